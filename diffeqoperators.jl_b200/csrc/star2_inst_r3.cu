// Instantiates the persistent tiled kernel for stencil radius 3 (Float32 and Float64, 3-D tile and 2-D strip).
#include "kernel_star2.cuh"

namespace deo {
DEO_STAR2_INSTANTIATE(3)
}  // namespace deo
