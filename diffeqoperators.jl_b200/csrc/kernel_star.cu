// placeholder: the tiled kernel is added in a later step
#include "common.hpp"
namespace deo {
int32_t star_configure(deo_plan*) { return DEO_OK; }
int32_t launch_star(const deo_plan*, void*, const void*, long long, long long, cudaStream_t) {
    set_error("star kernel not built");
    return DEO_ERR_UNSUPPORTED;
}
}  // namespace deo
