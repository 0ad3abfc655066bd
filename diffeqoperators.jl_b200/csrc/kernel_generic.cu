// Per-point kernel: one thread per output point, every operator / BC / layout the plan format can
// express.  It is the always-correct path (odd shapes, tiny grids, per-face BC tables, operators the
// tiled kernel does not take) and the in-kernel reference the tiled kernel is tested against.
//
// du[c] = sum_k  row_k(c[axis_k]) . q_k      with q_k the ghost-padded pencil through c along axis_k:
//   q[0]   = ghost_lo = b_l + a_l . u[0:K_l]          (bc_operators.jl:188-191; computed here, never stored)
//   q[j+1] = u[j]
//   q[n+1] = ghost_hi = b_r + a_r . u[n-K_r:n]
// or q taken straight from a pre-padded input (derivative_operator_functions.jl:50-57).
#include "generic_device.cuh"

namespace deo {

template <typename T>
__global__ void __launch_bounds__(256)
k_generic(const __grid_constant__ DevPlan<T> P, const T* __restrict__ u, T* __restrict__ du, int z0, int z1) {
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c0 >= P.n_out[0]) return;
    for (int c2 = z0 + blockIdx.z; c2 < z1; c2 += gridDim.z) {
        for (int c1 = blockIdx.y * blockDim.y + threadIdx.y; c1 < P.n_out[1]; c1 += gridDim.y * blockDim.y) {
            const long long o = (long long)c0 * P.out_stride[0] + (long long)c1 * P.out_stride[1] + (long long)c2 * P.out_stride[2];
            T tot = P.accumulate ? du[o] : T(0);
            for (int k = 0; k < P.nops; ++k) {
                const T r = apply_op<T>(P, P.ops[k], u, c0, c1, c2);
                tot = (k == 0 && !P.accumulate) ? r : tot + r;   // sum(op -> op*x, ops): left fold
            }
            du[o] = tot;
        }
    }
}

template <typename T>
static int32_t launch_generic_t(const deo_plan* plan, void* du, const void* u, long long z0, long long z1, cudaStream_t s) {
    const DevPlan<T>& P = *reinterpret_cast<const DevPlan<T>*>(plan->devplan.data());
    const long long n0 = P.n_out[0], n1 = P.n_out[1];
    dim3 block, grid;
    if (n1 == 1) block = dim3(256, 1, 1);
    else if (n0 >= 64) block = dim3(64, 4, 1);
    else if (n0 >= 16) block = dim3(16, 16, 1);
    else block = dim3(4, 64, 1);
    grid.x = (unsigned)((n0 + block.x - 1) / block.x);
    long long gy = (n1 + block.y - 1) / block.y;
    grid.y = (unsigned)(gy > 65535 ? 65535 : gy);
    long long gz = z1 - z0;
    grid.z = (unsigned)(gz > 65535 ? 65535 : gz);
    k_generic<T><<<grid, block, 0, s>>>(P, (const T*)u, (T*)du, (int)z0, (int)z1);
    DEO_CUDA(cudaGetLastError());
    return DEO_OK;
}

int32_t launch_generic(const deo_plan* plan, void* du, const void* u, long long z0, long long z1, cudaStream_t s) {
    return plan->dtype == DEO_F64 ? launch_generic_t<double>(plan, du, u, z0, z1, s)
                                  : launch_generic_t<float>(plan, du, u, z0, z1, s);
}

}  // namespace deo
