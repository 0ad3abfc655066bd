// Microbenchmark: DFMA / FFMA dependent-issue latency and per-scheduler throughput on this GPU, as a function of the number
// of independent chains per warp and of warps per scheduler (the stencil kernels' inner loops are chains of 2R+1 FMAs).
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int CHAINS>
__global__ void k_chain(T* out, T a, T b, int iters, long long* cyc) {
    T acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * T(1e-3) + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    const long long t1 = clock64();
    T s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <typename T, int CHAINS>
void run(const char* name, int threads, T* out, long long* cyc, int sms) {
    const int iters = 2048;
    k_chain<T, CHAINS><<<sms, threads>>>(out, (T)1.0000001, (T)1e-9, iters, cyc);
    cudaDeviceSynchronize();
    k_chain<T, CHAINS><<<sms, threads>>>(out, (T)1.0000001, (T)1e-9, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp = (double)c / (8.0 * iters * CHAINS);            // cycles between two FMA issues of one warp
    const double warps_per_sched = threads / 32 / 4.0;
    printf("%s chains=%2d warps/scheduler=%4.1f : %6.2f cycles per FMA per warp, %6.2f cycles per warp-FMA per scheduler\n", name, CHAINS,
           warps_per_sched, per_warp, per_warp / warps_per_sched);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, 8 * 148 * 1024 * 2);
    long long* cyc; cudaMalloc(&cyc, 8);
    printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
    for (int threads : {128, 256, 512, 1024}) {
        run<double, 1>("DFMA", threads, out, cyc, p.multiProcessorCount);
        run<double, 2>("DFMA", threads, out, cyc, p.multiProcessorCount);
        run<double, 4>("DFMA", threads, out, cyc, p.multiProcessorCount);
        run<double, 8>("DFMA", threads, out, cyc, p.multiProcessorCount);
    }
    for (int threads : {128, 512}) {
        run<float, 1>("FFMA", threads, (float*)out, cyc, p.multiProcessorCount);
        run<float, 2>("FFMA", threads, (float*)out, cyc, p.multiProcessorCount);
        run<float, 4>("FFMA", threads, (float*)out, cyc, p.multiProcessorCount);
        run<float, 8>("FFMA", threads, (float*)out, cyc, p.multiProcessorCount);
    }
    return 0;
}
