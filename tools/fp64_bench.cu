// Microbenchmark: peak FP64 FMA rate on this GPU (vector DFMA with register / constant operands, DMMA m8n8k4).
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cw[16];
template <int CHAINS, bool CONSTOP>
__global__ void k_dfma(double* out, double a, double b, int iters) {
    double acc[CHAINS];
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = CONSTOP ? fma(acc[i], cw[i & 15], b) : fma(acc[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    for (int it = 0; it < iters; ++it) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 4096;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    double h[16]; for (int i = 0; i < 16; ++i) h[i] = 1.0 + 1e-9 * i; cudaMemcpyToSymbol(cw, h, sizeof h);
    printf("%s  SMs=%d  clock=%d MHz\n", p.name, sms, p.clockRate / 1000);
    auto report = [&](const char* name, float ms, double fma_per_thread) {
        double fmas = fma_per_thread * blocks * threads;
        printf("%-34s %8.3f ms  %7.2f TFLOPS  %6.1f FMA/clk/SM (at %d MHz)\n", name, ms, 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / (p.clockRate * 1e3), p.clockRate / 1000);
    };
    report("DFMA reg operands, 8 chains", timeit([&] { k_dfma<8, false><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }), 8.0 * iters);
    report("DFMA reg operands, 16 chains", timeit([&] { k_dfma<16, false><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }), 16.0 * iters);
    report("DFMA const operand, 8 chains", timeit([&] { k_dfma<8, true><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); }), 8.0 * iters);
    report("DMMA m8n8k4 (4 indep accum)", timeit([&] { k_dmma<<<blocks, threads>>>(out, iters); }), 4.0 * iters * 8 * 8 * 4 / 32.0);
    return 0;
}
