// Do the FP64 tensor pipe (DMMA.8x8x4) and the FP64 FMA pipe (DFMA) add up on sm_100a, or do they share the units?
// Register-only loops, 16 warps per SM (2 CTAs x 8 warps): all-DMMA, all-DFMA, and half/half.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_pipes_probe tools/fp64_pipes_probe.cu && build/fp64_pipes_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode: 0 all warps DMMA, 1 all warps DFMA, 2 even warps DMMA / odd warps DFMA, 3 warps 0-3 DMMA / 4-7 DFMA
__global__ void __launch_bounds__(256, 2) probe(int mode, int iters, double* out, double seed) {
    const int warp = threadIdx.x >> 5;
    const bool tensor = mode == 0 || (mode == 2 && (warp & 1) == 0) || (mode == 3 && warp < 4);
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i * 1e-3;
    if (tensor) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma884(c[i], c[i + 1], a, b);      // 8 independent DMMAs per iteration
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);            // 64 independent-ish DFMAs per iteration
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}
int main() {
    double* d; cudaMalloc(&d, 8);
    const int iters = 20000, ctas = 148 * 2;
    for (int mode = 0; mode < 4; ++mode) {
        probe<<<ctas, 256>>>(mode, 100, d, 0.5);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<<<ctas, 256>>>(mode, iters, d, 0.5);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        // flops: DMMA warp-iteration = 8 x 512 flop; DFMA warp-iteration = 64 x 32 lanes x 2 flop = 4096 flop -> identical per iteration
        const double warps_t = (mode == 0) ? 8 : (mode == 1 ? 0 : 4), warps_f = 8 - warps_t;
        const double ft = warps_t * ctas * (double)iters * 4096.0, ff = warps_f * ctas * (double)iters * 4096.0;
        printf("mode %d: %.3f ms  tensor %.1f TF/s  fma %.1f TF/s  total %.1f TF/s\n", mode, ms, ft / ms / 1e9, ff / ms / 1e9, (ft + ff) / ms / 1e9);
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
