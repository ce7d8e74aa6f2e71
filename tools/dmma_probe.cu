// DMMA.8x8x4 issue behaviour on sm_100a: dependent-issue latency and throughput as a function of resident warps per SM
// and of independent accumulator chains per warp (the block solver's main loop has 16 chains, its TRSM 8, and only
// 2-4 warps per SM sub-partition).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_probe tools/dmma_probe.cu && build/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void probe(int iters, double* out, double seed, long long* clk) {
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
    double c[2 * ILP];
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) c[i] = i * 1e-3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16 / ILP; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);      // 16 DMMAs per iteration
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) clk[0] = t1 - t0;
}
template <int ILP>
void run(int warps_per_sm, double* d, long long* dc) {
    const int iters = 4000;
    const int threads = 32 * (warps_per_sm >= 8 ? 8 : warps_per_sm);
    const int ctas = 148 * (warps_per_sm >= 8 ? warps_per_sm / 8 : 1);
    probe<ILP><<<ctas, threads>>>(10, d, 0.5, dc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<ILP><<<ctas, threads>>>(iters, d, 0.5, dc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk; cudaMemcpy(&clk, dc, 8, cudaMemcpyDeviceToHost);
    const double n_dmma = (double)iters * 16;
    const double tf = (double)ctas * (threads / 32) * n_dmma * 512.0 / ms / 1e9;
    printf("warps/SM %2d  ILP %2d : %7.1f clk per DMMA per warp   %6.2f TF/s\n", warps_per_sm, ILP, clk / n_dmma, tf);
}
int main() {
    double* d; cudaMalloc(&d, 8);
    long long* dc; cudaMalloc(&dc, 8);
    for (int w : {4, 8, 16, 32}) {
        run<1>(w, d, dc); run<2>(w, d, dc); run<4>(w, d, dc); run<8>(w, d, dc); run<16>(w, d, dc);
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
