// Phase timing of chol_diag_kernel's body (clock64 stamps by thread 0 of one CTA).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/diag_probe tools/diag_probe.cu && build/diag_probe
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__device__ long long g_stamps[32];
#define DIAG_STAMP(n) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_stamps[(n)] = clock64(); } while (0)
#include "../dbslmm_b200/csrc/chol.cu"
using namespace dbslmm;
int main() {
    const int m = 256, mp = 256, nrows = mp + 8;
    std::vector<double> S((size_t)nrows * mp, 0.0);
    for (int i = 0; i < m; ++i) for (int j = 0; j <= i; ++j) S[(size_t)i * mp + j] = (i == j) ? 2.0 : 0.5 / (1.0 + i - j);
    BlockDesc bd{}; bd.moff = 0; bd.m = m; bd.ms = m; bd.mp = mp; bd.ld = mp; bd.nrows = nrows;
    const int NBLK = 296;
    std::vector<BlockDesc> blocks(NBLK, bd);
    std::vector<int> items(NBLK);
    for (int i = 0; i < NBLK; ++i) { items[i] = i; blocks[i].moff = (int64_t)i * nrows * mp; }
    double *dS, *dL, *dW; BlockDesc* dB; int *dI, *dSt;
    cudaMalloc(&dS, sizeof(double) * S.size() * NBLK); cudaMalloc(&dL, sizeof(double) * S.size() * NBLK);
    cudaMalloc(&dW, sizeof(double) * 64 * 64 * NBLK * 2);
    cudaMalloc(&dB, sizeof(BlockDesc) * NBLK); cudaMalloc(&dI, sizeof(int) * NBLK); cudaMalloc(&dSt, sizeof(int) * NBLK);
    for (int i = 0; i < NBLK; ++i) cudaMemcpy(dS + (size_t)i * S.size(), S.data(), sizeof(double) * S.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, blocks.data(), sizeof(BlockDesc) * NBLK, cudaMemcpyHostToDevice);
    cudaMemcpy(dI, items.data(), sizeof(int) * NBLK, cudaMemcpyHostToDevice);
    cudaMemset(dSt, 0, sizeof(int) * NBLK);
    chol_configure();
    for (int rep = 0; rep < 3; ++rep) {
        for (int n : {1, NBLK}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            launch_chol_diag(dB, dI, n, 0, dS, dL, dW, 64 * 64 * NBLK, 0.1, dSt, 0);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long st[32]; cudaMemcpyFromSymbol(st, g_stamps, sizeof st);
            printf("rep %d ctas %d: kernel %.2f us; stamps (clk from start):", rep, n, ms * 1e3);
            for (int i = 1; i < 24; ++i) if (st[i]) printf(" [%d]%lld", i, st[i] - st[0]);
            printf("\n");
            cudaMemset(dSt, 0, 4);
        }
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
