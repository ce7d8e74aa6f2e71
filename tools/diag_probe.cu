// Phase timing of chol_diag_kernel's body (clock64 stamps by thread 0 of one CTA).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/diag_probe tools/diag_probe.cu && build/diag_probe
#include <cstdio>
#include <cmath>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
__device__ long long g_stamps[32];
#define DIAG_STAMP(n) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_stamps[(n)] = clock64(); } while (0)
#include "../dbslmm_b200/csrc/chol.cu"
using namespace dbslmm;
// the 128-thread form of the diagonal-tile code (as run inside the 64-row panel kernel)
__global__ void __launch_bounds__(128, 3) diag128_probe_kernel(const BlockDesc* blocks, const int32_t* items, int32_t k, const double* sigma,
                                                               double* Lbuf, double* wbuf, double ridge, int32_t* status) {
    extern __shared__ __align__(16) double smem[];
    const int blk = items[blockIdx.x];
    const BlockDesc bd = blocks[blk];
    diag_body<128>(bd, blk, k, sigma, Lbuf, wbuf, ridge, status, smem, []() {});
}
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
    cudaFuncSetAttribute(diag128_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DIAG);
    for (int nt : {256, 128}) {
    for (int rep = 0; rep < 3; ++rep) {
        for (int n : {1, NBLK, 3 * 148}) {
            if (n > NBLK && nt == 256) continue;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (nt == 256) launch_chol_diag(dB, dI, n, 0, dS, dL, dW, 64 * 64 * NBLK, 0.1, dSt, 0);
            else diag128_probe_kernel<<<std::min(n, NBLK), 128, SMEM_DIAG>>>(dB, dI, 0, dS, dL, dW, 0.1, dSt);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long st[32]; cudaMemcpyFromSymbol(st, g_stamps, sizeof st);
            printf("threads %d rep %d ctas %d: kernel %.2f us; stamps (clk from start):", nt, rep, std::min(n, NBLK), ms * 1e3);
            for (int i = 1; i < 24; ++i) if (st[i]) printf(" [%d]%lld", i, st[i] - st[0]);
            printf("\n");
            cudaMemset(dSt, 0, 4);
        }
    }
    // ---- check L (lower), W^T (strict upper of the tile) and the W image against a host factorisation
    {
        std::vector<double> A(64 * 64), L(64 * 64, 0.0), W(64 * 64, 0.0);
        for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) A[i * 64 + j] = S[(size_t)std::max(i, j) * mp + std::min(i, j)] + (i == j ? 0.1 : 0.0);
        for (int j = 0; j < 64; ++j) {
            double d = A[j * 64 + j];
            for (int k2 = 0; k2 < j; ++k2) d -= L[j * 64 + k2] * L[j * 64 + k2];
            L[j * 64 + j] = sqrt(d);
            for (int i = j + 1; i < 64; ++i) {
                double v = A[i * 64 + j];
                for (int k2 = 0; k2 < j; ++k2) v -= L[i * 64 + k2] * L[j * 64 + k2];
                L[i * 64 + j] = v / L[j * 64 + j];
            }
        }
        for (int c = 0; c < 64; ++c)
            for (int i = c; i < 64; ++i) {
                double v = (i == c) ? 1.0 : 0.0;
                for (int k2 = c; k2 < i; ++k2) v -= L[i * 64 + k2] * W[k2 * 64 + c];
                W[i * 64 + c] = v / L[i * 64 + i];
            }
        std::vector<double> dl((size_t)nrows * mp), dw(64 * 64);
        cudaMemcpy(dl.data(), dL + (size_t)5 * S.size(), sizeof(double) * dl.size(), cudaMemcpyDeviceToHost);   // block 5
        cudaMemcpy(dw.data(), dW + (size_t)5 * 64 * 64, sizeof(double) * 64 * 64, cudaMemcpyDeviceToHost);
        double eL = 0, eWt = 0, eW = 0;
        for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) {
            if (j <= i) eL = fmax(eL, fabs(dl[(size_t)i * mp + j] - L[i * 64 + j]));
            else eWt = fmax(eWt, fabs(dl[(size_t)i * mp + j] - W[j * 64 + i]));
            if ((j >> 4) <= (i >> 4)) eW = fmax(eW, fabs(dw[w_img_off(i, j)] - W[i * 64 + j]));
        }
        printf("threads %d max abs err: L %.3e  W^T (tile upper) %.3e  W (image) %.3e\n", nt, eL, eWt, eW);
    }
    cudaMemset(dL, 0, sizeof(double) * S.size() * NBLK); cudaMemset(dW, 0, sizeof(double) * 64 * 64 * NBLK * 2);
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
