// microbench8.cu -- where does the 32-row tcgen05 GEMM (gemm_i8_tc_small_kernel) spend its time?  Synthetic buffers, no
// correctness check: event-timed launches over weights larger than L2, then one launch with SM-clock stamps of CTA (0,0):
// when the producer issued each k-step, when the MMA thread saw it complete, when the epilogue started / ended.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -lcuda -o scripts/microbench8 scripts/microbench8.cu
#include <cstdio>
#include <vector>
#include "../turboinfer_b200/csrc/gemm_tc.cuh"
using namespace tib;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

static int run(int K, int N, int S, bool resid) {
    const int KB = K / 128, tiles = N / 128, copies = 6;
    const size_t wbytes = (size_t)tiles * KB * kGemmTileBytes;
    uint8_t *wt, *xt; float *sx, *cs, *y; long long* sxf; unsigned long long* ws; unsigned int* cnt; long long* dbg;
    CK(cudaMalloc(&wt, wbytes * copies)); CK(cudaMemset(wt, 1, wbytes * copies));
    CK(cudaMalloc(&xt, (size_t)KB * 12288)); CK(cudaMemset(xt, 1, (size_t)KB * 12288));
    CK(cudaMalloc(&sx, 128)); CK(cudaMemset(sx, 0, 128)); CK(cudaMalloc(&sxf, 256)); CK(cudaMemset(sxf, 0, 256));
    CK(cudaMalloc(&cs, 4 * (size_t)N)); CK(cudaMemset(cs, 0, 4 * (size_t)N)); CK(cudaMalloc(&y, 4 * (size_t)N * 32)); CK(cudaMemset(y, 0, 4 * (size_t)N * 32));
    CK(cudaMalloc(&ws, 8 * (size_t)N * 32)); CK(cudaMemset(ws, 0, 8 * (size_t)N * 32)); CK(cudaMalloc(&cnt, 4 * tiles)); CK(cudaMemset(cnt, 0, 4 * tiles));
    CK(cudaMalloc(&dbg, 8 * 256)); CK(cudaMemset(dbg, 0, 8 * 256));
    CK(cudaFuncSetAttribute(gemm_i8_tc_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmemBytes));
    GemmArgs g{};
    g.M = 32; g.N = N; g.K = K; g.m_pad = 128; g.k_pad = K; g.a_signed_b = 0; g.woff = 8; g.sx = sx; g.sxf = sxf; g.colscale = cs; g.y = y;
    g.resid = resid ? y : nullptr;
    SplitKArgs sk{xt, ws, cnt, N, nullptr};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 30;
    for (int r = 0; r < 3; ++r) { g.wt = wt + (size_t)(r % copies) * wbytes; gemm_i8_tc_small_kernel<<<dim3(tiles, S), kSmallThreads, kSmallSmemBytes>>>(g, sk); }
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) { g.wt = wt + (size_t)(r % copies) * wbytes; gemm_i8_tc_small_kernel<<<dim3(tiles, S), kSmallThreads, kSmallSmemBytes>>>(g, sk); }
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("K=%5d N=%6d S=%d resid=%d: %7.2f us/launch (back to back), %6.0f GB/s of weights\n", K, N, S, (int)resid, ms * 1e3 / reps, wbytes / (ms * 1e-3 / reps) / 1e9);
    sk.dbg = dbg; g.wt = wt;
    gemm_i8_tc_small_kernel<<<dim3(tiles, S), kSmallThreads, kSmallSmemBytes>>>(g, sk);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(256);
    CK(cudaMemcpy(h.data(), dbg, 8 * 256, cudaMemcpyDeviceToHost));
    const long long t0 = h[64];
    const int nk = KB / S < 40 ? KB / S : 40;
    printf("   CTA(0,0), us since its first TMA issue @1.965GHz:  producer issue k-step:");
    for (int i = 0; i < nk; i += 4) printf(" %.2f", (h[64 + i] - t0) / 1965.0);
    printf("\n   MMA thread saw k-step complete:");
    for (int i = 0; i < nk; i += 4) printf(" %.2f", (h[128 + i] - t0) / 1965.0);
    printf("\n   epilogue: waits from %.2f, accumulators ready %.2f, done %.2f\n", (h[0] - t0) / 1965.0, (h[1] - t0) / 1965.0, (h[2] - t0) / 1965.0);
    cudaFree(wt); cudaFree(xt); cudaFree(y); cudaFree(ws);
    return 0;
}
int main() {
    run(4096, 4096, 1, true);
    run(4096, 4096, 1, false);
    run(4096, 12288, 1, false);
    run(4096, 22016 / 128 * 128, 1, false);
    run(11008 / 128 * 128, 4096, 3, true);
    return 0;
}
