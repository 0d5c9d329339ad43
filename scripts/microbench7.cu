// microbench7.cu -- latency of mbarrier.try_wait on an ALREADY completed phase, as the decode kernel's consumers see it:
// 16 warps of a CTA polling the same shared-memory barrier.  Also test_wait, and a plain volatile shared-memory flag.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/microbench7 scripts/microbench7.cu
#include <cstdio>
#include <cstdint>
#include "../turboinfer_b200/csrc/ptx.cuh"
using namespace tib;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k(long long* out, int bulk) {
    __shared__ uint64_t bar[4];
    __shared__ volatile int flag;
    __shared__ __align__(128) uint8_t buf[4096];
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); flag = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (bulk) { mbar_arrive_expect_tx(&bar[0], 4096); bulk_g2s(buf, out + 1024, 4096, &bar[0]); }   // completed by the TMA engine
        else mbar_arrive(&bar[0]);
        mbar_arrive(&bar[1]);
        flag = 1;
    }
    __syncthreads();
    for (volatile int spin = 0; spin < 20000; ++spin) {}   // let everything land
    __syncthreads();
    long long t0 = clock64();
    mbar_wait(&bar[0], 0);
    long long t1 = clock64();
    const bool ok = mbar_test_wait(&bar[1], 0);
    long long t2 = clock64();
    while (flag == 0) {}
    long long t3 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) {
        out[(threadIdx.x >> 5) * 4 + 0] = t1 - t0;
        out[(threadIdx.x >> 5) * 4 + 1] = t2 - t1 + (ok ? 0 : 1000000);
        out[(threadIdx.x >> 5) * 4 + 2] = t3 - t2;
    }
}
int main() {
    long long* d; CK(cudaMalloc(&d, 8 * 4096)); long long h[64];
    for (int bulk = 0; bulk < 2; ++bulk) {
        k<<<148, 512>>>(d, bulk); CK(cudaDeviceSynchronize());
        k<<<148, 512>>>(d, bulk); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("phase completed by %s: cycles per warp  try_wait(+clock):", bulk ? "a bulk copy" : "a thread arrive");
        for (int w = 0; w < 16; ++w) printf(" %lld", h[w * 4]);
        printf("\n   test_wait:"); for (int w = 0; w < 16; ++w) printf(" %lld", h[w * 4 + 1]);
        printf("\n   volatile flag:"); for (int w = 0; w < 16; ++w) printf(" %lld", h[w * 4 + 2]);
        printf("\n");
    }
    return 0;
}
