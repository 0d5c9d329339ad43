// microbench5.cu -- is the legacy warp-level integer MMA (mma.sync m16n8k32 u8 x s8 -> s32, SASS IMMA.16832) fast
// enough on a B200 to carry the decode GEMV's main loop?  Measures (1) the issue rate and latency of the instruction and
// (2) a candidate inner loop: LDS.128 of packed INT4 weights + 8 LOP3 + LDS.128 of activation digits + 2 IMMA per
// 512-byte item, out of shared memory, 16 warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o scripts/microbench5 scripts/microbench5.cu
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void imma_u8s8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// CHAINS independent accumulator chains per warp, `iters` MMAs per chain
template <int CHAINS>
__global__ void k_imma_rate(int iters, long long* cycles, int* sink) {
    int acc[CHAINS][4];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[c][i] = 0;
    const uint32_t a = threadIdx.x * 0x01010101u, b = threadIdx.x * 0x00010203u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) imma_u8s8(acc[c], a, a + c, a ^ 5u, a + 7u, b, b + c);
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    if (s == 0x12345) sink[0] = s;
    if (threadIdx.x == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}

// candidate main loop: `items` 512-byte items per warp out of a shared-memory ring of `ring_items` items per warp
template <int MODE>   // 0: full loop, 1: no MMA (loads + unpack), 2: loads only
__global__ void __launch_bounds__(544, 1) k_loop(int quads, long long* cycles, int* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp >= 16) return;
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem);            // 160 KiB
    uint32_t* xd = reinterpret_cast<uint32_t*>(smem + 160 * 1024);   // 12 KiB of digits
    for (int i = tid; i < 160 * 1024 / 4; i += 512) ring[i] = i * 2654435761u;
    for (int i = tid; i < 12 * 1024 / 4; i += 512) xd[i] = i * 40503u;
    asm volatile("bar.sync 1, 512;");
    const int g = lane >> 2, t = lane & 3;
    const uint32_t xbase = (uint32_t)__cvta_generic_to_shared(xd) + ((g < 3 ? g : 0) * 4 + t) * 16;
    const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(ring) + lane * 16;
    int accL[2][4] = {}, accH[2][4] = {};
    int junk = 0;
    const long long t0 = clock64();
    for (int q = 0; q < quads; ++q) {
        const int st = q % 5;                       // 5 stages of 32 KiB; warp w's quad at offset w * 2 KiB
        const uint32_t wq = wbase + st * 32768 + warp * 2048;
        const uint32_t xq = xbase + (q & 15) * 768;
        uint4 w[4], x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[i].x), "=r"(w[i].y), "=r"(w[i].z), "=r"(w[i].w) : "r"(wq + i * 512));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[i].x), "=r"(x[i].y), "=r"(x[i].z), "=r"(x[i].w) : "r"(xq + i * 192));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (MODE == 2) { junk ^= w[i].x ^ w[i].y ^ w[i].z ^ w[i].w ^ x[i].x ^ x[i].y ^ x[i].z ^ x[i].w; continue; }
            const uint32_t l0 = w[i].x & 0x0F0F0F0Fu, l1 = w[i].y & 0x0F0F0F0Fu, l2 = w[i].z & 0x0F0F0F0Fu, l3 = w[i].w & 0x0F0F0F0Fu;
            const uint32_t h0 = w[i].x & 0xF0F0F0F0u, h1 = w[i].y & 0xF0F0F0F0u, h2 = w[i].z & 0xF0F0F0F0u, h3 = w[i].w & 0xF0F0F0F0u;
            if (MODE == 1) { junk ^= l0 ^ l1 ^ l2 ^ l3 ^ h0 ^ h1 ^ h2 ^ h3 ^ x[i].x ^ x[i].y ^ x[i].z ^ x[i].w; continue; }
            imma_u8s8(accL[i & 1], l0, l1, l2, l3, x[i].x, x[i].y);
            imma_u8s8(accH[i & 1], h0, h1, h2, h3, x[i].z, x[i].w);
        }
    }
    const long long t1 = clock64();
    int s = junk;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += accL[0][i] + accL[1][i] + accH[0][i] + accH[1][i];
    if (s == 0x12345) sink[0] = s;
    if (lane == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}

template <int CHAINS>
static int rate(int warps, long long* cyc, int* sink) {
    const int iters = 4096;
    long long h;
    CK(cudaMemset(cyc, 0, 8));
    k_imma_rate<CHAINS><<<148, warps * 32>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double per_sm = (double)h / ((double)iters * CHAINS * warps);
    printf("IMMA.16832 u8*s8: %2d warps/SM x %d chains: %.2f cycles per MMA per SM (%.2f per SM sub-partition), %.0f int-ops/clk/SM; dependent-chain step %.1f cycles\n",
           warps, CHAINS, per_sm, per_sm * 4, 16.0 * 8 * 32 * 2 / per_sm, (double)h / iters);
    return 0;
}

template <int MODE>
static int loop(const char* label, long long* cyc, int* sink) {
    const int quads = 2000;
    long long h;
    CK(cudaFuncSetAttribute(k_loop<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 172 * 1024));
    CK(cudaMemset(cyc, 0, 8));
    k_loop<MODE><<<148, 544, 172 * 1024>>>(quads, cyc, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double per_item = (double)h / ((double)quads * 4 * 16);
    printf("main loop candidate, %-22s: %.2f cycles per 512-B item per SM -> %.0f GB/s-equivalent at 1.9 GHz on 148 SMs\n", label, per_item,
           512.0 / per_item * 1.9 * 148);
    return 0;
}

int main() {
    long long* cyc; int* sink;
    CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&sink, 4));
    rate<1>(4, cyc, sink); rate<2>(4, cyc, sink); rate<4>(4, cyc, sink);
    rate<1>(16, cyc, sink); rate<2>(16, cyc, sink); rate<4>(16, cyc, sink);
    loop<2>("loads only", cyc, sink);
    loop<1>("loads + unpack", cyc, sink);
    loop<0>("loads + unpack + IMMA", cyc, sink);
    return 0;
}
