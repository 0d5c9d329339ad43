// microbench2.cu -- B200 ground truth for the integer (dp4a) GEMV inner loop: pipe rate of IDP4A and the rate of
// candidate per-item loops out of shared memory (no HBM traffic), 16 warps per SM like the real consumers.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/microbench2 scripts/microbench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint4 lds128(const void* p) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ uint2 lds64(const void* p) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {  // a unsigned bytes, b signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- pipe rate of IDP4A: 8 independent chains per thread, 16 warps per SM --------------------------------
template <int OP>
__global__ void __launch_bounds__(512, 1) k_pipe(int* out, int iters, long long* cycles) {
    int acc[8];
    uint32_t a[8], b[8];
    for (int i = 0; i < 8; ++i) { acc[i] = i; a[i] = threadIdx.x * 2654435761u + i; b[i] = threadIdx.x * 40503u + 7 * i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) acc[i] = dp4a_uu(a[i], b[i], acc[i]);
            if (OP == 1) acc[i] = acc[i] * (int)a[i] + (int)b[i];            // IMAD
            if (OP == 2) acc[i] = (acc[i] & 0x0F0F0F0F) ^ (int)a[i];          // LOP3
            if (OP == 3) acc[i] = __funnelshift_r(acc[i], (int)a[i], 4);      // SHF
            if (OP == 4) { acc[i] = dp4a_uu(a[i] & 0x0F0F0F0Fu, b[i], acc[i]); a[i] += 0x01010101u; }  // LOP3 + DP4A + IADD
        }
    }
    const long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// ---- candidate item loops ------------------------------------------------------------------------------
// item = 512 B = 32 lanes x 16 B = 4 columns x 256 k (INT4): lane holds one 32-bit word (8 nibbles = 8 k) per column.
// x digits of the lane's 8 k: words d0a d0b d1a d1b (LDS.128) d2a d2b (LDS.64) per 256-k chunk.
// VARIANT 0: hi nibbles via shift (3 unpack instr / word), 12 accumulators
// VARIANT 1: hi nibbles kept in place (x16), 24 accumulators
// VARIANT 2: like 0 but only two x digits (16-bit activations)
template <int VARIANT, int ITEMS_PER_ROUND>
__global__ void __launch_bounds__(544, 1) k_item(int* out, int rounds, int flush_every, long long* cycles) {
    extern __shared__ uint8_t smraw[];
    uint8_t* ring = smraw;                    // 6 x 32 KiB of "weights"
    uint8_t* xd = smraw + 4 * 32768;          // digits: 64 chunks x 32 lanes x 32 B
    int* part = reinterpret_cast<int*>(xd + 64 * 32 * 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp >= 16) return;
    for (int i = tid; i < 4 * 32768 / 4; i += 512) reinterpret_cast<uint32_t*>(ring)[i] = i * 2654435761u;
    for (int i = tid; i < 64 * 32 * 8; i += 512) reinterpret_cast<uint32_t*>(xd)[i] = i * 40503u + 17;
    for (int i = tid; i < 1024; i += 512) part[i] = 0;
    asm volatile("bar.sync 1, 512;");
    int acc[VARIANT == 1 ? 24 : 12];
#pragma unroll
    for (int i = 0; i < (VARIANT == 1 ? 24 : 12); ++i) acc[i] = 0;
    const long long t0 = clock64();
    int chunk = warp;
    for (int r = 0; r < rounds; ++r) {
        const uint8_t* wbase = ring + (size_t)(r & 3) * 32768 + (size_t)warp * ITEMS_PER_ROUND * 512 + lane * 16;
        uint4 w[ITEMS_PER_ROUND];
#pragma unroll
        for (int g = 0; g < ITEMS_PER_ROUND; ++g) w[g] = lds128(wbase + g * 512);
#pragma unroll
        for (int g = 0; g < ITEMS_PER_ROUND; ++g) {
            const uint8_t* xp = xd + ((size_t)(chunk & 63) * 32 + lane) * 32;
            chunk++;
            const uint4 xa = lds128(xp);
            const uint2 xb = VARIANT == 2 ? make_uint2(0, 0) : lds64(xp + 16);
            const uint32_t ww[4] = {w[g].x, w[g].y, w[g].z, w[g].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t lo = ww[c] & 0x0F0F0F0Fu;
                if (VARIANT == 1) {
                    const uint32_t hi = ww[c] & 0xF0F0F0F0u;
                    acc[6 * c + 0] = dp4a_uu(lo, xa.x, acc[6 * c + 0]);
                    acc[6 * c + 1] = dp4a_uu(lo, xa.z, acc[6 * c + 1]);
                    acc[6 * c + 2] = dp4a_us(lo, xb.x, acc[6 * c + 2]);
                    acc[6 * c + 3] = dp4a_uu(hi, xa.y, acc[6 * c + 3]);
                    acc[6 * c + 4] = dp4a_uu(hi, xa.w, acc[6 * c + 4]);
                    acc[6 * c + 5] = dp4a_us(hi, xb.y, acc[6 * c + 5]);
                } else {
                    const uint32_t hi = (ww[c] >> 4) & 0x0F0F0F0Fu;
                    acc[3 * c + 0] = dp4a_uu(lo, xa.x, acc[3 * c + 0]);
                    acc[3 * c + 0] = dp4a_uu(hi, xa.y, acc[3 * c + 0]);
                    acc[3 * c + 1] = dp4a_uu(lo, xa.z, acc[3 * c + 1]);
                    acc[3 * c + 1] = dp4a_uu(hi, xa.w, acc[3 * c + 1]);
                    if (VARIANT != 2) {
                        acc[3 * c + 2] = dp4a_us(lo, xb.x, acc[3 * c + 2]);
                        acc[3 * c + 2] = dp4a_us(hi, xb.y, acc[3 * c + 2]);
                    }
                }
            }
        }
        if ((r + 1) % flush_every == 0) {
            // warp-reduce 12 accumulators (butterfly, no transposition trick: upper bound on the cost) and publish
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                int v = VARIANT == 1 ? acc[(i / 3) * 6 + i % 3] + (acc[(i / 3) * 6 + 3 + i % 3] >> 4) : acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == i) atomicAdd(&part[(r & 15) * 16 + i], v);
            }
#pragma unroll
            for (int i = 0; i < (VARIANT == 1 ? 24 : 12); ++i) acc[i] = 0;
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < (VARIANT == 1 ? 24 : 12); ++i) s += acc[i];
    out[blockIdx.x * 512 + tid] = s + part[tid];
    if (lane == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}

// INT8: item = 512 B = 4 columns x 128 k: lane word per column = 4 k; digits of the lane's 4 k: d0 d1 d2 pad (LDS.128)
template <int ITEMS_PER_ROUND>
__global__ void __launch_bounds__(544, 1) k_item8(int* out, int rounds, int flush_every, long long* cycles) {
    extern __shared__ uint8_t smraw[];
    uint8_t* ring = smraw;
    uint8_t* xd = smraw + 4 * 32768;
    int* part = reinterpret_cast<int*>(xd + 64 * 32 * 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp >= 16) return;
    for (int i = tid; i < 4 * 32768 / 4; i += 512) reinterpret_cast<uint32_t*>(ring)[i] = i * 2654435761u;
    for (int i = tid; i < 64 * 32 * 8; i += 512) reinterpret_cast<uint32_t*>(xd)[i] = i * 40503u + 17;
    for (int i = tid; i < 1024; i += 512) part[i] = 0;
    asm volatile("bar.sync 1, 512;");
    int acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0;
    const long long t0 = clock64();
    int chunk = warp;
    for (int r = 0; r < rounds; ++r) {
        const uint8_t* wbase = ring + (size_t)(r & 3) * 32768 + (size_t)warp * ITEMS_PER_ROUND * 512 + lane * 16;
        uint4 w[ITEMS_PER_ROUND];
#pragma unroll
        for (int g = 0; g < ITEMS_PER_ROUND; ++g) w[g] = lds128(wbase + g * 512);
#pragma unroll
        for (int g = 0; g < ITEMS_PER_ROUND; ++g) {
            const uint4 xa = lds128(xd + ((size_t)(chunk & 127) * 32 + lane) * 16);
            chunk++;
            const uint32_t ww[4] = {w[g].x, w[g].y, w[g].z, w[g].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[3 * c + 0] = dp4a_uu(ww[c], xa.x, acc[3 * c + 0]);
                acc[3 * c + 1] = dp4a_uu(ww[c], xa.y, acc[3 * c + 1]);
                acc[3 * c + 2] = dp4a_us(ww[c], xa.z, acc[3 * c + 2]);
            }
        }
        if ((r + 1) % flush_every == 0) {
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                int v = acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == i) atomicAdd(&part[(r & 15) * 16 + i], v);
                acc[i] = 0;
            }
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += acc[i];
    out[blockIdx.x * 512 + tid] = s + part[tid];
    if (lane == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    int* out; long long* cyc; long long h;
    CK(cudaMalloc(&out, sizeof(int) * sms * 544));
    CK(cudaMalloc(&cyc, 8));
    const int iters = 20000;
    const char* names[] = {"IDP4A", "IMAD", "LOP3", "SHF", "LOP3+IDP4A+IADD"};
#define RUN_OP(OP) { k_pipe<OP><<<sms, 512>>>(out, 100, cyc); k_pipe<OP><<<sms, 512>>>(out, iters, cyc); CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); \
    printf("%-18s %7.3f cycles per warp-op(group) per SMSP (4 warps/SMSP)\n", names[OP], (double)h / ((double)iters * 8 * 4)); }
    RUN_OP(0) RUN_OP(1) RUN_OP(2) RUN_OP(3) RUN_OP(4)
    const size_t smem = 4 * 32768 + 64 * 32 * 32 + 4096;
#define RUN_ITEM(VAR, IPR, FLUSH) { \
        CK(cudaFuncSetAttribute(k_item<VAR, IPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
        k_item<VAR, IPR><<<sms, 544, smem>>>(out, 16, FLUSH, cyc); CK(cudaMemset(cyc, 0, 8)); \
        k_item<VAR, IPR><<<sms, 544, smem>>>(out, 4096, FLUSH, cyc); \
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); \
        const double cpi = h / (4096.0 * IPR * 16); \
        printf("int4 item loop variant %d, %d items/round, flush every %3d rounds: %6.2f cycles/item/SM -> %6.0f GB/s at 1.9 GHz\n", VAR, IPR, FLUSH, cpi, 512.0 / cpi * 1.9 * sms); }
    RUN_ITEM(0, 4, 4) RUN_ITEM(0, 4, 2) RUN_ITEM(0, 4, 1) RUN_ITEM(0, 4, 64) RUN_ITEM(0, 2, 8) RUN_ITEM(1, 4, 4) RUN_ITEM(1, 4, 64) RUN_ITEM(2, 4, 4)
#define RUN_ITEM8(IPR, FLUSH) { \
        CK(cudaFuncSetAttribute(k_item8<IPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
        k_item8<IPR><<<sms, 544, smem>>>(out, 16, FLUSH, cyc); CK(cudaMemset(cyc, 0, 8)); \
        k_item8<IPR><<<sms, 544, smem>>>(out, 4096, FLUSH, cyc); \
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); \
        const double cpi = h / (4096.0 * IPR * 16); \
        printf("int8 item loop, %d items/round, flush every %3d rounds: %6.2f cycles/item/SM -> %6.0f GB/s at 1.9 GHz\n", IPR, FLUSH, cpi, 512.0 / cpi * 1.9 * sms); }
    RUN_ITEM8(4, 8) RUN_ITEM8(4, 64)
    return 0;
}
