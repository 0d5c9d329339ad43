// microbench4.cu -- the floor of one phase hand-over on a B200: 148 resident CTAs, each iteration
//   (a) every CTA stores a slice of a vector, (b) release-arrives on a counter, (c) one thread polls with acquire,
//   (d) every thread loads its share of the WHOLE vector (written by the other SMs) through L2.
// Variants: barrier only; barrier + exchange; exchange with 4 B / 16 B loads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/microbench4 scripts/microbench4.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_rel(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void red_rlx(unsigned* p) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }

template <int MODE>   // 0: barrier only, 1: + store slice before / load whole vector after, 2: like 1 but relaxed arrive (no fence)
__global__ void __launch_bounds__(512, 1) k_phase(unsigned* bar, float* vec, int H, int iters, long long* cycles, float* sink) {
    const int tid = threadIdx.x;
    unsigned target = 0;
    float acc = 0.f;
    const int per = H / gridDim.x;   // slice per CTA
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE >= 1 && tid < per) vec[blockIdx.x * per + tid] = acc + (float)i;
        asm volatile("bar.sync 1, 512;");
        if (tid == 0) {
            if (MODE == 2) red_rlx(bar); else red_rel(bar);
            target += gridDim.x;
            while (ld_acq(bar) < target) {}
        } else target += gridDim.x;
        asm volatile("bar.sync 1, 512;");
        if (MODE >= 1) {
            const float4* v4 = reinterpret_cast<const float4*>(vec);
            for (int v = tid; v < H / 4; v += 512) { const float4 t = __ldcg(v4 + v); acc += t.x + t.w; }
        }
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    sink[blockIdx.x * 512 + tid] = acc;
}


__device__ __forceinline__ float4 ld_vol4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
#define SENT 0x7FC0DEADu
// Lamport hand-over: no barrier, no flag, no fence -- the data words are their own flags.  Three rotating buffers,
// cleared to a sentinel two iterations before they are written again.
__global__ void __launch_bounds__(512, 1) k_lamport(float* bufs, int H, int Hpad, int iters, long long* cycles, float* sink) {
    const int tid = threadIdx.x;
    float acc = 0.f;
    const int per = H / gridDim.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        float* cur = bufs + (size_t)(i % 3) * Hpad;
        float* old = bufs + (size_t)((i + 2) % 3) * Hpad;
        if (tid < per) cur[blockIdx.x * per + tid] = (float)(i & 1023) + 0.001f * acc;
        const float4* v4 = reinterpret_cast<const float4*>(cur);
        for (int v = tid; v < H / 4; v += 512) {
            float4 t;
            do { t = ld_vol4(v4 + v); } while (__float_as_uint(t.x) == SENT || __float_as_uint(t.y) == SENT || __float_as_uint(t.z) == SENT || __float_as_uint(t.w) == SENT);
            acc += t.x - t.w;
        }
        asm volatile("bar.sync 1, 512;");   // (the real kernel has CTA-wide syncs between reading x and writing y too)
        if (tid < per) old[blockIdx.x * per + tid] = __uint_as_float(SENT);
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    sink[blockIdx.x * 512 + tid] = acc;
}
// flag slots instead of an atomic counter: every CTA release-stores its iteration number into its own word; one warp
// per CTA polls the 148 words, then the CTA loads the vector.
template <int FENCE>
__global__ void __launch_bounds__(512, 1) k_flags(unsigned* flags, float* vec, int H, int iters, long long* cycles, float* sink) {
    const int tid = threadIdx.x, lane = tid & 31;
    float acc = 0.f;
    const int per = H / gridDim.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (tid < per) vec[blockIdx.x * per + tid] = acc + (float)i;
        asm volatile("bar.sync 1, 512;");
        if (tid < 32) {
            if (tid == 0) {
                if (FENCE) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(i + 1) : "memory");
                else asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(i + 1) : "memory");
            }
            bool ok;
            do {
                ok = true;
                for (int c = lane; c < (int)gridDim.x; c += 32) ok &= ld_rlx(flags + c) >= (unsigned)(i + 1);
            } while (!__all_sync(0xffffffffu, ok));
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        asm volatile("bar.sync 1, 512;");
        const float4* v4 = reinterpret_cast<const float4*>(vec);
        for (int v = tid; v < H / 4; v += 512) { const float4 t = __ldcg(v4 + v); acc += t.x + t.w; }
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    sink[blockIdx.x * 512 + tid] = acc;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned* bar; float *vec, *sink; long long* cyc; long long h;
    CK(cudaMalloc(&bar, 256)); CK(cudaMalloc(&vec, 4 * 16384)); CK(cudaMalloc(&sink, 4 * sms * 512)); CK(cudaMalloc(&cyc, 8));
    const int iters = 2000;
    for (int H : {2048, 4096, 11008}) {
#define RUN(MODE, label) { CK(cudaMemset(bar, 0, 256)); int Hh = H / sms * sms / 4 * 4; Hh = (H / (4 * sms)) * 4 * sms; \
        void* args[] = {&bar, &vec, &Hh, (void*)&iters, &cyc, &sink}; \
        CK(cudaLaunchCooperativeKernel((void*)k_phase<MODE>, dim3(sms), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize()); \
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); \
        printf("H=%5d %-44s %7.0f cycles = %.2f us per hand-over\n", Hh, label, (double)h / iters, (double)h / iters / 1965.0); }
        RUN(0, "barrier only (release add + acquire poll)")
        RUN(1, "store slice + barrier + load whole vector")
        RUN(2, "same, relaxed arrive (no fence)")
        {
            int Hh = (H / (4 * sms)) * 4 * sms, Hpad = 16384;
            float* bufs; CK(cudaMalloc(&bufs, 4 * 3 * Hpad));
            unsigned* init = new unsigned[3 * Hpad]; for (int i = 0; i < 3 * Hpad; ++i) init[i] = SENT;
            CK(cudaMemcpy(bufs, init, 4 * 3 * Hpad, cudaMemcpyHostToDevice)); delete[] init;
            void* args[] = {&bufs, &Hh, &Hpad, (void*)&iters, &cyc, &sink};
            CK(cudaLaunchCooperativeKernel((void*)k_lamport, dim3(sms), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("H=%5d %-44s %7.0f cycles = %.2f us per hand-over\n", Hh, "Lamport (sentinel data words, no barrier)", (double)h / iters, (double)h / iters / 1965.0);
            cudaFree(bufs);
            unsigned* flags; CK(cudaMalloc(&flags, 4096)); 
            void* a2[] = {&flags, &vec, &Hh, (void*)&iters, &cyc, &sink};
            CK(cudaMemset(flags, 0, 4096));
            CK(cudaLaunchCooperativeKernel((void*)k_flags<1>, dim3(sms), dim3(512), a2, 0, 0)); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("H=%5d %-44s %7.0f cycles = %.2f us per hand-over\n", Hh, "flag words (release store, one warp polls)", (double)h / iters, (double)h / iters / 1965.0);
            CK(cudaMemset(flags, 0, 4096));
            CK(cudaLaunchCooperativeKernel((void*)k_flags<0>, dim3(sms), dim3(512), a2, 0, 0)); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
            printf("H=%5d %-44s %7.0f cycles = %.2f us per hand-over\n", Hh, "flag words, relaxed store (no fence)", (double)h / iters, (double)h / iters / 1965.0);
            cudaFree(flags);
        }
    }
    return 0;
}
