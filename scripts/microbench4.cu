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
    }
    return 0;
}
