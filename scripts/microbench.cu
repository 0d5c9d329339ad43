// microbench.cu -- B200 ground truth for the GEMV design: issue rates of the ops in the unpack loop, the rate of the
// real per-item dot product out of shared memory, the bandwidth of the bulk-copy ring alone, grid-barrier latency.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o microbench scripts/microbench.cu
#include <cstdio>
#include <vector>
#include "../turboinfer_b200/csrc/mega.cuh"
using namespace tib;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void __launch_bounds__(512, 1) k_ops(float* out, int iters, long long* cycles) {
    float a[8], b[8];
    f32x2 p[8], q[8];
    uint32_t u[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = 1.0001f + i * 1e-4f; p[i] = pack2(a[i], b[i]); q[i] = pack2(b[i], a[i]); u[i] = threadIdx.x * 2654435761u + i; }
    const f32x2 c2 = pack2(0.999f, 1.001f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], b[i], 0.5f);
            if (OP == 1) a[i] = a[i] + b[i];
            if (OP == 2) p[i] = fma2(p[i], q[i], c2);
            if (OP == 3) p[i] = add2(p[i], c2);
            if (OP == 4) u[i] = and_or(u[i], 0xF0F0F0F0u, u[(i + 1) & 7]);
            if (OP == 5) { p[i] = fma2(add2(pack2u(and_or(u[i], 0xF0u, 0x4B000000u), and_or(u[i], 0xF00u, 0x4B000000u)), c2), q[i], p[i]); }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) { float x, y; unpack2(p[i], x, y); s += a[i] + x + y + u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// the real per-item math out of shared memory, no HBM: 16 warps, each `rounds` x 4 items
__global__ void __launch_bounds__(512, 1) k_dot(float* out, int rounds, long long* cycles) {
    extern __shared__ uint8_t smraw[];
    uint8_t* ring = smraw;                       // 32 KiB of "weights"
    float* xs = reinterpret_cast<float*>(smraw + kStageBytes);  // 4096 floats
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kStageBytes / 4; i += 512) reinterpret_cast<uint32_t*>(ring)[i] = i * 2654435761u;
    for (int i = tid; i < 4096; i += 512) xs[i] = 0.001f * i;
    __syncthreads();
    f32x2 xr[16];
    const float* xp = xs + 4 * lane;
    for (int j = 0; j < 8; ++j) { const uint4 q = lds128(xp + 128 * j); xr[2 * j] = pack2u(q.x, q.y); xr[2 * j + 1] = pack2u(q.z, q.w); }
    float acc = 0.f;
    const Q4Consts kc = q4_consts();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const uint8_t* wbase = ring + (size_t)warp * 4 * kItemBytes + lane * 16;
        float v[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) v[g] = dot_q4(lds128(wbase + g * kItemBytes), xr, kc);
        acc += reduce4(v[0], v[1], v[2], v[3], lane);
    }
    const long long t1 = clock64();
    out[blockIdx.x * 512 + tid] = acc;
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// ring alone: producer streams `rounds` stages of 32 KiB per CTA, consumers only wait + arrive
__global__ void __launch_bounds__(kGemvThreads, 1) k_stream(const uint8_t* src, size_t bytes_per_cta, int stages, float* out) {
    extern __shared__ uint8_t smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uintptr_t p = (reinterpret_cast<uintptr_t>(smraw) + 127) & ~uintptr_t(127);
    uint8_t* ring = reinterpret_cast<uint8_t*>(p);
    uint64_t* full = reinterpret_cast<uint64_t*>(p + (size_t)stages * kStageBytes);
    uint64_t* empty = full + 8;
    if (tid == 0) { for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kConsumerWarps); } fence_mbar_init(); }
    __syncthreads();
    const int rounds = (int)(bytes_per_cta / kStageBytes);
    const uint8_t* s = src + (size_t)blockIdx.x * bytes_per_cta;
    if (warp == kConsumerWarps) {
        if (lane == 0)
            for (int r = 0; r < rounds; ++r) {
                const int st = r % stages, use = r / stages;
                if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
                mbar_arrive_expect_tx(&full[st], kStageBytes);
                bulk_g2s_evict_first(ring + (size_t)st * kStageBytes, s + (size_t)r * kStageBytes, kStageBytes, &full[st]);
            }
        return;
    }
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        const int st = r % stages;
        mbar_wait(&full[st], (r / stages) & 1);
        acc += reinterpret_cast<const uint32_t*>(ring + (size_t)st * kStageBytes)[tid];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
    out[blockIdx.x * 512 + tid] = (float)acc;
}

__global__ void __launch_bounds__(512, 1) k_barrier(unsigned int* bar, int n, long long* cycles) {
    unsigned int target = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        bar_sync(1, 512);
        if (threadIdx.x == 0) {
            red_release_add(bar, 1u);
            target += gridDim.x;
            while (ld_acquire_u32(bar) < target) {}
        } else target += gridDim.x;
        bar_sync(1, 512);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int DBG>
static int run_real(const char* label, int K, int N, int sms, int reps) {
    QLayout L = make_layout(K, N, 4, sms);
    int stages = 0; size_t smem = 0;
    for (int s = kMaxStages; s >= 2; --s) if (gemv_smem_bytes(L, s) <= 227 * 1024) { stages = s; smem = gemv_smem_bytes(L, s); break; }
    const size_t bytes = layout_bytes(L);
    const int copies = 8;   // cycle through several matrices so the working set exceeds L2
    uint8_t* w; float *cs, *x, *y;
    CK(cudaMalloc(&w, bytes * copies)); CK(cudaMemset(w, 0x5A, bytes * copies));
    CK(cudaMalloc(&cs, 4 * (size_t)4 * L.U)); CK(cudaMalloc(&x, 4 * (size_t)K)); CK(cudaMalloc(&y, 4 * (size_t)N));
    CK(cudaMemset(cs, 0, 4 * (size_t)4 * L.U)); CK(cudaMemset(x, 0, 4 * (size_t)K));
    CK(cudaFuncSetAttribute(gemv_kernel<4, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GemvArgs a{};
    a.colscale = cs; a.L = L; a.stages = stages; a.x = x; a.epi = EPI_STORE; a.out = y;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 3; ++r) { a.wq = w + (size_t)(r % copies) * bytes; gemv_kernel<4, DBG><<<L.P, kGemvThreads, smem>>>(a); }
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) { a.wq = w + (size_t)(r % copies) * bytes; gemv_kernel<4, DBG><<<L.P, kGemvThreads, smem>>>(a); }
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("real gemv_kernel<4> %-22s K=%5d N=%5d stages %d: %7.2f us/launch, %6.0f GB/s\n", label, K, N, stages, ms * 1e3 / reps, bytes / (ms * 1e-3 / reps) / 1e9);
    cudaFree(w); cudaFree(cs); cudaFree(x); cudaFree(y);
    return 0;
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    float* out; long long* cyc; long long h;
    CK(cudaMalloc(&out, sizeof(float) * sms * 544));
    CK(cudaMalloc(&cyc, 8));
    const int iters = 20000;
    const char* names[] = {"FFMA", "FADD", "FFMA2", "FADD2", "LOP3", "LOP3x2+FADD2+FFMA2"};
#define RUN_OP(OP) { k_ops<OP><<<sms, 512>>>(out, 100, cyc); k_ops<OP><<<sms, 512>>>(out, iters, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost)); \
    double per = (double)h / ((double)iters * 8 * 4); printf("%-22s %7.3f cycles per warp-instr(group) per SMSP (4 warps/SMSP)\n", names[OP], per); }
    RUN_OP(0) RUN_OP(1) RUN_OP(2) RUN_OP(3) RUN_OP(4) RUN_OP(5)
    {
        CK(cudaFuncSetAttribute(k_dot, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        const int rounds = 2000;
        k_dot<<<sms, 512, 48 * 1024 + 4096 * 4>>>(out, 10, cyc);
        k_dot<<<sms, 512, 48 * 1024 + 4096 * 4>>>(out, rounds, cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double items = (double)rounds * 4 * 16;
        printf("dot_q4 from smem: %.1f cycles per item per SM (%.1f per SMSP-item); at 1.9 GHz x %d SMs -> %.0f GB/s of INT4 weights\n",
               h / items, 4 * h / items, sms, 512.0 / (h / items) * 1.9e9 * sms / 1e9);
    }
    {
        CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        const size_t per_cta = (size_t)64 * kStageBytes;  // 2 MiB per CTA
        uint8_t* src;
        CK(cudaMalloc(&src, per_cta * sms * 4));
        CK(cudaMemset(src, 1, per_cta * sms * 4));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int stages : {1, 2, 4, 6}) {
            const size_t smem = (size_t)stages * kStageBytes + 256 + 128;
            k_stream<<<sms, kGemvThreads, smem>>>(src, per_cta, stages, out);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            for (int rep = 0; rep < 4; ++rep) k_stream<<<sms, kGemvThreads, smem>>>(src + (size_t)rep * per_cta * sms, per_cta, stages, out);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("bulk-copy ring, %d stages x 32 KiB: %.0f GB/s (%.1f us per 2 MiB/CTA launch)\n", stages, 4.0 * per_cta * sms / (ms * 1e-3) / 1e9, ms * 1e3 / 4);
        }
    }
    {
        unsigned int* bar;
        CK(cudaMalloc(&bar, 4));
        CK(cudaMemset(bar, 0, 4));
        const int n = 2000;
        void* args[] = {&bar, (void*)&n, &cyc};
        CK(cudaLaunchCooperativeKernel((const void*)k_barrier, dim3(sms), dim3(512), args, 0, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("grid barrier over %d CTAs: %.0f cycles (%.2f us at 1.9 GHz)\n", sms, (double)h / n, (double)h / n / 1900.0);
    }
    for (auto kn : std::vector<std::pair<int,int>>{{4096, 22016}, {4096, 4096}, {11008, 4096}, {2048, 2048}, {8192, 57344}}) {
        run_real<0>("full", kn.first, kn.second, sms, 64);
        run_real<1>("no math (LDS only)", kn.first, kn.second, sms, 64);
        run_real<2>("no LDS, no math", kn.first, kn.second, sms, 64);
    }
    return 0;
}
