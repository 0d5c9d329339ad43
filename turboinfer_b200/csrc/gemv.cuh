// gemv.cuh -- dequant-fused INT4 / INT8 GEMV for batch-1 decode on sm_100a.
//
// Replaces, on the decode path, TensorEngine::matmul (src/core/tensor_engine.cpp:490-640: convert_dtype of the
// whole weight to fp32 followed by the scalar column-strided i-j-k loop) plus the ops the reference runs around
// it as separate passes: rms_norm (:1452-1508) as a prologue, and add / silu*multiply / relu / apply_rope /
// KVCache append (:1626-1743, :900-923, :1510-1624, inference_engine.cpp:78-160) as epilogues.
//
// One CTA per SM, 17 warps:
//   warp 16 (one elected lane)  producer: streams the CTA's contiguous slab of packed weights HBM -> shared memory
//                               with 1-D bulk async copies (TMA engine, cp.async.bulk + mbarrier complete_tx),
//                               <= 32 KiB per stage, kStages deep, weights marked L2 evict-first
//   warps 0..15                 consumers: stage x into shared memory (fused RMSNorm, nibble-position prescale),
//                               then per stage: LDS.128 weights, unpack in registers (LOP3 magic-number trick, no
//                               int->float converts), FADD2 / FFMA2 fp32 accumulate, 6-shuffle transposed warp
//                               reduction of 4 items, partial sums to shared memory
//   epilogue                    per column: fixed-order sum of the k-superchunk partials, * scale (+ zero-point
//                               term), then the fused residual / SwiGLU / ReLU / RoPE+KV-append / logits+argmax
// HBM-bound by design: algorithmic bytes = K*N*bits/8 + O(K + N) floats; every packed byte is read exactly once.
#pragma once
#include "ptx.cuh"
#include "qlayout.cuh"

namespace tib {

constexpr int kGemvThreads = (kConsumerWarps + 1) * 32;  // 544
constexpr int kConsumerThreads = kConsumerWarps * 32;    // 512
constexpr int kMaxStages = 6;

enum GemvEpilogue : int {
    EPI_STORE = 0,    // out[n] = y
    EPI_RESIDUAL = 1, // out[n] = resid[n] + y                         (TensorEngine::add, :1626)
    EPI_SWIGLU = 2,   // columns interleaved (gate_i, up_i): out[i] = up * silu(gate)   (:900-923, :1680)
    EPI_RELU = 3,     // out[n] = max(0, y)                           (:828)
    EPI_QKV = 4,      // [q | k | v]: optional RoPE on q,k; q -> out, k/v -> paged KV cache at position pos
    EPI_LOGITS = 5    // out[n] = y and a running (max, first index) in argmax_key
};

struct GemvArgs {
    // weights
    const uint8_t* wq;
    const float* colscale;  // [4*U] scale per column
    const float* colzterm;  // [4*U] zero-point term per column (y += scale*zterm*sum(x)), or nullptr
    QLayout L;
    int stages;
    // prologue
    const float* x;        // [K]
    const float* norm_w;   // RMSNorm weight [K] or nullptr
    float rms_eps;
    // epilogue
    int epi;
    float* out;
    const float* resid;
    // EPI_QKV
    int hidden;             // H: q = cols [0,H), k = [H,2H), v = [2H,3H)
    int rope_dim;           // 0: no RoPE; else the rotation group size (head_dim or H)
    const float* inv_freq;  // [rope_dim/2], computed on the host with powf like the reference (:1562-1565)
    const int* pos_ptr;     // device scalar: position of this token = tokens already in the cache
    float* k_pool;          // KV pools of this layer, page-major: [page][page_tokens][H]
    float* v_pool;
    const int* page_table;  // logical page -> physical page
    int page_tokens;
    // EPI_LOGITS
    unsigned long long* argmax_key;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the 512 consumer threads, identical (and identically ordered) in every thread
__device__ __forceinline__ float consumer_block_sum(float v, float* red, int warp, int lane) {
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    bar_sync(1, kConsumerThreads);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) s += red[i];
    bar_sync(1, kConsumerThreads);
    return s;
}

// ---- per-item dot products -------------------------------------------------------------------------
// INT4: a lane's 16 B hold 32 nibbles u = q + 8.  (w & (0xF << 4p)) | 0x4B000000 is the float 2^23 + u*16^p;
// adding -(2^23 + 8*16^p) leaves (u - 8)*16^p exactly, and x was pre-multiplied by 16^-p when it was staged,
// so one LOP3 + half an FADD2 + half an FFMA2 per weight, all exact until the fp32 accumulate.
__device__ __forceinline__ float dot_q4(const uint4& wv, const f32x2 (&xr)[16]) {
    const uint32_t MAGIC = 0x4B000000u;
    const f32x2 C01 = pack2(-8388616.f, -8388736.f);
    const f32x2 C23 = pack2(-8390656.f, -8421376.f);
    const f32x2 C42 = pack2(-8912896.f, -8390656.f);
    const f32x2 C34 = pack2(-8421376.f, -8912896.f);
    f32x2 acc0 = 0ull, acc1 = 0ull;
    const uint32_t words[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t w = words[i];
        const uint32_t wh = w >> 12;
        const uint32_t m0 = and_or(w, 0x0000000Fu, MAGIC);
        const uint32_t m1 = and_or(w, 0x000000F0u, MAGIC);
        const uint32_t m2 = and_or(w, 0x00000F00u, MAGIC);
        const uint32_t m3 = and_or(w, 0x0000F000u, MAGIC);
        const uint32_t m4 = and_or(w, 0x000F0000u, MAGIC);
        const uint32_t m5 = and_or(wh, 0x00000F00u, MAGIC);
        const uint32_t m6 = and_or(wh, 0x0000F000u, MAGIC);
        const uint32_t m7 = and_or(wh, 0x000F0000u, MAGIC);
        acc0 = fma2(add2(pack2u(m0, m1), C01), xr[4 * i + 0], acc0);
        acc1 = fma2(add2(pack2u(m2, m3), C23), xr[4 * i + 1], acc1);
        acc0 = fma2(add2(pack2u(m4, m5), C42), xr[4 * i + 2], acc0);
        acc1 = fma2(add2(pack2u(m6, m7), C34), xr[4 * i + 3], acc1);
    }
    float a, b;
    unpack2(add2(acc0, acc1), a, b);
    return a + b;
}

// INT8: a lane's 16 B hold 16 bytes q + 128; PRMT drops byte e into the mantissa of 2^23.
__device__ __forceinline__ float dot_q8(const uint4& wv, const f32x2 (&xr)[8]) {
    const uint32_t MAGIC = 0x4B000000u;
    const f32x2 C = pack2(-8388736.f, -8388736.f);
    f32x2 acc0 = 0ull, acc1 = 0ull;
    const uint32_t words[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t w = words[i];
        const uint32_t m0 = prmt(w, MAGIC, 0x7650u);
        const uint32_t m1 = prmt(w, MAGIC, 0x7651u);
        const uint32_t m2 = prmt(w, MAGIC, 0x7652u);
        const uint32_t m3 = prmt(w, MAGIC, 0x7653u);
        acc0 = fma2(add2(pack2u(m0, m1), C), xr[2 * i + 0], acc0);
        acc1 = fma2(add2(pack2u(m2, m3), C), xr[2 * i + 1], acc1);
    }
    float a, b;
    unpack2(add2(acc0, acc1), a, b);
    return a + b;
}

// Reduce four per-lane values over the warp with 6 shuffles; afterwards lane 8*i (i < 4) holds the full
// sum of v[i].  Every v[i] goes through the same addition tree, so equal inputs give bit-equal sums.
__device__ __forceinline__ float reduce4(float v0, float v1, float v2, float v3, int lane) {
    const bool hi16 = lane & 16;
    float keep0 = hi16 ? v2 : v0, keep1 = hi16 ? v3 : v1;
    float send0 = hi16 ? v0 : v2, send1 = hi16 ? v1 : v3;
    keep0 += __shfl_xor_sync(0xffffffffu, send0, 16);
    keep1 += __shfl_xor_sync(0xffffffffu, send1, 16);
    const bool hi8 = lane & 8;
    float keep = hi8 ? keep1 : keep0;
    float send = hi8 ? keep0 : keep1;
    keep += __shfl_xor_sync(0xffffffffu, send, 8);
    keep += __shfl_xor_sync(0xffffffffu, keep, 4);
    keep += __shfl_xor_sync(0xffffffffu, keep, 2);
    keep += __shfl_xor_sync(0xffffffffu, keep, 1);
    return keep;
}

__device__ __forceinline__ unsigned long long argmax_pack(float v, int idx) {
    // order-preserving map of the float, index stored inverted so that atomicMax keeps the FIRST maximum
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (uint32_t)(0x7FFFFFFF - idx);
}

// ---- shared-memory carve-up ------------------------------------------------------------------------
struct GemvSmem {
    uint8_t* ring;      // stages * 32 KiB
    float* xs;          // kpad floats
    float* part;        // max items floats
    float* red;         // 32 floats
    uint64_t* full;     // [stages]
    uint64_t* empty;    // [stages]
};

TIB_HD size_t gemv_smem_bytes(const QLayout& L, int stages) {
    size_t b = (size_t)stages * kStageBytes;
    b += (size_t)layout_kpad(L) * 4;
    b += (size_t)slab_max_items(L) * 4;
    b += 32 * 4;
    b += (size_t)2 * kMaxStages * 8;
    return b + 128;
}

__device__ __forceinline__ GemvSmem gemv_carve(uint8_t* base, const QLayout& L, int stages) {
    GemvSmem s;
    uintptr_t p = (reinterpret_cast<uintptr_t>(base) + 127) & ~uintptr_t(127);
    s.ring = reinterpret_cast<uint8_t*>(p);
    p += (size_t)stages * kStageBytes;
    s.xs = reinterpret_cast<float*>(p);
    p += (size_t)layout_kpad(L) * 4;
    s.part = reinterpret_cast<float*>(p);
    p += (size_t)slab_max_items(L) * 4;
    s.red = reinterpret_cast<float*>(p);
    p += 32 * 4;
    s.full = reinterpret_cast<uint64_t*>(p);
    s.empty = s.full + kMaxStages;
    return s;
}

// ---- building blocks shared by the stand-alone GEMV kernel and the persistent decode kernel (mega.cuh) -------

// cross-CTA data (written by another SM earlier in the same launch) must not be served from a stale L1 line
__device__ __forceinline__ float ld_act(const float* p, bool coherent) { return coherent ? __ldcg(p) : *p; }

struct PhaseCtx {
    bool coherent;   // activations were produced by other CTAs of the same launch: read them through L2
    int pos;         // cache position of the current token (EPI_QKV); < 0: read *pos_ptr
    unsigned long long* key;  // EPI_LOGITS: argmax key to use instead of GemvArgs::argmax_key (nullptr: keep)
};

// producer: stream this CTA's slab through the ring.  `it` counts stages over the whole launch.
__device__ __forceinline__ void gemv_produce(const GemvArgs& a, const Slab& slab, const GemvSmem& sm, uint32_t& it) {
    const int S = a.stages;
    const uint8_t* src = a.wq + slab.byte0;
    for (int r = 0; r < slab.rounds; ++r, ++it) {
        const uint32_t st = it % S, use = it / S;
        if (use > 0) mbar_wait(&sm.empty[st], (use - 1) & 1);
        const uint32_t bytes = (uint32_t)round_total(slab, r) * kItemBytes;
        mbar_arrive_expect_tx(&sm.full[st], bytes);
        bulk_g2s_evict_first(sm.ring + (size_t)st * kStageBytes, src, bytes, &sm.full[st]);
        src += bytes;
    }
}

// consumers, prologue: stage x into shared memory (fused RMSNorm, INT4 nibble-position prescale).
// Returns sum(x') (only meaningful when want_sum).
template <int BITS>
__device__ __forceinline__ float gemv_stage_x(const GemvArgs& a, const float* x, const GemvSmem& sm, bool coherent, bool want_sum,
                                              int tid, int warp, int lane) {
    const QLayout& L = a.L;
    const int K = L.K, kpad = layout_kpad(L);
    float ss = 0.f;
    for (int k = tid; k < kpad; k += kConsumerThreads) {
        const float v = k < K ? ld_act(x + k, coherent) : 0.f;
        sm.xs[k] = v;
        ss = fmaf(v, v, ss);
    }
    float rms = 1.f;
    if (a.norm_w != nullptr) {
        const float tot = consumer_block_sum(ss, sm.red, warp, lane);
        rms = sqrtf(tot / (float)K + a.rms_eps);  // :1501
    }
    float sx = 0.f;
    for (int k = tid; k < kpad; k += kConsumerThreads) {
        float v = sm.xs[k];
        if (a.norm_w != nullptr && k < K) v = (v / rms) * a.norm_w[k];  // :1504-1506, same two roundings
        sx += v;
        if (BITS == 4) {
            const int p = q4_pos((k & 1023) >> 7, k & 3);
            v *= __uint_as_float((uint32_t)(127 - 4 * p) << 23);  // 16^-p, exact
        }
        sm.xs[k] = v;
    }
    float sumx = 0.f;
    if (want_sum) sumx = consumer_block_sum(sx, sm.red, warp, lane);
    bar_sync(1, kConsumerThreads);
    return sumx;
}

// consumers, main loop: per ring stage LDS.128 weights, unpack, FFMA2, reduce, partials to shared memory
template <int BITS>
__device__ __forceinline__ void gemv_consume(const GemvArgs& a, const Slab& slab, const GemvSmem& sm, uint32_t& it, int warp, int lane) {
    const QLayout& L = a.L;
    const int S = a.stages;
    constexpr int NX = BITS == 4 ? 16 : 8;  // float2 pairs of x per lane per superchunk
    f32x2 xr[NX];
    int cur_s = -1;
    const int my_n = warp_items(slab, warp);
    const int first = warp_first_item(slab, warp);
    int it_s = first / slab.ncols;  // superchunk of the next item
    int it_c = first - it_s * slab.ncols;
    int item = first;
    for (int r = 0; r < slab.rounds; ++r, ++it) {
        const uint32_t st = it % S;
        const int g_n = round_items(my_n, r);
        mbar_wait(&sm.full[st], (it / S) & 1);
        const uint8_t* wbase = sm.ring + (size_t)st * kStageBytes + (size_t)round_warp_offset(slab, r, warp) * kItemBytes + lane * 16;
        float v[kItemsPerRound] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int g = 0; g < kItemsPerRound; ++g) {
            if (g < g_n) {
                if (it_s != cur_s) {
                    cur_s = it_s;
                    const float* xp = sm.xs + (size_t)cur_s * L.ksc + 4 * lane;
#pragma unroll
                    for (int j = 0; j < NX / 2; ++j) {
                        const uint4 q = lds128(xp + 128 * j);
                        xr[2 * j] = pack2u(q.x, q.y);
                        xr[2 * j + 1] = pack2u(q.z, q.w);
                    }
                }
                const uint4 wv = lds128(wbase + g * kItemBytes);
                if constexpr (BITS == 4) v[g] = dot_q4(wv, xr);
                else v[g] = dot_q8(wv, xr);
                if (++it_c == slab.ncols) { it_c = 0; ++it_s; }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);  // this warp is done reading the stage
        const float tot = reduce4(v[0], v[1], v[2], v[3], lane);
        const int gi = lane >> 3;
        if ((lane & 7) == 0 && gi < g_n) sm.part[item + gi] = tot;  // part[s*ncols + c]
        item += g_n;
    }
    bar_sync(1, kConsumerThreads);
}

// consumers, epilogue: fixed-order sum of the superchunk partials, scale, fused tail op
__device__ __forceinline__ void gemv_epilogue(const GemvArgs& a, const Slab& slab, const GemvSmem& sm, float sumx, const float* resid,
                                              const PhaseCtx& ctx, int tid, int lane) {
    const QLayout& L = a.L;
    const int ncols = slab.ncols, nsc = L.nsc;
    auto column = [&](int c) -> float {
        float acc = 0.f;
        for (int s = 0; s < nsc; ++s) acc += sm.part[s * ncols + c];
        const int n = slab.col0 + c;
        if (a.colzterm != nullptr) acc = fmaf(a.colzterm[n], sumx, acc);
        return a.colscale[n] * acc;
    };
    if (a.epi == EPI_SWIGLU || a.epi == EPI_QKV) {
        for (int pc = tid; pc < ncols / 2; pc += kConsumerThreads) {  // column pairs
            const int n0 = slab.col0 + 2 * pc;
            if (n0 >= L.N) continue;
            const float y0 = column(2 * pc), y1 = column(2 * pc + 1);
            if (a.epi == EPI_SWIGLU) {
                const float sg = y0 / (1.0f + expf(-y0));  // silu(gate), :918
                a.out[n0 >> 1] = y1 * sg;                   // multiply(up, silu(gate))
            } else {
                const int H = a.hidden;
                const int seg = n0 / H, d = n0 - seg * H;
                const int pos = ctx.pos >= 0 ? ctx.pos : *a.pos_ptr;
                float o0 = y0, o1 = y1;
                if (seg < 2 && a.rope_dim > 0) {
                    const int i = (d % a.rope_dim) >> 1;
                    float sn, cs;
                    sincosf((float)pos * a.inv_freq[i], &sn, &cs);
                    o0 = __fsub_rn(__fmul_rn(y0, cs), __fmul_rn(y1, sn));  // :1584-1585, un-fused like the build
                    o1 = __fadd_rn(__fmul_rn(y0, sn), __fmul_rn(y1, cs));
                }
                if (seg == 0) {
                    a.out[d] = o0; a.out[d + 1] = o1;
                } else {
                    const int page = a.page_table[pos / a.page_tokens];
                    const size_t off = ((size_t)page * a.page_tokens + (pos % a.page_tokens)) * H + d;
                    float* dst = seg == 1 ? a.k_pool : a.v_pool;
                    dst[off] = o0; dst[off + 1] = o1;
                }
            }
        }
    } else {
        float best = -INFINITY;
        int besti = 0x7FFFFFFF;
        for (int c = tid; c < ncols; c += kConsumerThreads) {
            const int n = slab.col0 + c;
            if (n >= L.N) continue;
            float y = column(c);
            if (a.epi == EPI_RESIDUAL) y = ld_act(resid + n, ctx.coherent) + y;
            else if (a.epi == EPI_RELU) y = fmaxf(y, 0.f);
            a.out[n] = y;
            if (a.epi == EPI_LOGITS && (y > best)) { best = y; besti = n; }
        }
        if (a.epi == EPI_LOGITS) {
            unsigned long long key = besti == 0x7FFFFFFF ? 0ull : argmax_pack(best, besti);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            if (lane == 0 && key != 0ull) atomicMax(ctx.key ? ctx.key : a.argmax_key, key);
        }
    }
}

__device__ __forceinline__ void gemv_init_barriers(const GemvSmem& sm, int stages) {
    for (int i = 0; i < stages; ++i) {
        mbar_init(&sm.full[i], 1);
        mbar_init(&sm.empty[i], kConsumerWarps);
    }
    fence_mbar_init();
}

// ---- the stand-alone kernel: one GEMV per launch -----------------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(kGemvThreads, 1) gemv_kernel(const __grid_constant__ GemvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Slab slab = make_slab(a.L, blockIdx.x);
    const GemvSmem sm = gemv_carve(smem_raw, a.L, a.stages);
    if (tid == 0) gemv_init_barriers(sm, a.stages);
    __syncthreads();
    uint32_t it = 0;
    if (warp == kConsumerWarps) {
        // the weights do not depend on the previous kernel: start streaming at once
        if (lane == 0) gemv_produce(a, slab, sm, it);
        return;
    }
    pdl_wait_prior_grid();  // x (and resid / pos) come from the previous kernel in the stream
    const PhaseCtx ctx{false, -1, nullptr};
    const float sumx = gemv_stage_x<BITS>(a, a.x, sm, false, a.colzterm != nullptr, tid, warp, lane);
    gemv_consume<BITS>(a, slab, sm, it, warp, lane);
    gemv_epilogue(a, slab, sm, sumx, a.resid, ctx, tid, lane);
}

}  // namespace tib
