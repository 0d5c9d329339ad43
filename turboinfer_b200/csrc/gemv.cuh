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
struct Q4Consts {
    uint32_t magic;
    f32x2 c01, c23, c42, c34;
};
// The constants are made opaque to the compiler once per kernel so they live in registers instead of being
// re-materialised (UMOV pairs) for every item.
__device__ __forceinline__ Q4Consts q4_consts() {
    Q4Consts c;
    c.magic = 0x4B000000u;
    c.c01 = pack2(-8388616.f, -8388736.f);
    c.c23 = pack2(-8390656.f, -8421376.f);
    c.c42 = pack2(-8912896.f, -8390656.f);
    c.c34 = pack2(-8421376.f, -8912896.f);
    asm volatile("" : "+r"(c.magic), "+l"(c.c01), "+l"(c.c23), "+l"(c.c42), "+l"(c.c34));
    return c;
}
__device__ __forceinline__ float dot_q4(const uint4& wv, const f32x2 (&xr)[16], const Q4Consts& k) {
    const uint32_t MAGIC = k.magic;
    const f32x2 C01 = k.c01, C23 = k.c23, C42 = k.c42, C34 = k.c34;
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
__device__ __forceinline__ float dot_q8(const uint4& wv, const f32x2 (&xr)[8], const Q4Consts& k) {
    const uint32_t MAGIC = k.magic;
    const f32x2 C = k.c01 == 0ull ? 0ull : pack2(-8388736.f, -8388736.f);
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
// Fast path (K % 4 == 0, 16-byte aligned pointers): 128-bit loads, four per thread in flight before first use, so
// the global-memory latency is paid once per batch instead of once per element.
__device__ __forceinline__ float4 ld_act4(const float* p, bool coherent) {
    return coherent ? __ldcg(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
}
template <int BITS>
__device__ __forceinline__ float4 q4_prescale4(float4 v, int k0) {
    if (BITS == 4) {
        // element e of the float4 at k0 sits at nibble position p(j, e), j = (k0 % 1024) / 128  (qlayout.cuh q4_pos)
        if (((k0 & 1023) >> 7) & 1) { v.x *= 1.0f / 65536.0f; v.y *= 1.0f / 256.0f; v.z *= 1.0f / 4096.0f; v.w *= 1.0f / 65536.0f; }
        else { v.y *= 1.0f / 16.0f; v.z *= 1.0f / 256.0f; v.w *= 1.0f / 4096.0f; }
    }
    return v;
}

template <int BITS>
__device__ __forceinline__ float gemv_stage_x(const GemvArgs& a, const float* x, const GemvSmem& sm, bool coherent, bool want_sum,
                                              int tid, int warp, int lane) {
    const QLayout& L = a.L;
    const int K = L.K, kpad = layout_kpad(L);
    const float* nw = a.norm_w;
    const bool vec = (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(nw)) & 15) == 0;
    float ss = 0.f, sx = 0.f, rms = 1.f;
    if (vec) {
        constexpr int B = 4;  // float4 loads in flight per thread
        const int nvec = kpad >> 2, kvec = K >> 2;
        for (int v0 = tid; v0 < nvec; v0 += B * kConsumerThreads) {
            float4 xv[B];
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const int v = v0 + i * kConsumerThreads;
                xv[i] = v < kvec ? ld_act4(x + 4 * v, coherent) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const int v = v0 + i * kConsumerThreads;
                if (v < nvec) {
                    ss = fmaf(xv[i].x, xv[i].x, ss); ss = fmaf(xv[i].y, xv[i].y, ss);
                    ss = fmaf(xv[i].z, xv[i].z, ss); ss = fmaf(xv[i].w, xv[i].w, ss);
                    if (nw == nullptr) {  // no second pass needed: finish now
                        sx += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
                        xv[i] = q4_prescale4<BITS>(xv[i], 4 * v);
                    }
                    *reinterpret_cast<float4*>(sm.xs + 4 * v) = xv[i];
                }
            }
        }
        if (nw != nullptr) {
            // the norm weights do not depend on the reduction: get the first batch moving before the barriers
            float4 wv[B];
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const int v = tid + i * kConsumerThreads;
                wv[i] = v < kvec ? *reinterpret_cast<const float4*>(nw + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float tot = consumer_block_sum(ss, sm.red, warp, lane);
            rms = sqrtf(tot / (float)K + a.rms_eps);  // :1501
            for (int v0 = tid; v0 < kvec; v0 += B * kConsumerThreads) {
                if (v0 != tid) {
#pragma unroll
                    for (int i = 0; i < B; ++i) {
                        const int v = v0 + i * kConsumerThreads;
                        wv[i] = v < kvec ? *reinterpret_cast<const float4*>(nw + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int i = 0; i < B; ++i) {
                    const int v = v0 + i * kConsumerThreads;
                    if (v < kvec) {
                        float4 t = *reinterpret_cast<const float4*>(sm.xs + 4 * v);
                        t.x = (t.x / rms) * wv[i].x;  // :1504-1506, same two roundings
                        t.y = (t.y / rms) * wv[i].y;
                        t.z = (t.z / rms) * wv[i].z;
                        t.w = (t.w / rms) * wv[i].w;
                        sx += (t.x + t.y) + (t.z + t.w);
                        *reinterpret_cast<float4*>(sm.xs + 4 * v) = q4_prescale4<BITS>(t, 4 * v);
                    }
                }
            }
        }
    } else {
        for (int k = tid; k < kpad; k += kConsumerThreads) {
            const float v = k < K ? ld_act(x + k, coherent) : 0.f;
            sm.xs[k] = v;
            ss = fmaf(v, v, ss);
        }
        if (nw != nullptr) {
            const float tot = consumer_block_sum(ss, sm.red, warp, lane);
            rms = sqrtf(tot / (float)K + a.rms_eps);
        }
        for (int k = tid; k < kpad; k += kConsumerThreads) {
            float v = sm.xs[k];
            if (nw != nullptr && k < K) v = (v / rms) * nw[k];
            sx += v;
            if (BITS == 4) {
                const int p = q4_pos((k & 1023) >> 7, k & 3);
                v *= __uint_as_float((uint32_t)(127 - 4 * p) << 23);  // 16^-p, exact
            }
            sm.xs[k] = v;
        }
    }
    float sumx = 0.f;
    if (want_sum) sumx = consumer_block_sum(sx, sm.red, warp, lane);
    bar_sync(1, kConsumerThreads);
    return sumx;
}

// consumers, main loop: per ring stage LDS.128 weights, unpack, FFMA2, reduce, partials to shared memory.
// A round hands each warp up to 4 items.  The common case -- 4 items of the same k-superchunk -- is straight-line
// code, so the four unpack / FMA chains interleave (ILP 4); rounds that end the warp's range or straddle a
// superchunk boundary take the per-item path.  The warp reduction of round r is issued after round r+1's
// shared-memory loads, hiding the shuffle latency behind them.
template <int BITS, int DBG = 0>
__device__ __forceinline__ void gemv_consume(const GemvArgs& a, const Slab& slab, const GemvSmem& sm, uint32_t& it, int warp, int lane) {
    const QLayout& L = a.L;
    const int S = a.stages;
    const int ksc = L.ksc, ncols = slab.ncols, nrounds = slab.rounds;
    constexpr int NX = BITS == 4 ? 16 : 8;  // float2 pairs of x per lane per superchunk
    const Q4Consts kc = q4_consts();
    f32x2 xr[NX];
    int cur_s = -1;
    const int my_n = warp_items(slab, warp);
    const int first = warp_first_item(slab, warp);
    int it_s = first / ncols;  // superchunk / column of the next item
    int it_c = first - it_s * ncols;
    int item = first;
    const int w_lo = warp < slab.m ? warp : slab.m, w_hi = warp - w_lo;  // warps before this one with b+1 / b items
    const float* xlane = sm.xs + 4 * lane;
    auto load_x = [&](int s) {
        cur_s = s;
        const float* xp = xlane + (size_t)s * ksc;
#pragma unroll
        for (int j = 0; j < NX / 2; ++j) {
            const uint4 q = lds128(xp + 128 * j);
            xr[2 * j] = pack2u(q.x, q.y);
            xr[2 * j + 1] = pack2u(q.z, q.w);
        }
    };
    auto dot = [&](const uint4& wv) -> float {
        if constexpr (DBG == 1) return __uint_as_float(wv.x ^ wv.y ^ wv.z ^ wv.w);
        else if constexpr (BITS == 4) return dot_q4(wv, xr, kc);
        else return dot_q8(wv, xr, kc);
    };
    float pv0 = 0.f, pv1 = 0.f, pv2 = 0.f, pv3 = 0.f;  // previous round's per-lane sums, reduced one round late
    int p_item = 0, p_n = 0;
    auto flush_prev = [&]() {
        if (p_n > 0) {
            const float tot = reduce4(pv0, pv1, pv2, pv3, lane);
            const int gi = lane >> 3;
            if ((lane & 7) == 0 && gi < p_n) sm.part[p_item + gi] = tot;  // part[s*ncols + c]
        }
    };
    for (int r = 0; r < nrounds; ++r, ++it) {
        const uint32_t st = it % S;
        const int g_n = round_items(my_n, r);
        const int woff = w_lo * round_items(slab.b + 1, r) + w_hi * round_items(slab.b, r);
        mbar_wait(&sm.full[st], (it / S) & 1);
        const uint8_t* wbase = sm.ring + (size_t)st * kStageBytes + (size_t)woff * kItemBytes + lane * 16;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        if (DBG == 2) {
            flush_prev();
        } else if (g_n == kItemsPerRound && it_c + kItemsPerRound <= ncols) {
            const uint4 w0 = lds128(wbase), w1 = lds128(wbase + kItemBytes), w2 = lds128(wbase + 2 * kItemBytes),
                        w3 = lds128(wbase + 3 * kItemBytes);
            if (it_s != cur_s) load_x(it_s);
            flush_prev();
            v0 = dot(w0); v1 = dot(w1); v2 = dot(w2); v3 = dot(w3);
            it_c += kItemsPerRound;
            if (it_c == ncols) { it_c = 0; ++it_s; }
        } else {
            flush_prev();
            float v[kItemsPerRound] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int g = 0; g < kItemsPerRound; ++g) {
                if (g < g_n) {
                    if (it_s != cur_s) load_x(it_s);
                    v[g] = dot(lds128(wbase + g * kItemBytes));
                    if (++it_c == ncols) { it_c = 0; ++it_s; }
                }
            }
            v0 = v[0]; v1 = v[1]; v2 = v[2]; v3 = v[3];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);  // this warp is done reading the stage
        pv0 = v0; pv1 = v1; pv2 = v2; pv3 = v3;
        p_item = item; p_n = g_n;
        item += g_n;
    }
    flush_prev();
    bar_sync(1, kConsumerThreads);
}

// Values the epilogue needs that do not depend on this phase's arithmetic; loaded early (before the grid barrier
// in the persistent kernel) so their latency is off the critical path.  They cover the thread's first column
// (or column pair); further columns, if a slab has more than 512, load theirs in place.
struct EpiPre {
    float cs0, cs1;   // colscale
    float zt0, zt1;   // colzterm (0 when absent)
    float r0;         // residual input (EPI_RESIDUAL)
    float invf;       // inv_freq entry (EPI_QKV with RoPE)
    int page;         // physical KV page of the current position (EPI_QKV)
};

__device__ __forceinline__ EpiPre gemv_epilogue_prefetch(const GemvArgs& a, const Slab& slab, const float* resid, const PhaseCtx& ctx, int tid) {
    EpiPre p{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0};
    const bool pairs = a.epi == EPI_SWIGLU || a.epi == EPI_QKV;
    const int c = pairs ? 2 * tid : tid;
    const int n = slab.col0 + c;
    if (c < slab.ncols && n < a.L.N) {
        p.cs0 = a.colscale[n];
        if (a.colzterm) p.zt0 = a.colzterm[n];
        if (pairs) {
            p.cs1 = a.colscale[n + 1];
            if (a.colzterm) p.zt1 = a.colzterm[n + 1];
        }
        if (a.epi == EPI_RESIDUAL) p.r0 = ld_act(resid + n, ctx.coherent);
        if (a.epi == EPI_QKV) {
            const int pos = ctx.pos >= 0 ? ctx.pos : *a.pos_ptr;
            p.page = a.page_table[pos / a.page_tokens];
            if (a.rope_dim > 0) p.invf = a.inv_freq[((n % a.hidden) % a.rope_dim) >> 1];
        }
    }
    return p;
}

// consumers, epilogue: fixed-order sum of the superchunk partials, scale, fused tail op
__device__ __forceinline__ void gemv_epilogue(const GemvArgs& a, const Slab& slab, const GemvSmem& sm, float sumx, const float* resid,
                                              const PhaseCtx& ctx, const EpiPre& pre, int tid, int lane) {
    const QLayout& L = a.L;
    const int ncols = slab.ncols, nsc = L.nsc;
    auto colsum = [&](int c) -> float {
        float acc = 0.f;
        for (int s = 0; s < nsc; ++s) acc += sm.part[s * ncols + c];
        return acc;
    };
    if (a.epi == EPI_SWIGLU || a.epi == EPI_QKV) {
        for (int pc = tid; pc < ncols / 2; pc += kConsumerThreads) {  // column pairs
            const int n0 = slab.col0 + 2 * pc;
            if (n0 >= L.N) continue;
            const bool first = pc == tid;
            const float cs0 = first ? pre.cs0 : a.colscale[n0], cs1 = first ? pre.cs1 : a.colscale[n0 + 1];
            const float zt0 = first ? pre.zt0 : (a.colzterm ? a.colzterm[n0] : 0.f);
            const float zt1 = first ? pre.zt1 : (a.colzterm ? a.colzterm[n0 + 1] : 0.f);
            const float y0 = cs0 * fmaf(zt0, sumx, colsum(2 * pc)), y1 = cs1 * fmaf(zt1, sumx, colsum(2 * pc + 1));
            if (a.epi == EPI_SWIGLU) {
                const float sg = y0 / (1.0f + expf(-y0));  // silu(gate), :918
                a.out[n0 >> 1] = y1 * sg;                   // multiply(up, silu(gate))
            } else {
                const int H = a.hidden;
                const int seg = n0 / H, d = n0 - seg * H;
                const int pos = ctx.pos >= 0 ? ctx.pos : *a.pos_ptr;
                float o0 = y0, o1 = y1;
                if (seg < 2 && a.rope_dim > 0) {
                    const float invf = first ? pre.invf : a.inv_freq[(d % a.rope_dim) >> 1];
                    float sn, cs;
                    sincosf((float)pos * invf, &sn, &cs);
                    o0 = __fsub_rn(__fmul_rn(y0, cs), __fmul_rn(y1, sn));  // :1584-1585, un-fused like the build
                    o1 = __fadd_rn(__fmul_rn(y0, sn), __fmul_rn(y1, cs));
                }
                if (seg == 0) {
                    *reinterpret_cast<float2*>(a.out + d) = make_float2(o0, o1);
                } else {
                    const int page = first ? pre.page : a.page_table[pos / a.page_tokens];
                    const size_t off = ((size_t)page * a.page_tokens + (pos % a.page_tokens)) * H + d;
                    float* dst = seg == 1 ? a.k_pool : a.v_pool;
                    *reinterpret_cast<float2*>(dst + off) = make_float2(o0, o1);
                }
            }
        }
    } else {
        float best = -INFINITY;
        int besti = 0x7FFFFFFF;
        for (int c = tid; c < ncols; c += kConsumerThreads) {
            const int n = slab.col0 + c;
            if (n >= L.N) continue;
            const bool first = c == tid;
            const float cs = first ? pre.cs0 : a.colscale[n];
            const float zt = first ? pre.zt0 : (a.colzterm ? a.colzterm[n] : 0.f);
            float y = cs * fmaf(zt, sumx, colsum(c));
            if (a.epi == EPI_RESIDUAL) y = (first ? pre.r0 : ld_act(resid + n, ctx.coherent)) + y;
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
template <int BITS, int DBG = 0>
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
    const EpiPre pre = gemv_epilogue_prefetch(a, slab, a.resid, ctx, tid);
    gemv_consume<BITS, DBG>(a, slab, sm, it, warp, lane);
    gemv_epilogue(a, slab, sm, sumx, a.resid, ctx, pre, tid, lane);
}

}  // namespace tib
