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
//   warps 0..15                 consumers
//     prologue   x (fp32, optionally RMS-normalised with the reference's roundings) is converted ONCE per GEMV to
//                24-bit block fixed point, x_k ~= s_x * xf_k with |xf_k| < 2^23, and stored in shared memory as three
//                signed 8-bit digit planes (xf = d2*65536 + d1*256 + d0, every digit in [-128, 127])
//     main loop  warp-level integer MMAs (mma.sync.m16n8k32 u8/s8 x s8 -> s32, SASS IMMA.16832): the A operand is 16
//                weight columns x 32 k, stored in HBM in fragment order so that one LDS.128 per lane is the fragment (INT4:
//                of two MMAs -- `w & 0x0F0F0F0F` is the first, `w & 0xF0F0F0F0` the second, 16x too large, which a
//                shift of its own accumulator undoes exactly); the B operand's 8 columns carry the three digit planes
//                (5 columns idle: the tensor pipe has >10x headroom here).  Integer accumulation is exact and order-
//                independent, so the result does not depend on how k is split over warps or (tensor-parallel) GPUs.
//                B200 measurement (scripts/microbench5.cu): IMMA.16832 issues every ~2-3 cycles per SM; the loop runs at
//                9 cycles per 512-byte item per SM out of shared memory (bound by the LDS wavefronts) = 2.4x the HBM
//                rate.  The IDP4A loop it replaces needed 19-27 (scripts/microbench2/3.cu), the fp32 unpack before 18+.
//     reduction  none across lanes: the C fragment of lane (g, t) already holds the sums of columns g, g+8 for digit
//                planes 2t, 2t+1; one shared-memory integer atomic per (column, digit) when a warp leaves a group
//   epilogue                    per column: y = colscale * s_x * (A2*65536 + A1*256 + A0 - off*sum(xf) [+ zero-point
//                               term]); the integer part is exact (int64), then three fp32 roundings; fused residual / SwiGLU /
//                               ReLU / RoPE+KV-append / logits+argmax
// Accuracy: the only approximation is the 24-bit quantisation of x relative to max|x| (absolute error <= 2^-24 max|x|
// per element, i.e. what an fp32 mantissa holds for the largest elements); products and sums are exact.  Measured
// error against the fp64 dot product is below that of the reference's own sequential fp32 accumulation.
// HBM-bound by design: algorithmic bytes = K*N*bits/8 + O(K + N) floats; every packed byte is read exactly once.
#pragma once
#include "ptx.cuh"
#include "qlayout.cuh"

namespace tib {

constexpr int kGemvThreads = (kConsumerWarps + 1) * 32;  // 544
constexpr int kConsumerThreads = kConsumerWarps * 32;    // 512
constexpr int kMaxStages = 6;
constexpr float kXQMax = 8355000.0f;   // bound of |xf|: (1 + 2e-5) * 8355000 + 0.5 < 127*65536 + 127*256 + 127, the largest value three signed digits hold
constexpr int kXCache = 8;             // float4 vectors of x a thread keeps in registers between the two prologue passes

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
    int woff;               // offset added to the stored INT4 values (8 symmetric, 0 asymmetric); INT8 is stored signed
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
    int col_off;            // added to the column index in the arg-max key (tensor parallel: this rank's slice of the vocabulary)
    // persistent kernel: the RMSNorm weight the NEXT GEMV applies to this one's output (nullptr: none); used to hand
    // the next prologue its sum of squares and max|x * w| bound (see XStats)
    const float* next_norm_w;
};

// What a GEMV prologue needs to know about its input before it can convert it to fixed point in ONE pass: the sum of
// squares (RMSNorm) and a bound of max|x_k * w_k| (w = 1 without norm).  In the persistent kernel the phase that
// PRODUCES x leaves one partial per CTA in global memory; the consumers combine them in CTA order (deterministic).
struct XStats {
    float ss, am;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- warp-level integer MMA ----------------------------------------------------------------------------
// D[16 x 8] += A[16 x 32] * B[32 x 8], A = weights (unsigned nibbles-in-bytes for INT4, signed bytes for INT8), B = digits
__device__ __forceinline__ void imma_u8s8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma_s8s8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Accumulators of one warp for the group it is working on.  INT4: lo collects the low-nibble MMAs, hi the high-nibble ones
// (operands 16x too large: total = lo + (hi >> 4), exact) -- two independent dependency chains per warp; INT8: lo only.
// (A second set per parity was dropped: four warps per scheduler hide the MMA latency, and the eight registers it cost
// pushed loop-carried values of the persistent kernel into local memory.)
#ifndef TIB_ACC_SETS
#define TIB_ACC_SETS 1
#endif
#ifndef TIB_ZERO_IN_EPI
#define TIB_ZERO_IN_EPI 1
#endif
#ifndef TIB_LOOP_SYNC
#define TIB_LOOP_SYNC 0   // 0: every warp waits / probes / arrives on the ring's mbarriers itself; 1: no probe of the next stage;
                          // 2: the 16 consumer warps meet at named barriers and ONE thread talks to the mbarriers
#endif
#ifndef TIB_MASK_FREE_HI
#define TIB_MASK_FREE_HI 1
#endif
template <int BITS>
struct QuadAcc {
    int lo[TIB_ACC_SETS][4];
    int hi[TIB_ACC_SETS][4];
    int lb[TIB_MASK_FREE_HI ? TIB_ACC_SETS : 1][4];   // INT4, mask-free high nibbles: low nibbles x the digits of the SECOND 32 k (see kitem_mma)
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int s = 0; s < TIB_ACC_SETS; ++s)
#pragma unroll
            for (int i = 0; i < 4; ++i) { lo[s][i] = 0; hi[s][i] = 0; lb[TIB_MASK_FREE_HI ? s : 0][i] = 0; }
    }
    __device__ __forceinline__ int total(int i) const {
        int l = 0, h = 0;
#pragma unroll
        for (int s = 0; s < TIB_ACC_SETS; ++s) { l += lo[s][i]; h += hi[s][i] - (TIB_MASK_FREE_HI ? lb[s][i] : 0); }
        if constexpr (BITS == 4) return l + (h >> 4);
        else return l;
    }
};

// One k-item: the lane's A fragment registers (w.x .. w.w = registers 0..3) against the digits of the item's k range
// (INT4: xv = {b0, b1 of the first 32 k, b0, b1 of the next 32 k}; INT8: xv.x, xv.y).
// INT4: a byte b = lo + 16 hi holds the nibbles of k and k + 32.  The low nibbles need their mask (4 LOP3); the high nibbles
// are taken WITHOUT masking: b . x2 = lo . x2 + 16 hi . x2, and lo . x2 is one more MMA on the already masked registers, kept
// in its own accumulator and subtracted when the group's sums are flushed -- exact, since everything is an integer.  One
// extra IMMA (the tensor pipe is ~20 % busy in this loop) replaces four LOP3 on the ALU pipe.
template <int BITS>
__device__ __forceinline__ void kitem_mma(QuadAcc<BITS>& acc, int par, const uint4& w, const uint4& xv) {
    if constexpr (BITS == 4) {
        const uint32_t l0 = w.x & 0x0F0F0F0Fu, l1 = w.y & 0x0F0F0F0Fu, l2 = w.z & 0x0F0F0F0Fu, l3 = w.w & 0x0F0F0F0Fu;
        imma_u8s8(acc.lo[par], l0, l1, l2, l3, xv.x, xv.y);
#if TIB_MASK_FREE_HI
        imma_u8s8(acc.hi[par], w.x, w.y, w.z, w.w, xv.z, xv.w);
        imma_u8s8(acc.lb[par], l0, l1, l2, l3, xv.z, xv.w);
#else
        imma_u8s8(acc.hi[par], w.x & 0xF0F0F0F0u, w.y & 0xF0F0F0F0u, w.z & 0xF0F0F0F0u, w.w & 0xF0F0F0F0u, xv.z, xv.w);
#endif
    } else {
        imma_s8s8(acc.lo[par], w.x, w.y, w.z, w.w, xv.x, xv.y);
    }
}
// B fragment of one k-item.  Only B columns 0..2 carry digit planes (lanes g = lane >> 2 < 3); the other lanes read plane 0
// again (their product columns are never used).  Predicating those lanes off would halve the wavefronts of this load, but
// costs four register clears per load pair in a loop that is bound by issue slots, not by shared-memory bandwidth: measured
// slower (profiles/r02_*), so every lane loads.
template <int BITS>
__device__ __forceinline__ uint4 load_xfrag(uint32_t addr, bool on) {
    (void)on;
    if constexpr (BITS == 4) return lds128s(addr);
    else { const uint2 v = lds64s(addr); return make_uint4(v.x, v.y, 0u, 0u); }
}

__device__ __forceinline__ unsigned long long argmax_pack(float v, int idx) {
    // order-preserving map of the float, index stored inverted so that atomicMax keeps the FIRST maximum
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (uint32_t)(0x7FFFFFFF - idx);
}

// ---- shared-memory carve-up ------------------------------------------------------------------------
// The block starts at the kernel's dynamic shared memory (declared 128-byte aligned) and is carved with plain pointer
// arithmetic on that array: the compiler then knows every pointer is a shared-memory one and emits LDS / STS / ATOMS and
// constant shared-window addresses.  (Carving through uintptr_t, as the first version did, loses the address space: every
// access became a generic LD / ST / ATOM, several times slower, and every barrier address a ~10-instruction conversion.)
// The mbarriers and the ring come first so that their addresses are compile-time offsets from the base.
constexpr int kBarBlockBytes = 128;   // full[kMaxStages], empty[kMaxStages], padding
struct GemvSmem {
    uint64_t* full;     // [stages]
    uint64_t* empty;    // [stages]
    uint8_t* ring;      // stages * 32 KiB
    uint8_t* xd;        // digit planes of x: 3 * kpad bytes
    int* acc;           // [ncols][3] integer column sums
    float* red;         // 128 floats of reduction scratch
    long long* sxf;     // [16] per-warp partial sums of xf over k
    unsigned int* spare;         // 128 spare bytes
};

TIB_HD size_t gemv_smem_bytes_for(int stages, int kpad, int max_units) {
    size_t b = kBarBlockBytes;
    b += (size_t)stages * kStageBytes;
    b += (size_t)3 * kpad;
    b += (size_t)max_units * 4 * 3 * 4;
    b += 128 * 4 + 32 * 8;
    return b + 128;
}
TIB_HD size_t gemv_smem_bytes(const QLayout& L, int stages) { return gemv_smem_bytes_for(stages, layout_kpad(L), slab_max_units(L)); }

// `base` must be the (128-byte aligned) dynamic shared memory array itself
__device__ __forceinline__ GemvSmem gemv_carve_for(uint8_t* base, int stages, int kpad, int max_units, uint8_t** end = nullptr) {
    GemvSmem s;
    s.full = reinterpret_cast<uint64_t*>(base);
    s.empty = s.full + kMaxStages;
    s.ring = base + kBarBlockBytes;
    s.xd = s.ring + (size_t)stages * kStageBytes;
    s.acc = reinterpret_cast<int*>(s.xd + (size_t)3 * kpad);   // kpad is a multiple of 128: stays 16-byte aligned
    s.red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s.acc) + (size_t)max_units * 4 * 3 * 4);
    s.sxf = reinterpret_cast<long long*>(s.red + 128);
    s.spare = reinterpret_cast<unsigned int*>(s.sxf + 16);
    if (end) *end = reinterpret_cast<uint8_t*>(s.spare + 32);
    return s;
}
__device__ __forceinline__ GemvSmem gemv_carve(uint8_t* base, const QLayout& L, int stages) {
    return gemv_carve_for(base, stages, layout_kpad(L), slab_max_units(L));
}

// ---- building blocks shared by the stand-alone GEMV kernel and the persistent decode kernel (mega.cuh) -------

// cross-CTA data (written by another SM earlier in the same launch) must not be served from a stale L1 line
__device__ __forceinline__ float ld_act(const float* p, bool coherent) { return coherent ? __ldcg(p) : *p; }

struct PhaseCtx {
    bool coherent;   // activations were produced by other CTAs of the same launch: read them through L2
    int pos;         // cache position of the current token (EPI_QKV); < 0: read *pos_ptr
    unsigned long long* key;  // EPI_LOGITS: argmax key to use instead of GemvArgs::argmax_key (nullptr: keep)
    // tensor parallel, row-parallel GEMV (EPI_STORE): the partial output goes to every rank's partial buffer instead of
    // a.out -- peers[r][peer_off + n], r < npeers (0: single GPU)
    float* const* peers;
    int npeers;
    size_t peer_off;
    // != 0: the partial buffers hold 8-byte words {value, ll_seq} (one single-copy-atomic store per element: the reader polls the
    // tag inside the word it needs, no fence and no flag; the element index is the same, the stride 8 bytes)
    unsigned int ll_seq;
};

// Position in the ring, carried across the GEMVs of a launch (no divisions per phase): stage and the parity of its use.
struct RingPos {
    uint32_t st = 0, par = 0;
    __device__ __forceinline__ void advance(uint32_t S) { if (++st == S) { st = 0; par ^= 1u; } }
};

// producer (the whole warp): stream this CTA's slab through the ring.  `it` counts stages over the whole launch.
// Lane w < 16 computes how many items consumer warp w takes from the stage, one REDUX sums them, lane 0 drives the
// mbarriers and the bulk copy -- the per-stage bookkeeping must stay far below the ~0.7 us a stage lasts at HBM rate.
__device__ __forceinline__ void gemv_produce(const GemvArgs& a, const Slab& slab, const GemvSmem& sm, RingPos& it, int lane) {
    const int S = a.stages;
    const uint8_t* src = a.wq + slab.byte0;
    const int l_nq = warp_quads(slab, lane & 15), l_fq = warp_first_quad(slab, lane & 15);
    for (int r = 0; r < slab.rounds; ++r, it.advance(S)) {
        const int l_items = (lane < kConsumerWarps && r < l_nq) ? (l_fq + r >= slab.qfull ? slab.nlast : 4) : 0;
        const uint32_t bytes = (uint32_t)__reduce_add_sync(0xffffffffu, l_items) * kItemBytes;
        if (lane == 0) {
            const uint32_t st = it.st;
            mbar_wait(&sm.empty[st], it.par ^ 1u);   // a fresh mbarrier counts its "previous" phase as complete
            mbar_arrive_expect_tx(&sm.full[st], bytes);
            bulk_g2s_evict_first(sm.ring + (size_t)st * kStageBytes, src, bytes, &sm.full[st]);
        }
        src += bytes;
    }
    __syncwarp();
}

// ---- prologue: x -> (optional RMSNorm) -> 24-bit block fixed point -> digit planes in shared memory -----------
__device__ __forceinline__ float4 ld_x4(const float* x, int v, int K, bool vec, bool coherent) {
    const int k = 4 * v;
    if (vec) {
        if (k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        return coherent ? __ldcg(reinterpret_cast<const float4*>(x + k)) : *reinterpret_cast<const float4*>(x + k);
    }
    float4 r;
    r.x = k + 0 < K ? ld_act(x + k + 0, coherent) : 0.f;
    r.y = k + 1 < K ? ld_act(x + k + 1, coherent) : 0.f;
    r.z = k + 2 < K ? ld_act(x + k + 2, coherent) : 0.f;
    r.w = k + 3 < K ? ld_act(x + k + 3, coherent) : 0.f;
    return r;
}

// two block-wide reductions at once (sum of a, max of b) over the 512 consumer threads; every thread gets the same,
// identically ordered result.  Uses red[slot*32 .. slot*32+31]; callers alternate slots instead of a trailing barrier.
__device__ __forceinline__ void consumer_sum_max(float& a, float& b, float* red, int warp, int lane) {
    a = warp_sum(a);
    b = warp_max(b);
    if (lane == 0) { red[warp] = a; red[16 + warp] = b; }
    bar_sync(1, kConsumerThreads);
    float s = 0.f, m = 0.f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) { s += red[i]; m = fmaxf(m, red[16 + i]); }
    a = s;
    b = m;
}

template <int BITS>
__device__ __forceinline__ void x_store_digits(uint8_t* xd, int v, float4 y, float inv_s, long long& sxf) {
    const int f0 = __float2int_rn(y.x * inv_s), f1 = __float2int_rn(y.y * inv_s), f2 = __float2int_rn(y.z * inv_s),
              f3 = __float2int_rn(y.w * inv_s);
    sxf += (long long)((f0 + f1) + (f2 + f3));
    // signed digits: the bytes of u = f + 0x808080 are d + 128 (no carries to track), so digit = byte ^ 0x80
    const int u0 = f0 + 0x808080, u1 = f1 + 0x808080, u2 = f2 + 0x808080, u3 = f3 + 0x808080;
    // byte b of each of the four values -> one word per digit
    const uint32_t lo01 = __byte_perm(u0, u1, 0x5140), lo23 = __byte_perm(u2, u3, 0x5140);  // [u0.b0 u1.b0 u0.b1 u1.b1]
    const uint32_t d0 = __byte_perm(lo01, lo23, 0x5410) ^ 0x80808080u;
    const uint32_t d1 = __byte_perm(lo01, lo23, 0x7632) ^ 0x80808080u;
    const uint32_t hi01 = __byte_perm(u0, u1, 0x0062), hi23 = __byte_perm(u2, u3, 0x0062);  // [u0.b2 u1.b2 . .]
    const uint32_t d2 = __byte_perm(hi01, hi23, 0x5410) ^ 0x80808080u;
    const int k = 4 * v;
    *reinterpret_cast<uint32_t*>(xd + xdigit_word_offset(BITS, k, 0)) = d0;
    *reinterpret_cast<uint32_t*>(xd + xdigit_word_offset(BITS, k, 1)) = d1;
    *reinterpret_cast<uint32_t*>(xd + xdigit_word_offset(BITS, k, 2)) = d2;
}

// Returns s_x (x_k ~= s_x * xf_k).  Also zeroes the column accumulators of this phase and leaves sum(xf) in *sm.sxf.
template <int BITS>
__device__ __forceinline__ float gemv_stage_x(const GemvArgs& a, const float* x, const GemvSmem& sm, const Slab& slab, bool coherent,
                                              int tid, int warp, int lane) {
    const QLayout& L = a.L;
    const int K = L.K, kpad = layout_kpad(L);
    const float* nw = a.norm_w;
    const bool vec = (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(nw)) & 15) == 0;
    const int nvec = kpad >> 2;
    const bool cached = nvec <= kXCache * kConsumerThreads;
    for (int i = tid; i < slab.ncols * 3; i += kConsumerThreads) sm.acc[i] = 0;
    float ss = 0.f, amax = 0.f;
    float4 xv[kXCache];
    // pass 1: sum of squares (RMSNorm) and a tight bound of max|y|, y = (x / rms) * w
    if (cached) {
#pragma unroll
        for (int i = 0; i < kXCache; ++i) {
            const int v = tid + i * kConsumerThreads;
            xv[i] = v < nvec ? ld_x4(x, v, K, vec, coherent) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < kXCache; ++i) {
            const int v = tid + i * kConsumerThreads;
            if (v < nvec) {
                const float4 t = xv[i];
                ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
                float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
                if (nw != nullptr) w = ld_x4(nw, v, K, vec, false);
                amax = fmaxf(fmaxf(amax, fabsf(t.x * w.x)), fmaxf(fabsf(t.y * w.y), fmaxf(fabsf(t.z * w.z), fabsf(t.w * w.w))));
            }
        }
    } else {
        for (int v = tid; v < nvec; v += kConsumerThreads) {
            const float4 t = ld_x4(x, v, K, vec, coherent);
            ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
            float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
            if (nw != nullptr) w = ld_x4(nw, v, K, vec, false);
            amax = fmaxf(fmaxf(amax, fabsf(t.x * w.x)), fmaxf(fabsf(t.y * w.y), fmaxf(fabsf(t.z * w.z), fabsf(t.w * w.w))));
        }
    }
    consumer_sum_max(ss, amax, sm.red, warp, lane);
    float rms = 1.f;
    if (nw != nullptr) {
        rms = sqrtf(ss / (float)K + a.rms_eps);  // :1501
        amax = (amax / rms) * 1.000001f;          // covers the two roundings of (x / rms) * w
    }
    const bool finite = amax > 0.f && amax < INFINITY;
    const float inv_s = finite ? kXQMax / amax : 0.f;
    const float s_x = finite ? amax / kXQMax : 0.f;
    // pass 2: y, fixed point, digit planes
    long long sxf = 0;
    auto emit = [&](int v, float4 t) {
        if (nw != nullptr) {
            const float4 w = ld_x4(nw, v, K, vec, false);
            t.x = (t.x / rms) * w.x;  // :1504-1506, same two roundings
            t.y = (t.y / rms) * w.y;
            t.z = (t.z / rms) * w.z;
            t.w = (t.w / rms) * w.w;
        }
        x_store_digits<BITS>(sm.xd, v, t, inv_s, sxf);
    };
    if (cached) {
#pragma unroll
        for (int i = 0; i < kXCache; ++i) {
            const int v = tid + i * kConsumerThreads;
            if (v < nvec) emit(v, xv[i]);
        }
    } else {
        for (int v = tid; v < nvec; v += kConsumerThreads) emit(v, ld_x4(x, v, K, vec, coherent));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sxf += __shfl_xor_sync(0xffffffffu, sxf, o);
    if (lane == 0) sm.sxf[tid >> 5] = sxf;
    bar_sync(1, kConsumerThreads);
    return s_x;
}

// Single-pass prologue of the persistent kernel: sum(x^2) and the bound of max|x*w| arrive with the grid barrier (the
// phase that produced x reduced them into the barrier word), so x is read once.  Written for the persistent kernel's
// latency: 128-bit loads only (the host checks K % 4 == 0 and the alignment once); the RMSNorm weights of a phase are
// static and usually come from HBM, not L2, so the thread's share is fetched BEFORE the grid barrier (XPre) and the x
// loads are issued the moment the barrier opens.
constexpr int kXBatch = 6;   // float4 vectors per thread per batch: K <= 12288 converts in one batch
// The three scalars the conversion needs, from the statistics of x: y = (x * inv_rms) * w, xf = round(y * inv_s), y ~= s_x * xf
struct XScale {
    float inv_rms, inv_s, s_x;
};
__device__ __forceinline__ XScale make_xscale(XStats st, bool has_norm, int K, float rms_eps) {
    XScale sc;
    sc.inv_rms = 1.f;
    float amax = st.am;
    if (has_norm) {
        sc.inv_rms = rsqrtf(st.ss / (float)K + rms_eps);  // :1501
        amax = amax * sc.inv_rms * 1.00001f;                 // a bound of max|y|, y = x * inv_rms * w
    }
    const bool finite = amax > 0.f && amax < INFINITY;
    sc.inv_s = finite ? __fdividef(kXQMax, amax) : 0.f;   // |y * inv_s| <= kXQMax * (1 + 2e-5)
    sc.s_x = finite ? amax * (1.0f / kXQMax) : 0.f;
    return sc;
}
struct XPre {
    float4 w[kXBatch];
};
__device__ __forceinline__ void gemv_x_prefetch(const GemvArgs& a, int tid, XPre& p) {
    const float4* nw4 = reinterpret_cast<const float4*>(a.norm_w);
    const int kvec = a.L.K >> 2;
#pragma unroll
    for (int i = 0; i < kXBatch; ++i) {
        const int v = tid + i * kConsumerThreads;
        p.w[i] = (nw4 != nullptr && v < kvec) ? ldg_stream4(nw4 + v) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
}
template <int BITS>
__device__ __forceinline__ float gemv_stage_x_lean(const GemvArgs& a, const float* x, const GemvSmem& sm, const Slab& slab, bool coherent,
                                                   XScale sc, XPre& pre, bool load_w, int tid, int lane, long long* ts = nullptr) {
    const int K = a.L.K, nvec = layout_kpad(a.L) >> 2, kvec = K >> 2;
    const float4* nw4 = reinterpret_cast<const float4*>(a.norm_w);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    constexpr int B = kXBatch;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f), one4 = make_float4(1.f, 1.f, 1.f, 1.f);
    float4 xv[B];
    auto issue = [&](int v0, bool with_w) {
#pragma unroll
        for (int i = 0; i < B; ++i) {
            const int v = v0 + i * kConsumerThreads;
            xv[i] = v < kvec ? (coherent ? __ldcg(x4 + v) : x4[v]) : zero4;
            if (with_w) pre.w[i] = (nw4 != nullptr && v < kvec) ? ldg_stream4(nw4 + v) : one4;
        }
    };
    issue(tid, load_w);
    if (ts) ts[0] = clock64();
    // (the column accumulators are zero already: every epilogue clears what it has read)
#if !TIB_ZERO_IN_EPI
    for (int i = tid; i < slab.ncols * 3; i += kConsumerThreads) sm.acc[i] = 0;
#endif
    const float inv_rms = sc.inv_rms, inv_s = sc.inv_s, s_x = sc.s_x;
    if (ts) ts[1] = clock64();
    long long sxf = 0;
    int v0 = tid;
    while (true) {
#pragma unroll
        for (int i = 0; i < B; ++i) {
            const int v = v0 + i * kConsumerThreads;
            if (v < nvec) {
                float4 t = xv[i];
                // the reference divides by rms (:1504-1506); multiplying by the reciprocal differs by ~1 ulp, far below
                // the 2^-24 max|y| granularity of the fixed-point conversion that follows
                t.x = (t.x * inv_rms) * pre.w[i].x;
                t.y = (t.y * inv_rms) * pre.w[i].y;
                t.z = (t.z * inv_rms) * pre.w[i].z;
                t.w = (t.w * inv_rms) * pre.w[i].w;
                x_store_digits<BITS>(sm.xd, v, t, inv_s, sxf);
            }
        }
        v0 += B * kConsumerThreads;
        if (v0 >= nvec) break;
        issue(v0, true);
    }
    if (ts) ts[2] = clock64();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sxf += __shfl_xor_sync(0xffffffffu, sxf, o);
    if (lane == 0) sm.sxf[tid >> 5] = sxf;
    bar_sync(1, kConsumerThreads);
    return s_x;
}

// consumers, main loop: one quad per warp per ring stage -- 4 k-items, each one LDS.128 of weights (the A fragments),
// one LDS of digits (the B fragments) and 2 (INT4) / 1 (INT8) IMMA; when the warp's run over a group ends, the C
// fragments are added to the column sums in shared memory (no cross-lane reduction needed).
// Where a warp starts in its slab: depends on the slab's geometry only, so the persistent kernel computes it BEFORE the
// grid barrier (it holds an integer division).
struct ConsumePlan {
    int my_nq;        // quads of this warp
    int grp, chunk;   // group / chunk of its first quad
    int l_nq, l_fq;   // quads / first quad of warp (lane & 15): lanes stand for warps when stage offsets are summed
};
__device__ __forceinline__ ConsumePlan make_consume_plan(const QLayout& L, const Slab& slab, int warp, int lane) {
    ConsumePlan p;
    p.my_nq = warp_quads(slab, warp);
    const int fq = warp_first_quad(slab, warp);
    p.grp = fq / L.nchunks;
    p.chunk = fq - p.grp * L.nchunks;
    p.l_nq = warp_quads(slab, lane & 15);
    p.l_fq = warp_first_quad(slab, lane & 15);
    return p;
}
// DBG: 0 production; 1 = no MMAs, 2 = no barrier waits (stand-alone kernel experiments); 3 = timeline stamps through `dbg`
// (the persistent kernel's debug instance; `xskip` then also takes the timing switches 8 = no MMAs, 16 = no digit loads).
//
// The loop is written for the instruction issue slots: a round is only 4 k-items per warp (4 LDS.128 of weights, 4 of
// digits, 8 IMMA and -- INT4 -- 32 mask LOP3), so every instruction of bookkeeping around them shows: the first version of
// this loop ran ~170 instructions per round and warp and was issue-bound at 0.8 us per stage with the whole ring already in
// shared memory (profiles/r02_timeline_*: `rounds`).  Hence: barrier and stage addresses as 32-bit shared-window values
// advanced incrementally, the per-round offset sum (REDUX) only on rounds that can be ragged -- a warp whose quad is not in
// the slab's last group has only full quads before it, so its offset is 4 * warp items -- and no debug code in the
// production instance.
template <int BITS, int DBG = 0>
__device__ __forceinline__ void gemv_consume(const GemvArgs& a, const Slab& slab, const GemvSmem& sm, RingPos& it, const ConsumePlan& pl,
                                             int warp, int lane, long long* dbg = nullptr, bool ready0 = false, int xskip = 0) {
    const uint32_t S = (uint32_t)a.stages;
    const int C = a.L.nchunks, nrounds = slab.rounds;
    constexpr int kChunkBytes = BITS == 4 ? 768 : 384;   // digits of one chunk: 4 k-items x 3 planes x 4 t x 16 / 8 B
    constexpr int kXItem = kChunkBytes / 4;
    const int my_nq = pl.my_nq;
    int grp = pl.grp, chunk = pl.chunk;   // group / chunk of the warp's next quad
    const int g = lane >> 2, t = lane & 3;
    // B column g = digit plane g.  Columns 3..7 of the product are never read; their lanes duplicate a plane that is read in
    // the SAME quarter-warp (g = 3 -> plane 2, g >= 4 -> plane 0), so that the extra lanes are pure broadcasts: pointing all of
    // them at plane 0 put planes 0 and 2 -- the same banks -- into one quarter-warp of the LDS.128, a 2-way bank conflict on
    // every digit load (15 % of the kernel's shared-memory load wavefronts in the ncu capture of profiles/r02_*)
    const uint32_t xlane = smem_u32(sm.xd) + ((g < 3 ? g : (g == 3 ? 2 : 0)) * 4 + t) * (BITS == 4 ? 16 : 8);
    const bool xon = g < 3;
    // lane l < 16 stands for warp l when the stage offsets are summed with one REDUX (ragged rounds only)
    const int l_nq = pl.l_nq, l_fq = pl.l_fq;
    const bool l_before = lane < warp;  // warp < 16
    const int last_grp = slab.nlast == 4 ? -1 : slab.ngroups - 1;   // the only group whose quads are smaller than 4 items
    QuadAcc<BITS> acc;
    acc.clear();
    bool dirty = false;
    if (DBG == 3 && dbg != nullptr) {
        dbg[6] = clock64();
        // how many ring stages have landed when the loop starts (the producer runs ahead during the hand-over), and how many
        // rounds the slab has; dbg[-10 .. -6] below = completion time of rounds 0..4
        int nready = 0;
        RingPos probe = it;
        for (int i = 0; i < (int)S && i < nrounds; ++i, probe.advance(S)) nready += mbar_test_wait(&sm.full[probe.st], probe.par) ? 1 : 0;
        dbg[-1] = nready;
        dbg[5] = nrounds;
    }
    auto flush = [&]() {
        // C fragment: c0, c1 = (row g, B columns 2t, 2t+1), c2, c3 = (row g + 8, same columns); B column = digit plane
        const int rows = grp == slab.ngroups - 1 ? 4 * slab.nlast : 16;
        if (t < 2) {
            int* dst = sm.acc + (grp * 16 + g) * 3 + 2 * t;
            if (g < rows) {
                atomicAdd(dst, acc.total(0));
                if (t == 0) atomicAdd(dst + 1, acc.total(1));
            }
            if (g + 8 < rows) {
                atomicAdd(dst + 24, acc.total(2));
                if (t == 0) atomicAdd(dst + 25, acc.total(3));
            }
        }
        acc.clear();
        dirty = false;
    };
    auto one = [&](int i, const uint4& wv, const uint4& xv) {
        if (DBG == 1 || (DBG == 3 && (xskip & 8))) acc.lo[0][0] += (int)(wv.x ^ wv.y ^ wv.z ^ wv.w);
        else kitem_mma<BITS>(acc, i % TIB_ACC_SETS, wv, xv);
    };
    // Ring position: stage index and parity; every address of a round is a constant offset from ONE register holding the
    // block's shared-window base (mbarriers at +0 / +48, ring at +128).  The base, the lane's offset and the digit address are
    // made opaque to the compiler once: left to itself it re-derives them from %tid and the CTA id (S2R + several ALU
    // instructions) in every round.
    uint32_t st = it.st, par = it.par;
    uint32_t sbase = smem_u32(sm.full);
    uint32_t wfast = (uint32_t)(kBarBlockBytes + warp * 4 * kItemBytes + lane * 16);   // the warp's slot in a stage when every warp before has a full quad
    uint32_t xq = xlane + chunk * kChunkBytes;
    asm volatile("" : "+r"(sbase), "+r"(wfast), "+r"(xq));
    // the NEXT stage's barrier is probed while this stage is being multiplied; the first one may have been probed by the
    // caller during its prologue (polling an mbarrier costs 0.1-0.2 us even when its phase is long complete)
    bool ready = ready0;
    for (int r = 0; r < nrounds; ++r) {
        const bool have = r < my_nq;
        const uint32_t stage = sbase + st * kStageBytes, fullb = sbase + st * 8, emptyb = fullb + kMaxStages * 8;
        const uint32_t nst = st + 1 == S ? 0u : st + 1, npar = par ^ (nst == 0u ? 1u : 0u);
        // wait until the stage has landed
#if TIB_LOOP_SYNC == 2
        if (warp == 0 && lane == 0 && !ready) mbar_wait_s(fullb, par);
        bar_sync(4, kConsumerThreads);
#endif
        if (have && grp != last_grp) {
            // ---- fast path: a full quad, and only full quads before it in the stage ----
            uint4 xa, xb;
            if (!(DBG == 3 && (xskip & 16))) {   // digits of the first two k-items: they do not depend on the stage, so before the wait
                xa = load_xfrag<BITS>(xq, xon);
                xb = load_xfrag<BITS>(xq + kXItem, xon);
            } else {
                xa = xb = make_uint4(1u, 2u, 3u, 4u);
            }
            if constexpr (DBG == 3) { if (dbg && r == 0) dbg[7] = clock64(); }
#if TIB_LOOP_SYNC != 2
            if (DBG != 2 && !ready) mbar_wait_s(fullb, par);
#endif
            if constexpr (DBG == 3) { if (dbg && r == 0) dbg[0] = clock64(); }
            const uint32_t wbase = stage + wfast;
            const uint4 w0 = lds128s(wbase), w1 = lds128s(wbase + 512);
            one(0, w0, xa); one(1, w1, xb);
            const uint4 w2 = lds128s(wbase + 1024), w3 = lds128s(wbase + 1536);
            if (!(DBG == 3 && (xskip & 16))) {
                xa = load_xfrag<BITS>(xq + 2 * kXItem, xon);
                xb = load_xfrag<BITS>(xq + 3 * kXItem, xon);
            }
#if TIB_LOOP_SYNC == 0
            ready = (DBG != 2 && r + 1 < nrounds) ? mbar_test_wait_s(sbase + nst * 8, npar) : false;
#else
            ready = false;
#endif
            one(2, w2, xa); one(3, w3, xb);
            dirty = true;
            xq += kChunkBytes;
            if (++chunk == C) { flush(); chunk = 0; ++grp; xq = xlane; }
        } else {
            // ---- generic path: ragged last group (quads of nlast items), or no quad for this warp in the last round ----
            const int l_items = (l_before && r < l_nq) ? (l_fq + r >= slab.qfull ? slab.nlast : 4) : 0;
            const int woff = __reduce_add_sync(0xffffffffu, l_items);   // 512-byte items of the warps before this one in the stage
            if constexpr (DBG == 3) { if (dbg && r == 0) dbg[7] = clock64(); }
#if TIB_LOOP_SYNC != 2
            if (DBG != 2 && !ready) mbar_wait_s(fullb, par);
#endif
            if constexpr (DBG == 3) { if (dbg && r == 0) dbg[0] = clock64(); }
#if TIB_LOOP_SYNC == 0
            ready = (DBG != 2 && r + 1 < nrounds) ? mbar_test_wait_s(sbase + nst * 8, npar) : false;
#else
            ready = false;
#endif
            if (have) {
                const int nl = slab.nlast;   // rows 0..7 in part A, rows 8..11 (nl = 3) in part B
                const bool in_a = nl > 1 || lane < 16, in_b = nl == 3 && lane < 16;
                const uint32_t wbase = stage + kBarBlockBytes + woff * kItemBytes + lane * 8;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint2 pa = make_uint2(0u, 0u), pb = make_uint2(0u, 0u);
                    if (in_a) pa = lds64s(wbase + i * nl * 128);
                    if (in_b) pb = lds64s(wbase + i * nl * 128 + 256);
                    const uint4 xv = load_xfrag<BITS>(xq + i * kXItem, xon);
                    one(i, make_uint4(pa.x, pb.x, pa.y, pb.y), xv);
                }
                dirty = true;
                xq += kChunkBytes;
                if (++chunk == C) { flush(); chunk = 0; ++grp; xq = xlane; }
            }
        }
#if TIB_LOOP_SYNC == 2
        bar_sync(5, kConsumerThreads);
        if (warp == 0 && lane == 0) mbar_arrive_s(emptyb);   // every warp is done reading the stage (the empty barriers count 1 arrival)
#else
        __syncwarp();
        if (DBG != 2 && lane == 0) mbar_arrive_s(emptyb);  // this warp is done reading the stage
#endif
        if constexpr (DBG == 3) { if (dbg && r < 5) dbg[r - 10] = clock64(); }
        st = nst;
        par = npar;
    }
    it.st = st;
    it.par = par;
    if constexpr (DBG == 3) { if (dbg) dbg[1] = clock64(); }
    if (dirty) flush();
    if constexpr (DBG == 3) { if (dbg) dbg[2] = clock64(); }
    bar_sync(1, kConsumerThreads);
    if constexpr (DBG == 3) { if (dbg) dbg[3] = clock64(); }
}

// Values the epilogue needs that do not depend on this phase's arithmetic; loaded early (before the grid barrier
// in the persistent kernel) so their latency is off the critical path.  They cover the thread's first column
// (or column pair); further columns, if a slab has more than 512, load theirs in place.
struct EpiPre {
    float cs0, cs1;   // colscale
    float zt0, zt1;   // colzterm (0 when absent)
    float r0;         // residual input (EPI_RESIDUAL)
    float nw0;        // RMSNorm weight the next GEMV applies to this column (1 when none): for the output statistics
    float invf;       // inv_freq entry (EPI_QKV with RoPE)
    int page;         // physical KV page of the current position (EPI_QKV)
};

__device__ __forceinline__ EpiPre gemv_epilogue_prefetch(const GemvArgs& a, const Slab& slab, const float* resid, const PhaseCtx& ctx, int tid) {
    EpiPre p{0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0};
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
        if (a.epi == EPI_RESIDUAL)
            p.r0 = ld_act(resid + n, ctx.coherent);
        if (!pairs && a.next_norm_w) p.nw0 = a.next_norm_w[n];
        if (a.epi == EPI_QKV) {
            const int pos = ctx.pos >= 0 ? ctx.pos : *a.pos_ptr;
            p.page = a.page_table[pos / a.page_tokens];
            if (a.rope_dim > 0) p.invf = a.inv_freq[((n % a.hidden) % a.rope_dim) >> 1];
        }
    }
    return p;
}

// consumers, epilogue: combine the digit sums exactly, scale once, fused tail op
// Returns this thread's contribution to the XStats of the output (EPI_RESIDUAL / EPI_SWIGLU / EPI_RELU / EPI_STORE).
__device__ __forceinline__ XStats gemv_epilogue(const GemvArgs& a, const Slab& slab, const GemvSmem& sm, float s_x, const float* resid,
                                                const PhaseCtx& ctx, const EpiPre& pre, int tid, int lane) {
    XStats out_st{0.f, 0.f};
    const QLayout& L = a.L;
    const int ncols = slab.ncols;
    long long sxf = 0;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) sxf += sm.sxf[i];
    const long long offterm = (long long)a.woff * sxf;
    const float fsxf = (float)sxf;
    // y = cs * s_x * (sum_k u_k xf_k - off * sum xf + zt * sum xf)
    // reads the three digit sums of column c and leaves them zero for the next GEMV that uses this shared memory (the
    // persistent kernel's prologue no longer spends a pass on clearing them)
    auto colval = [&](int c, float cs, float zt) -> float {
        int* p = sm.acc + c * 3;
        const int a0 = p[0], a1 = p[1], a2 = p[2];
#if TIB_ZERO_IN_EPI
        p[0] = 0; p[1] = 0; p[2] = 0;
#endif
        const long long t = ((long long)a2 << 16) + ((long long)a1 << 8) + (long long)a0 - offterm;
        // |t| < 2^47: hi * 2^23 + lo with both parts exact in fp32, so tf is the correctly rounded t (no fp64 needed)
        const int hi = (int)(t >> 23), lo = (int)(t & 0x7FFFFF);
        const float tf = fmaf((float)hi, 8388608.0f, (float)lo);
        return (fmaf(zt, fsxf, tf) * s_x) * cs;
    };
    if (a.epi == EPI_SWIGLU || a.epi == EPI_QKV) {
        for (int pc = tid; pc < ncols / 2; pc += kConsumerThreads) {  // column pairs
            const int n0 = slab.col0 + 2 * pc;
            if (n0 >= L.N) {   // padding columns of the last slab: nothing to store, but their sums must not survive
                for (int i = 0; i < 6; ++i) sm.acc[6 * pc + i] = 0;
                continue;
            }
            const bool first = pc == tid;
            const float cs0 = first ? pre.cs0 : a.colscale[n0], cs1 = first ? pre.cs1 : a.colscale[n0 + 1];
            const float zt0 = first ? pre.zt0 : (a.colzterm ? a.colzterm[n0] : 0.f);
            const float zt1 = first ? pre.zt1 : (a.colzterm ? a.colzterm[n0 + 1] : 0.f);
            const float y0 = colval(2 * pc, cs0, zt0), y1 = colval(2 * pc + 1, cs1, zt1);
            if (a.epi == EPI_SWIGLU) {
                const float sg = y0 / (1.0f + expf(-y0));  // silu(gate), :918
                const float o = y1 * sg;                    // multiply(up, silu(gate))
                a.out[n0 >> 1] = o;
                out_st.am = fmaxf(out_st.am, fabsf(o));
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
            if (n >= L.N) {
                for (int i = 0; i < 3; ++i) sm.acc[3 * c + i] = 0;
                continue;
            }
            const bool first = c == tid;
            const float cs = first ? pre.cs0 : a.colscale[n];
            const float zt = first ? pre.zt0 : (a.colzterm ? a.colzterm[n] : 0.f);
            float y = colval(c, cs, zt);
            if (a.epi == EPI_RESIDUAL) y = (first ? pre.r0 : ld_act(resid + n, ctx.coherent)) + y;
            else if (a.epi == EPI_RELU) y = fmaxf(y, 0.f);
            if (ctx.npeers > 0) {
                if (ctx.ll_seq != 0u) {
                    const unsigned long long word = ((unsigned long long)ctx.ll_seq << 32) | (unsigned long long)__float_as_uint(y);
                    for (int r = 0; r < ctx.npeers; ++r)
                        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(ctx.peers[r]) + ctx.peer_off + n), "l"(word)
                                     : "memory");
                } else {
                    for (int r = 0; r < ctx.npeers; ++r) ctx.peers[r][ctx.peer_off + n] = y;   // NVLink stores (one local)
                }
            } else {
                a.out[n] = y;
            }
            if (a.epi == EPI_LOGITS) {
                if (y > best) { best = y; besti = n + a.col_off; }
            } else {
                out_st.ss = fmaf(y, y, out_st.ss);
                out_st.am = fmaxf(out_st.am, fabsf(y * (first ? pre.nw0 : (a.next_norm_w ? a.next_norm_w[n] : 1.f))));
            }
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
    return out_st;
}

__device__ __forceinline__ void gemv_init_barriers(const GemvSmem& sm, int stages) {
    for (int i = 0; i < stages; ++i) {
        mbar_init(&sm.full[i], 1);
        mbar_init(&sm.empty[i], TIB_LOOP_SYNC == 2 ? 1 : kConsumerWarps);
    }
    fence_mbar_init();
}

// ---- the stand-alone kernel: one GEMV per launch -----------------------------------------------------------
template <int BITS, int DBG = 0>
__global__ void __launch_bounds__(kGemvThreads, 1) gemv_kernel(const __grid_constant__ GemvArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Slab slab = make_slab(a.L, blockIdx.x);
    const GemvSmem sm = gemv_carve(smem_raw, a.L, a.stages);
    if (tid == 0) gemv_init_barriers(sm, a.stages);
    __syncthreads();
    // programmatic dependent launch: the NEXT kernel of the stream may be scheduled as soon as this CTA's resources free up --
    // its producer then streams weights and its consumers park at griddepcontrol.wait until this grid has completed
    pdl_launch_dependents();
    RingPos it;
    if (warp == kConsumerWarps) {
        // the weights do not depend on the previous kernel: start streaming at once
        gemv_produce(a, slab, sm, it, lane);
        return;
    }
    pdl_wait_prior_grid();  // x (and resid / pos) come from the previous kernel in the stream
    const PhaseCtx ctx{false, -1, nullptr, nullptr, 0, 0, 0u};
    const EpiPre pre = gemv_epilogue_prefetch(a, slab, a.resid, ctx, tid);
    const float s_x = gemv_stage_x<BITS>(a, a.x, sm, slab, false, tid, warp, lane);
    gemv_consume<BITS, DBG>(a, slab, sm, it, make_consume_plan(a.L, slab, warp, lane), warp, lane);
    (void)gemv_epilogue(a, slab, sm, s_x, a.resid, ctx, pre, tid, lane);
}

}  // namespace tib
