// mega.cuh -- the persistent decode kernel: ONE cooperative launch runs whole forward passes (and whole greedy
// generations) of the decoder, replacing the per-op launches of InferenceEngine::forward_pass_incremental
// (src/model/inference_engine.cpp:1493-1552) -> TransformerLayer::forward_incremental (:244-401) and the host-side
// sampling loop of generate (:752-775).
//
// Why: at batch 1 a layer is ~100 MB (7B INT4) = ~15 us of HBM time split over 5 dependent GEMVs; separate kernels
// pay launch + drain + cold-pipeline latency (~6 us measured) on each of them.  Here one CTA per SM stays resident:
//   * the producer warp streams the CTA's weight slabs of ALL phases back to back through the shared-memory ring
//     (weights never depend on activations), so HBM keeps flowing while the consumers synchronise;
//   * consumers run the phases  [QKV+RMSNorm+RoPE+KV-append] [attention split-K + combine] [O-proj+residual]
//     [gate/up+RMSNorm+SwiGLU] [down+residual] per layer, then [lm_head+RMSNorm+argmax], separated by a
//     grid-wide barrier (release/acquire counter in global memory) instead of a kernel boundary;
//   * the greedy token is decoded from the argmax key by every CTA, so the next forward pass starts without
//     touching the host.
#pragma once
#include "kernels.cuh"

namespace tib {

enum MegaPhaseType : int { PH_GEMV = 0, PH_ATTN = 1, PH_REDUCE = 2, PH_KEYX = 3 };
constexpr int kMaxTp = 8;
struct MegaArgs;
constexpr int kStampsPerPhase = 32;  // debug timeline: 6 phase-level stamps, [11] ring stages ready, [12..] finer stamps
enum MegaSrc : int { SRC_PTR = 0, SRC_EMB = 1 };

struct MegaPhase {
    int type;
    int x_src;      // PH_GEMV: where x comes from (pointer in g.x, or the embedding row of the current token)
    int resid_src;  // PH_GEMV + EPI_RESIDUAL: same for the residual input
    int is_head;    // lm_head phase: only run on steps that sample
    // tensor parallel: a row-parallel GEMV phase (mgpu = 1) stores its partial output into EVERY rank's
    // partial buffer `part_sel` and ends with a barrier across all GPUs; the PH_REDUCE phase after it sums the partials in
    // rank order on top of the residual (g.x = this rank's partial buffer [tp][H], g.resid / g.out = the residual stream,
    // g.L.N = H, g.next_norm_w as in a residual epilogue).  mgpu = 2: point-to-point variant -- no barrier across the GPUs
    // and no PH_REDUCE phase: the CTA waits for the flags of its own column slice from every rank and reduces the slice
    // in the tail of the GEMV phase (g.resid / g.out / g.next_norm_w as in the single-GPU residual epilogue)
    int mgpu, part_sel;
    GemvArgs g;
    AttnArgs at;
};

// What the producer warp needs of a phase, compact enough for a lane to hold a whole record in registers: the
// producer runs AHEAD of the consumers, so it must never wait on a chain of dependent global loads per phase.
struct ProdRec {
    unsigned long long wq;   // packed weights
    QLayout L;
    int flags;               // bit 0: GEMV phase, bit 1: lm_head (only on sampling steps)
};
static_assert(sizeof(ProdRec) == 40, "ProdRec is shuffled as 10 words");

struct MegaArgs {
    const MegaPhase* __restrict__ phases;
    const ProdRec* __restrict__ prod;   // [nphases]
    int nphases;
    const float* emb;     // [V][H]
    int H, V;
    int literal_embed;    // unused in this kernel (compat path is separate)
    StepState* st;
    const StepIO* io;
    const int* prompt;    // tokens for steps < n_prompt
    int n_prompt;
    int n_steps;          // forward passes in this launch
    int first_sample;     // steps >= first_sample run the lm_head and pick a token
    unsigned int* grid_bar;       // kBarWords barrier words, 128 bytes apart; zeroed by the host before every launch
    unsigned int* head_cnt;       // [heads]; zero between phases by construction
    unsigned long long* keys;     // [2] argmax keys, zeroed by the host before every launch
    float* logits;
    int stages;
    int max_kpad, max_units, attn_floats;
    long long* dbg;   // optional: CTA 0 writes 6 clock64 stamps per phase of step 0 (debug timeline)
    int dbg_flags;    // A/B switches (debug): 1 = norm weights loaded after the barrier, 2 = no L1 prefetch of the next descriptor
    // tensor parallel: peer pointers (CUDA IPC mappings of every rank's exchange block, own rank included)
    int tp, tp_rank;
    float* peer_part[2][kMaxTp];       // [buffer][rank] -> that rank's partial buffer, layout [source rank][H]
    unsigned int* peer_bar[kMaxTp];    // [rank] -> that rank's multi-GPU barrier words (kBarWords x kBarStride)
    unsigned int* mg_seq;              // two device words carried across launches: [0] barriers across the GPUs passed so far, [1] point-to-point exchanges
    // point-to-point variant (tp_p2p): CTA b of every rank owns the same column slice, so only the P CTAs "b" have to meet:
    // flags [source rank][CTA] per partial buffer, written by the source with a system-scope release
    unsigned int* peer_flag[2][kMaxTp];
    int tp_p2p;
    // tp_ll (with tp_p2p): the partials travel as 8-byte words {value, exchange number} -- a reader polls the words themselves, so an
    // exchange costs ONE NVLink crossing (no release fence waiting for the stores' acknowledgements, no separate flag)
    int tp_ll;
    // tensor parallel with the lm_head sharded over the vocabulary: every rank's local (max, index) key of a
    // sampling step goes to slot [step & 1][rank] of every rank's exchange block (PH_KEYX phase), the barrier across the GPUs
    // follows, and the token is the maximum of the P keys -- n_head = 2 phases run on sampling steps only (lm_head, PH_KEYX)
    unsigned long long* peer_keys[kMaxTp];
    int n_head;
    const XStats* emb_stats;  // [V]: statistics of every embedding row (against the first norm weight)
    // the sequence's KV page table (static for the launch), copied to shared memory once: kv_pages entries (0: none)
    const int* kv_page_table;
    int kv_pages;
};
constexpr int kMaxSmemPages = 1024;   // 64 k tokens at 64 tokens per page; longer tables are read from global memory

TIB_HD size_t mega_smem_bytes(int stages, int max_kpad, int max_units, int attn_floats, int kv_pages = 0) {
    const size_t table = kv_pages <= kMaxSmemPages ? (((size_t)kv_pages * 4 + 15) & ~size_t(15)) : 0;
    return gemv_smem_bytes_for(stages, max_kpad, max_units) + 16 + (size_t)attn_floats * 4 + 16 + 3 * ((sizeof(MegaPhase) + 15) & ~size_t(15)) + table;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- the grid barrier -------------------------------------------------------------------------------------------
// One 16-byte word per barrier instance, {arrivals, max|y*w| as float bits, sum(y^2) in 2^-28 fixed point (64 bits)}:
// every CTA reduces its partial statistics of the phase's output into the word (integer atomics: exact, so the
// result does not depend on the arrival order) BEFORE its release-increment of the arrival count, and the one polling
// thread per CTA reads count and statistics with a single 16-byte load -- the next phase's prologue needs no second
// round trip to L2 (measured: gathering 148 per-CTA slots after the barrier cost 1.2-1.9 us per phase).  Four words in
// rotation, each in its own 128-byte line; CTA 0 clears word (k-1) & 3 after passing barrier k (every CTA arrived at
// k, so every CTA has finished reading k-1), three barriers before its next use.
constexpr int kBarWords = 4, kBarStride = 32;   // in 32-bit units
constexpr long long kBarWatchdog = 120000000000LL;   // ~60 s of SM clocks: a protocol bug traps instead of hanging the box; long enough for profiler replay and time slicing
constexpr float kSsScale = 268435456.0f;         // 2^28: resolution 3.7e-9, range 6.9e10
__device__ __forceinline__ void bar_arrive_stats(unsigned int* w, float ss, float am) {
    const unsigned long long q = __float2ull_rn(fminf(ss, 6.0e10f) * kSsScale);
    asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(w + 1), "r"(__float_as_uint(am)) : "memory");
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(w + 2), "l"(q) : "memory");
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(w) : "memory");
}
// the same across GPUs: arrivals are system-scope atomics on every rank's word (NVLink), the poll is on the local one
__device__ __forceinline__ void bar_arrive_sys(unsigned int* w) {
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(w) : "memory");
}
__device__ __forceinline__ unsigned int bar_poll_sys(const unsigned int* w) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(w) : "memory");
    return v;
}
__device__ __forceinline__ uint4 bar_poll(const unsigned int* w) {
    uint4 v;
    asm volatile("ld.acquire.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(w) : "memory");
    return v;
}

// ---- attention work item (head h, split j) for NT cooperating threads -----------------------------------
template <int NT>
__device__ __forceinline__ float nt_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    bar_sync(1, NT);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) s += red[i];
    bar_sync(1, NT);
    return s;
}
template <int NT>
__device__ __forceinline__ float nt_max(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    bar_sync(1, NT);
    float s = -INFINITY;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) s = fmaxf(s, red[i]);
    bar_sync(1, NT);
    return s;
}

TIB_HD int attn_fast_scratch_floats(int nt);
TIB_HD int attn_lean_scratch_floats(int nt);
TIB_HD int attn_scratch_floats(int D, int nt) {
    const int groups = D < nt ? nt / D : 1;
    const int generic = D + kAttnTokBlock + 32 + groups * D + 4;
    int fast = D <= 128 ? attn_fast_scratch_floats(nt) + 4 : 0;   // + the merge flag of mega_attention
    if (D <= 128 && attn_lean_scratch_floats(nt) > fast) fast = attn_lean_scratch_floats(nt);
    return generic > fast ? generic : fast;
}

// Computes the (m, l, o) partial of head h over tokens [t0, t1) and stores it.  q / K / V are read through L2.
// Returns the thread's max |value written to a.out| (direct items only).
template <int NT>
__device__ __forceinline__ float attn_item(const AttnArgs& a, int h, int j, int t0, int t1, float* sm, bool direct = false) {
    float out_am = 0.f;
    float* qs = sm;
    float* sc = qs + a.D;
    float* red = sc + kAttnTokBlock;
    float* ored = red + 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = a.D, hoff = h * D;
    for (int d = tid; d < D; d += NT) qs[d] = __ldcg(a.q + hoff + d);
    bar_sync(1, NT);
    const int groups = D < NT ? NT / D : 1;
    const int grp = D < NT ? tid / D : 0;
    const int d0 = D < NT ? tid % D : tid;
    const bool active = grp < groups;
    constexpr int kMaxDims = 8192 / NT;
    float o[kMaxDims];
#pragma unroll
    for (int i = 0; i < kMaxDims; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int tb = t0; tb < t1; tb += kAttnTokBlock) {
        const int nt = min(kAttnTokBlock, t1 - tb);
        for (int tt = warp; tt < nt; tt += NT / 32) {
            const float* kr = kv_row(a.k_pool, a.page_table, a.page_tokens, a.H, tb + tt) + hoff;
            float s = 0.f;
            for (int d = 4 * lane; d < D; d += 128) {
                const float4 kv = __ldcg(reinterpret_cast<const float4*>(kr + d));
                s = fmaf(qs[d], kv.x, s);
                s = fmaf(qs[d + 1], kv.y, s);
                s = fmaf(qs[d + 2], kv.z, s);
                s = fmaf(qs[d + 3], kv.w, s);
            }
            s = warp_sum(s);
            if (lane == 0) sc[tt] = s * a.scale;
        }
        bar_sync(1, NT);
        float mx = -INFINITY;
        for (int tt = tid; tt < nt; tt += NT) mx = fmaxf(mx, sc[tt]);
        mx = nt_max<NT>(mx, red);
        const float m_new = fmaxf(m_run, mx);
        float ls = 0.f;
        for (int tt = tid; tt < nt; tt += NT) {
            const float p = expf(sc[tt] - m_new);
            sc[tt] = p;
            ls += p;
        }
        ls = nt_sum<NT>(ls, red);
        const float corr = expf(m_run - m_new);
        l_run = l_run * corr + ls;
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < kMaxDims; ++i) {
            const int d = d0 + i * NT;
            if (i == 0 || d < D) {
                float acc = o[i] * corr;
                if (d < D && active)
                    for (int tt = grp; tt < nt; tt += groups) {
                        const float* vr = kv_row(a.v_pool, a.page_table, a.page_tokens, a.H, tb + tt) + hoff;
                        acc = fmaf(sc[tt], __ldcg(vr + d), acc);
                    }
                o[i] = acc;
            }
        }
        bar_sync(1, NT);
    }
    // direct: this item covers the whole context of the head -> normalise and write the output, no partials
    float* po = direct ? a.out + hoff : a.part_o + ((size_t)h * a.max_splits + j) * D;
    const float onorm = direct ? 1.0f / l_run : 1.0f;
    if (direct) {
#pragma unroll
        for (int i = 0; i < kMaxDims; ++i) o[i] *= onorm;
    }
    if (groups > 1) {
        if (active) ored[grp * D + d0] = o[0];
        bar_sync(1, NT);
        if (grp == 0) {
            float acc = 0.f;
            for (int g = 0; g < groups; ++g) acc += ored[g * D + d0];
            po[d0] = acc;
            if (direct) out_am = fabsf(acc);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kMaxDims; ++i) {
            const int d = d0 + i * NT;
            if (d < D) {
                po[d] = o[i];
                if (direct) out_am = fmaxf(out_am, fabsf(o[i]));
            }
        }
    }
    if (tid == 0 && !direct) {
        a.part_ml[((size_t)h * a.max_splits + j) * 2 + 0] = m_run;
        a.part_ml[((size_t)h * a.max_splits + j) * 2 + 1] = l_run;
    }
    return out_am;
}

// Low-latency item for head dims <= 128 (every Llama-family shape): a warp owns whole tokens, a lane owns 4 dims.
// Each warp issues the K AND V row loads of 4 tokens at once (they do not depend on q or on the scores), so a short
// context costs one memory round trip; scores are warp-shuffle dot products, the softmax is "online" per warp, and the
// 16 warps' (m, l, o) are merged through shared memory in warp order (deterministic).  Same arithmetic as
// attention_fast_incremental (:1254-1388) up to the order of the fp32 sums; expf is the full-precision one.
TIB_HD int attn_fast_scratch_floats(int nt) { return (nt / 32) * (128 + 2); }
template <int NT>
__device__ __forceinline__ float attn_item_fast(const AttnArgs& a, int h, int j, int t0, int t1, float* sm, bool direct, long long* ts = nullptr) {
    constexpr int NW = NT / 32, G = 5;   // 5 tokens in flight per warp: a 320-token item (4 splits of a 1280-token context) is one round trip
    float* wm = sm;                 // [NW]
    float* wl = wm + NW;            // [NW]
    float* wo = wl + NW;            // [NW][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = a.D, hoff = h * D;
    const bool lane_on = 4 * lane < D;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // the item's page ids, one per lane (an item of up to 32 pages): the K / V addresses then need no dependent global load
    const int ps = a.page_shift;
    const int p0 = ps >= 0 ? (t0 >> ps) : 0;
    const bool hoisted = ps >= 0 && ((t1 - 1) >> ps) - p0 < 32;
    int pg = 0;
    if (hoisted && p0 + lane <= ((t1 - 1) >> ps)) pg = a.page_table[p0 + lane];
    const int pmask = a.page_tokens - 1;
    auto row_of = [&](const float* pool, int t) -> const float* {
        if (hoisted) {
            const int page = __shfl_sync(0xffffffffu, pg, (t >> ps) - p0);
            return pool + ((size_t)page * a.page_tokens + (t & pmask)) * a.H;
        }
        return kv_row(pool, a.page_table, a.page_tokens, a.H, t);
    };
    // the first group's K / V loads go out before q is needed
    float4 kv[G], vv[G];
    auto issue = [&](int tbase) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int t = tbase + g * NW;
            const int tc = t < t1 ? t : t1 - 1;   // the shuffle inside row_of is warp-wide: every lane computes an in-range row
            const size_t row = (size_t)hoff + 4 * lane;
            const float* kr = row_of(a.k_pool, tc);
            const float* vr = row_of(a.v_pool, tc);
            if (t < t1 && lane_on) {
                kv[g] = __ldcg(reinterpret_cast<const float4*>(kr + row));
                vv[g] = __ldcg(reinterpret_cast<const float4*>(vr + row));
            } else {
                kv[g] = zero4;
                vv[g] = zero4;
            }
        }
    };
    if (ts) ts[11] = clock64();
    issue(t0 + warp);
    const float4 q = lane_on ? __ldcg(reinterpret_cast<const float4*>(a.q + hoff + 4 * lane)) : zero4;
    if (ts) ts[16] = clock64();
    float m_run = -INFINITY, l_run = 0.f;
    float4 o = zero4;
    for (int tbase = t0 + warp; tbase < t1; tbase += G * NW) {
        if (tbase != t0 + warp) issue(tbase);
        float s[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float d = q.x * kv[g].x;
            d = fmaf(q.y, kv[g].y, d);
            d = fmaf(q.z, kv[g].z, d);
            d = fmaf(q.w, kv[g].w, d);
            s[g] = d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int g = 0; g < G; ++g) s[g] += __shfl_xor_sync(0xffffffffu, s[g], off);
        }
        float mx = m_run;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            s[g] = tbase + g * NW < t1 ? s[g] * a.scale : -INFINITY;
            mx = fmaxf(mx, s[g]);
        }
        if (ts && tbase == t0 + warp) ts[22] = clock64() + (long long)(mx == 12345.678f);   // depends on the scores: after loads + shuffles
        const float corr = expf(m_run - mx);   // first group: exp(-inf) = 0
        l_run *= corr;
        o.x *= corr; o.y *= corr; o.z *= corr; o.w *= corr;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float p = expf(s[g] - mx);     // masked tokens: exp(-inf) = 0
            l_run += p;
            o.x = fmaf(p, vv[g].x, o.x);
            o.y = fmaf(p, vv[g].y, o.y);
            o.z = fmaf(p, vv[g].z, o.z);
            o.w = fmaf(p, vv[g].w, o.w);
        }
        m_run = mx;
    }
    if (ts) ts[7] = clock64();
    // merge the warps (a warp without tokens has m = -inf, l = 0, o = 0)
    if (lane == 0) { wm[warp] = m_run; wl[warp] = l_run; }
    *reinterpret_cast<float4*>(wo + warp * 128 + 4 * lane) = o;
    bar_sync(1, NT);
    float out_am = 0.f;
    if (tid < D) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) M = fmaxf(M, wm[w]);
        float Lsum = 0.f, acc = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float f = wm[w] == -INFINITY ? 0.f : expf(wm[w] - M);
            Lsum = fmaf(wl[w], f, Lsum);
            acc = fmaf(wo[w * 128 + tid], f, acc);
        }
        if (direct) {
            const float r = acc / Lsum;
            a.out[hoff + tid] = r;
            out_am = fabsf(r);
        } else {
            a.part_o[((size_t)h * a.max_splits + j) * D + tid] = acc;
            if (tid == 0) {
                a.part_ml[((size_t)h * a.max_splits + j) * 2 + 0] = M;
                a.part_ml[((size_t)h * a.max_splits + j) * 2 + 1] = Lsum;
            }
        }
    }
    bar_sync(1, NT);   // the scratch is reused by the next item; every thread's partial stores are issued
    if (ts) ts[8] = clock64();
    return out_am;
}

// ---- the attention phase of the persistent kernel for head dims <= 128: lean item ----------------------------------
// Same arithmetic as attn_item_fast, written for latency: the page table lives in shared memory (copied once per launch), the
// split geometry is computed once per step, and the 16 warps' (m, l, o) are merged with 16 exponentials instead of 16 per
// output element.  Contexts of up to min_chunk tokens (4 round trips of 80 tokens) are ONE item per head that writes the
// normalised output directly: at these lengths a split's counter + last-arriver merge (~2 us, and the merging CTA finishes
// after all others) costs more than the extra round trips.  (Deferring the merge to the o-projection's prologue was tried:
// 148 CTAs re-reading nsplit partial vectors from L2 at the same moment cost more than the merge it saved.)
// Returns the thread's max |output| (direct items only).
struct AttnStep {
    int nsplit, chunk;
};
__device__ __forceinline__ AttnStep attn_step_geometry(const AttnArgs& a, int t) {
    AttnStep g;
    attn_split_range(t, a.max_splits, a.min_chunk, g.nsplit, g.chunk);
    return g;
}
TIB_HD int attn_lean_scratch_floats(int nt) { return (nt / 32) * (128 + 3) + 4 + 4; }   // + the merge flag
template <int NT>
__device__ __forceinline__ float attn_item_lean(const AttnArgs& a, const int* spt, int h, int j, int t0, int t1, float* sm, bool direct, long long* ts) {
    constexpr int NW = NT / 32, G = 5;   // 5 tokens in flight per warp: an item of up to 80 tokens is one round trip.  (Two rounds in
                                         // flight -- double-buffered K / V registers -- was tried: the 48-64 extra registers spill.)
    float* wm = sm;                 // [NW] running max of each warp
    float* wl = wm + NW;            // [NW] its sum of exponentials
    float* wf = wl + NW;            // [NW] e^(m_w - M), then [NW] = M, [NW + 1] = L
    float* wo = wf + NW + 4;        // [NW][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = a.D, hoff = h * D;
    const bool lane_on = 4 * lane < D;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ps = a.page_shift, pmask = a.page_tokens - 1;
    const size_t lane_off = (size_t)hoff + 4 * lane;
    const int* table = spt != nullptr ? spt : a.page_table;
    float4 kv[G], vv[G];
    auto issue = [&](int tbase) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int t = tbase + g * NW;
            if (t < t1 && lane_on) {
                const size_t row = ((size_t)table[t >> ps] * a.page_tokens + (t & pmask)) * a.H + lane_off;
                kv[g] = __ldcg(reinterpret_cast<const float4*>(a.k_pool + row));
                vv[g] = __ldcg(reinterpret_cast<const float4*>(a.v_pool + row));
            } else {
                kv[g] = zero4;
                vv[g] = zero4;
            }
        }
    };
    if (ts) ts[11] = clock64();
    issue(t0 + warp);   // the first group's K / V loads go out before q is needed
    const float4 q = lane_on ? __ldcg(reinterpret_cast<const float4*>(a.q + hoff + 4 * lane)) : zero4;
    if (ts) ts[16] = clock64();
    float m_run = -INFINITY, l_run = 0.f;
    float4 o = zero4;
    for (int tbase = t0 + warp; tbase < t1; tbase += G * NW) {
        if (tbase != t0 + warp) issue(tbase);
        float s[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float d = q.x * kv[g].x;
            d = fmaf(q.y, kv[g].y, d);
            d = fmaf(q.z, kv[g].z, d);
            d = fmaf(q.w, kv[g].w, d);
            s[g] = d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int g = 0; g < G; ++g) s[g] += __shfl_xor_sync(0xffffffffu, s[g], off);
        }
        float mx = m_run;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            s[g] = tbase + g * NW < t1 ? s[g] * a.scale : -INFINITY;
            mx = fmaxf(mx, s[g]);
        }
        if (ts && tbase == t0 + warp) ts[22] = clock64() + (long long)(mx == 12345.678f);   // depends on the scores: after loads + shuffles
        const float corr = expf(m_run - mx);   // first group: exp(-inf) = 0
        l_run *= corr;
        o.x *= corr; o.y *= corr; o.z *= corr; o.w *= corr;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float p = expf(s[g] - mx);     // masked tokens: exp(-inf) = 0
            l_run += p;
            o.x = fmaf(p, vv[g].x, o.x);
            o.y = fmaf(p, vv[g].y, o.y);
            o.z = fmaf(p, vv[g].z, o.z);
            o.w = fmaf(p, vv[g].w, o.w);
        }
        m_run = mx;
    }
    if (ts) ts[7] = clock64();
    // merge the warps (a warp without tokens has m = -inf, l = 0, o = 0): warp 0 turns the 16 maxima into 16 factors
    if (lane == 0) { wm[warp] = m_run; wl[warp] = l_run; }
    *reinterpret_cast<float4*>(wo + warp * 128 + 4 * lane) = o;
    bar_sync(1, NT);
    if (warp == 0) {
        const float mw = lane < NW ? wm[lane] : -INFINITY;
        const float M = warp_max(mw);
        const float f = mw == -INFINITY ? 0.f : expf(mw - M);
        float L = lane < NW ? wl[lane] * f : 0.f;
        L = warp_sum(L);
        if (lane < NW) wf[lane] = f;
        if (lane == 0) { wf[NW] = M; wf[NW + 1] = L; }
    }
    bar_sync(1, NT);
    float out_am = 0.f;
    if (tid < D) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) acc = fmaf(wo[w * 128 + tid], wf[w], acc);
        if (direct) {
            const float r = acc / wf[NW + 1];
            a.out[hoff + tid] = r;
            out_am = fabsf(r);
        } else {
            a.part_o[((size_t)h * a.max_splits + j) * D + tid] = acc;
            if (tid == 0) {
                a.part_ml[((size_t)h * a.max_splits + j) * 2 + 0] = wf[NW];
                a.part_ml[((size_t)h * a.max_splits + j) * 2 + 1] = wf[NW + 1];
            }
        }
    }
    if (ts) ts[8] = clock64();
    return out_am;
}

// merge the nsplit partials of head h (fixed order, independent of which CTA does it)
template <int NT>
__device__ __forceinline__ float attn_merge_head(const AttnArgs& a, int h, int nsplit) {
    float out_am = 0.f;
    const float* ml = a.part_ml + (size_t)h * a.max_splits * 2;
    float M = -INFINITY;
    for (int j = 0; j < nsplit; ++j) M = fmaxf(M, __ldcg(ml + 2 * j));
    float Lsum = 0.f;
    for (int j = 0; j < nsplit; ++j) Lsum += __ldcg(ml + 2 * j + 1) * expf(__ldcg(ml + 2 * j) - M);
    for (int d = threadIdx.x; d < a.D; d += NT) {
        float acc = 0.f;
        for (int j = 0; j < nsplit; ++j)
            acc = fmaf(__ldcg(a.part_o + ((size_t)h * a.max_splits + j) * a.D + d), expf(__ldcg(ml + 2 * j) - M), acc);
        const float o = acc / Lsum;
        a.out[h * a.D + d] = o;
        out_am = fmaxf(out_am, fabsf(o));
    }
    return out_am;
}

template <int NT>
__device__ __forceinline__ float mega_attention_lean(const AttnArgs& a, const AttnStep g, const int* spt, int t, unsigned int* head_cnt, float* sm, long long* ts) {
    float out_am = 0.f;
    const int items = a.heads * g.nsplit;
    int* flag = reinterpret_cast<int*>(sm + attn_lean_scratch_floats(NT) - 4);
    for (int i = blockIdx.x; i < items; i += gridDim.x) {
        const int h = g.nsplit == 1 ? i : i / g.nsplit, j = i - h * g.nsplit;
        const int t0 = j * g.chunk, t1 = min(t, t0 + g.chunk);
        if (i != (int)blockIdx.x) bar_sync(1, NT);   // the scratch is reused
        if (g.nsplit == 1) {   // one item per head: normalised output, no partials / counter / merge
            out_am = fmaxf(out_am, attn_item_lean<NT>(a, spt, h, 0, 0, t, sm, true, ts));
            continue;
        }
        (void)attn_item_lean<NT>(a, spt, h, j, t0, t1, sm, false, ts);
        // Every thread's partial stores precede this CTA barrier; ONE acquire-release atomic by thread 0 then publishes them
        // (release, cumulative over what the barrier ordered before it) and, for the CTA that arrives last on the head's
        // counter, makes the other splits' partials visible (acquire) -- no separate fences.
        bar_sync(1, NT);
        if (threadIdx.x == 0) {
            unsigned int prev;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&head_cnt[h]) : "memory");
            *flag = (prev == (unsigned int)(g.nsplit - 1)) ? 1 : 0;
            if (*flag) head_cnt[h] = 0u;  // ready for the next layer (a grid barrier away)
        }
        bar_sync(1, NT);
        if (ts) ts[9] = clock64();
        if (*flag) out_am = fmaxf(out_am, attn_merge_head<NT>(a, h, g.nsplit));
        if (ts) ts[10] = clock64();
    }
    return out_am;
}

// stand-alone launches (TensorEngine::attention_fast_incremental / multi_head_attention entry points, per-op engine)
__device__ __forceinline__ AttnArgs attn_for_sequence(AttnArgs a, int z) {
    a.q += (size_t)z * a.zq;
    a.out += (size_t)z * a.zout;
    a.page_table += (size_t)z * a.ztable;
    a.part_o += z * a.zpart_o;
    a.part_ml += z * a.zpart_ml;
    return a;
}
__global__ void __launch_bounds__(kAttnThreads) attn_partial_kernel(const AttnArgs a0) {
    extern __shared__ float attn_dyn_smem[];
    pdl_wait_prior_grid();       // (no-op unless launched with programmatic stream serialization: the lockstep step's graph)
    pdl_launch_dependents();
    const AttnArgs a = attn_for_sequence(a0, blockIdx.z);
    const int t = *a.pos_ptr + a.t_bias;
    int nsplit, chunk;
    attn_split_range(t, a.max_splits, a.min_chunk, nsplit, chunk);
    const int j = blockIdx.y;
    if (j >= nsplit) return;
    if (a.D <= 128) (void)attn_item_fast<kAttnThreads>(a, blockIdx.x, j, j * chunk, min(t, (j + 1) * chunk), attn_dyn_smem, nsplit == 1);
    else (void)attn_item<kAttnThreads>(a, blockIdx.x, j, j * chunk, min(t, (j + 1) * chunk), attn_dyn_smem, nsplit == 1);
}
__global__ void __launch_bounds__(kAttnThreads) attn_combine_kernel(const AttnArgs a0) {
    pdl_wait_prior_grid();       // (no-op unless launched with programmatic stream serialization: the lockstep step's graph)
    pdl_launch_dependents();
    const AttnArgs a = attn_for_sequence(a0, blockIdx.z);
    const int t = *a.pos_ptr + a.t_bias;
    int nsplit, chunk;
    attn_split_range(t, a.max_splits, a.min_chunk, nsplit, chunk);
    if (nsplit == 1) return;  // attn_partial_kernel wrote the output directly
    (void)attn_merge_head<kAttnThreads>(a, blockIdx.x, nsplit);
}

// the attention phase of the persistent kernel: (head, split) items dealt round-robin to the CTAs; the CTA that
// completes the last split of a head merges that head's partials (no extra grid barrier, deterministic result)
__device__ __forceinline__ float mega_attention(const AttnArgs& a, int t, unsigned int* head_cnt, float* sm, long long* ts = nullptr) {
    constexpr int NT = kConsumerThreads;
    float out_am = 0.f;
    int nsplit, chunk;
    attn_split_range(t, a.max_splits, a.min_chunk, nsplit, chunk);
    const int items = a.heads * nsplit;
    int* flag = reinterpret_cast<int*>(sm + attn_scratch_floats(a.D, NT) - 4);
    for (int i = blockIdx.x; i < items; i += gridDim.x) {
        const int h = i / nsplit, j = i - h * nsplit;
        const int t0 = j * chunk, t1 = min(t, t0 + chunk);
        const bool fast = a.D <= 128;
        if (nsplit == 1) {  // short context: one CTA per head does everything, no partials / counters / merge
            out_am = fmaxf(out_am, fast ? attn_item_fast<NT>(a, h, 0, 0, t, sm, true, ts) : attn_item<NT>(a, h, 0, 0, t, sm, true));
            bar_sync(1, NT);
            continue;
        }
        if (fast) (void)attn_item_fast<NT>(a, h, j, t0, t1, sm, false, ts);
        else { (void)attn_item<NT>(a, h, j, t0, t1, sm); bar_sync(1, NT); }
        // Every thread's partial stores precede the CTA barrier that ended the item; ONE acquire-release atomic by thread 0
        // then publishes them (release, cumulative over what the barrier ordered before it) and, for the CTA that arrives last on the
        // head's counter, makes the other splits' partials visible (acquire) -- no separate fences.
        if (threadIdx.x == 0) {
            unsigned int prev;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&head_cnt[h]) : "memory");
            *flag = (prev == (unsigned int)(nsplit - 1)) ? 1 : 0;
            if (*flag) head_cnt[h] = 0u;  // ready for the next layer (a grid barrier away)
        }
        bar_sync(1, NT);
        if (ts) ts[9] = clock64();
        if (*flag) out_am = fmaxf(out_am, attn_merge_head<NT>(a, h, nsplit));
        bar_sync(1, NT);
        if (ts) ts[10] = clock64();
    }
    return out_am;
}

// Registers: more than 16 warps put 5 on an SM sub-partition (16 K registers each), which caps a uniform allocation at 96
// per thread -- 18 warps at 112 do not fit, measured.  The kernel is compiled for 96 (__maxnreg__); the producer warpgroup then
// shrinks to 40 and the four consumer warpgroups grow to 104 (the pool is the CTA's launch allocation: 640 x 96 >= 512 x 104
// + 128 x 40).  setmaxnreg is a warpgroup-wide instruction, so the producer warp comes with three siblings: warp 17 is the
// grid-barrier warp, warps 18 and 19 only execute the shrink and exit: the CTA has 20 warps.
constexpr int kMegaThreads = (kConsumerWarps + 4) * 32;  // 640
#ifndef TIB_PROD_REGS
#define TIB_PROD_REGS 32
#define TIB_CONS_REGS 112
#endif
// TL: the debug-timeline instance (SM-clock stamps along every phase, m.dbg); the production instance carries none of it
template <int BITS, bool TL = false>
__global__ void __maxnreg__(96) mega_decode_kernel(const __grid_constant__ MegaArgs m) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // carve: mbarriers | ring | x digits | column sums | reduction scratch | attention scratch | phase descriptors -- plain
    // pointer arithmetic on the shared array (every block is a multiple of 16 bytes), so that the accesses stay LDS / STS
    uint8_t* tail = nullptr;
    const GemvSmem sm = gemv_carve_for(smem_raw, m.stages, m.max_kpad, m.max_units, &tail);
    float* attn_sm = reinterpret_cast<float*>(tail);
    uint8_t* const ptail = tail + (((size_t)m.attn_floats * 4 + 15) & ~size_t(15));
    // the phase descriptors, staged by the consumers ONE PHASE AHEAD (three slots): while a phase runs out of slot i, the next
    // one is copied into slot i + 1 before the grid barrier, so that a phase's own header -- slab geometry, epilogue constants,
    // norm weights, all of it on the barrier's critical path -- reads its descriptor from shared memory instead of starting
    // with a chain of global loads; slot i + 2 may still be read by a warp finishing the previous phase's epilogue
    constexpr size_t kPhaseSlot = (sizeof(MegaPhase) + 15) & ~size_t(15);
    uint8_t* const sph_base = ptail;
    // the KV page table of the sequence, copied once: a page lookup in the attention phase is then an LDS
    int* const spt = (m.kv_pages > 0 && m.kv_pages <= kMaxSmemPages) ? reinterpret_cast<int*>(ptail + 3 * kPhaseSlot) : nullptr;
    if (tid == 0) gemv_init_barriers(sm, m.stages);
    __syncthreads();

    const int pos0 = m.st->pos;
    const int step0 = m.st->step;   // read once: thread 0 bumps st->step at the end of the launch while other warps may still publish
    RingPos it;

    if (warp >= kConsumerWarps) {
        // ===== producer warpgroup: warp 16 streams every GEMV phase of every step, back to back =====
        reg_dealloc<TIB_PROD_REGS>();
        if (warp == kConsumerWarps) {
            // lane l holds the record of phase base + l: one latency per 32 phases, then register shuffles only
            for (int s = 0; s < m.n_steps; ++s) {
                const bool sample = s >= m.first_sample;
                for (int base = 0; base < m.nphases; base += 32) {
                    const int lph = base + lane;
                    uint32_t cur[10];
                    {
                        const uint32_t* src = reinterpret_cast<const uint32_t*>(m.prod + (lph < m.nphases ? lph : 0));
#pragma unroll
                        for (int i = 0; i < 10; ++i) cur[i] = __ldg(src + i);
                        if (lph >= m.nphases) cur[9] = 0;
                    }
                    const int cnt = min(32, m.nphases - base);
                    for (int j = 0; j < cnt; ++j) {
                        const int flags = (int)__shfl_sync(0xffffffffu, cur[9], j);
                        if (!(flags & 1) || ((flags & 2) && !sample)) continue;
                        GemvArgs g;
                        QLayout& L = g.L;
                        L.K = (int)__shfl_sync(0xffffffffu, cur[2], j);
                        L.N = (int)__shfl_sync(0xffffffffu, cur[3], j);
                        L.bits = (int)__shfl_sync(0xffffffffu, cur[4], j);
                        L.kc = (int)__shfl_sync(0xffffffffu, cur[5], j);
                        L.nchunks = (int)__shfl_sync(0xffffffffu, cur[6], j);
                        L.U = (int)__shfl_sync(0xffffffffu, cur[7], j);
                        L.P = (int)__shfl_sync(0xffffffffu, cur[8], j);
                        if ((int)blockIdx.x >= L.P) continue;
                        const uint32_t lo = __shfl_sync(0xffffffffu, cur[0], j), hi = __shfl_sync(0xffffffffu, cur[1], j);
                        g.wq = reinterpret_cast<const uint8_t*>(((unsigned long long)hi << 32) | lo);
                        g.stages = m.stages;
                        const Slab slab = make_slab(L, blockIdx.x);
                        gemv_produce(g, slab, sm, it, lane);
                    }
                }
            }
        }
        // ===== warp 17: the grid barrier =====
        // Arriving (a gpu-scope release = MEMBAR, ~0.6 us) and polling are a serial chain of L2 round trips.  A consumer
        // thread doing them would hold its warp's share of the next phase's pre-barrier work behind that chain; this
        // warp has nothing else to do.  Consumers signal "phase done" with bar.arrive 2 and, after bar.sync 3, pick up from
        // shared memory the conversion scalars this warp derived from the statistics of the phase's output.
        if (warp == kConsumerWarps + 1) {
            unsigned int k = 0;                                  // local barriers of this launch
            unsigned int kk = m.tp > 1 ? *m.mg_seq : 0u;         // multi-GPU barriers since the group was formed
            for (int s = 0; s < m.n_steps; ++s) {
                const int nph = m.nphases - (s >= m.first_sample ? 0 : m.n_head);   // the lm_head (and the key exchange) run on sampling steps only
                for (int ph = 0; ph < nph; ++ph) {
                    // what the NEXT phase's prologue needs to turn the statistics into its conversion scalars (fetched
                    // while the consumers are still working on this phase)
                    int nK = 0;
                    bool nnorm = false;
                    float neps = 0.f;
                    if (lane == 0 && ph + 1 < nph) {
                        const MegaPhase* np = m.phases + ph + 1;
                        nK = __ldg(&np->g.L.K);
                        nnorm = __ldg(reinterpret_cast<const unsigned long long*>(&np->g.norm_w)) != 0ull;
                        neps = __ldg(&np->g.rms_eps);
                    }
                    const bool mgpu = m.tp > 1 && __ldg(&m.phases[ph].mgpu) == 1;
                    bar_sync(2, kConsumerThreads + 32);   // every consumer thread has stored its outputs and partials
                    // debug timeline ([25] CTA done, [26] barrier passed: SM clock; [27], [28] the same on the global timer,
                    // comparable across CTAs): how long each CTA waits for the slowest one
                    long long* bts = (TL && m.dbg != nullptr && s == 0 && lane == 0 && (blockIdx.x == 0 || (m.dbg_flags & 4)))
                                         ? m.dbg + ((size_t)blockIdx.x * m.nphases + ph) * kStampsPerPhase : nullptr;
                    if (bts) { bts[25] = clock64(); bts[27] = (long long)globaltimer_ns(); }
                    if (mgpu) {
                        // barrier across all GPUs: the partial outputs this CTA wrote into the peers' buffers become visible
                        // (system-scope release) before its arrival is counted on every rank's word
                        unsigned int* wl = m.peer_bar[m.tp_rank] + (kk & (kBarWords - 1)) * kBarStride;
                        if (lane < m.tp) bar_arrive_sys(m.peer_bar[lane] + (kk & (kBarWords - 1)) * kBarStride);
                        if (lane == 0) {
                            const long long t0 = clock64();
                            const unsigned int target = gridDim.x * (unsigned int)m.tp;
                            while (bar_poll_sys(wl) < target) {
                                if (clock64() - t0 > 120000000000LL) __trap();   // ~60 s: ranks may start a launch at different times
                            }
                            if (blockIdx.x == 0) m.peer_bar[m.tp_rank][((kk + kBarWords - 1) & (kBarWords - 1)) * kBarStride] = 0u;
                        }
                        ++kk;
                        __syncwarp();
                        bar_arrive(3, kConsumerThreads + 32);
                        continue;
                    }
                    if (lane == 0) {
                        float ss = 0.f, am = 0.f;
#pragma unroll
                        for (int i = 0; i < kConsumerWarps; ++i) { ss += sm.red[32 + i]; am = fmaxf(am, sm.red[48 + i]); }
                        unsigned int* w = m.grid_bar + (k & (kBarWords - 1)) * kBarStride;
                        bar_arrive_stats(w, ss, am);
                        const long long t0 = clock64();
                        uint4 v = bar_poll(w);
                        while (v.x < gridDim.x) {
                            if (clock64() - t0 > kBarWatchdog) __trap();
                            v = bar_poll(w);
                        }
                        XStats st;
                        st.ss = __ull2float_rn(((unsigned long long)v.w << 32) | v.z) * (1.0f / kSsScale);
                        st.am = __uint_as_float(v.y);
                        const XScale sc = make_xscale(st, nnorm, nK > 0 ? nK : 1, neps);
                        sm.red[0] = sc.inv_rms;
                        sm.red[1] = sc.inv_s;
                        sm.red[2] = sc.s_x;
                        if (blockIdx.x == 0)
                            *reinterpret_cast<uint4*>(m.grid_bar + ((k + kBarWords - 1) & (kBarWords - 1)) * kBarStride) = make_uint4(0u, 0u, 0u, 0u);
                        if (bts) { bts[26] = clock64(); bts[28] = (long long)globaltimer_ns(); }
                    }
                    ++k;
                    __syncwarp();
                    bar_arrive(3, kConsumerThreads + 32);  // barrier passed, scalars in sm.red[0..2]
                }
            }
            // the point-to-point variant keeps its own exchange count in the same word (written by consumer thread 0 at the
            // end of the launch); this warp passes no barrier across the GPUs in that mode and must not write a stale value back
            if (m.tp > 1 && blockIdx.x == 0 && lane == 0) m.mg_seq[0] = kk;   // (the exchange count lives in its own word, mg_seq[1])
        }
        return;
    }

    // ===== consumers =====
    reg_alloc<TIB_CONS_REGS>();
    for (int i = tid; i < m.max_units * 4 * 3; i += kConsumerThreads) sm.acc[i] = 0;   // from here on every epilogue re-zeroes what it reads
    if (spt != nullptr)
        for (int i = tid; i < m.kv_pages; i += kConsumerThreads) spt[i] = m.kv_page_table[i];
    // ends a phase: publishes this CTA's partial statistics of the phase's output, then arrives on the grid barrier
    bool need_wait = false;
    // ends a phase: leaves this CTA's partial statistics of the phase's output for the barrier warp and signals it
    auto grid_arrive = [&](XStats st, long long* ts) {
        st.ss = warp_sum(st.ss);
        st.am = warp_max(st.am);
        if (lane == 0) { sm.red[32 + warp] = st.ss; sm.red[48 + warp] = st.am; }
        __syncwarp();
        bar_arrive(2, kConsumerThreads + 32);   // does not wait: on to the next phase's pre-barrier work
        if (TL && ts) ts[21] = clock64();
        need_wait = true;
    };
    // waits until every CTA has finished the phase last arrived on; the next prologue's scalars are then in sm.red[0..2]
    auto grid_wait = [&]() {
        if (!need_wait) return;
        bar_sync(3, kConsumerThreads + 32);
        need_wait = false;
    };
    auto decode_key = [&](int s) -> int {
        unsigned long long key = __ldcg(&m.keys[s & 1]);
        if (m.n_head == 2) {   // vocabulary sharded over the tensor-parallel ranks: the maximum of every rank's key
            key = 0ull;
            for (int r = 0; r < m.tp; ++r) {
                const unsigned long long kr = __ldcg(m.peer_keys[m.tp_rank] + (size_t)(s & 1) * kMaxTp + r);
                key = kr > key ? kr : key;
            }
        }
        // key 0: no logit compared greater than -inf (all NaN / -inf): defined as token 0; the clamp keeps the embedding
        // lookup of the next step inside the table whatever the key holds
        const int tok = key == 0ull ? 0 : 0x7FFFFFFF - (int)(uint32_t)(key & 0xFFFFFFFFull);
        return min(max(tok, 0), m.V - 1);
    };
    // CTA 0 publishes the token picked in step s (called by all its consumer threads, after that step's barrier)
    auto publish = [&](int s, int tok) {
        if (blockIdx.x != 0) return;
        const int k = step0 + (s - m.first_sample);
        if (m.io->hist && k < m.io->hist_cap)
            for (int i = tid; i < m.V; i += kConsumerThreads) m.io->hist[(size_t)k * m.V + i] = __ldcg(m.logits + i);
        if (tid == 0) {
            if (m.io->out_tokens && k < m.io->out_cap) m.io->out_tokens[k] = tok;
            m.keys[(s + 1) & 1] = 0ull;
        }
    };

    unsigned int phase_seq = 0;   // phases executed so far: selects the descriptor slot
    auto stage_descriptor = [&](int ph, unsigned int slot) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(m.phases + ph);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sph_base + (size_t)slot * kPhaseSlot);
        for (int i = tid; i < (int)(sizeof(MegaPhase) / 4); i += kConsumerThreads) dst[i] = src[i];
    };
    stage_descriptor(0, 0);   // the first phase's descriptor; every later one is staged during its predecessor
    bar_sync(1, kConsumerThreads);
    unsigned int ex_seq = (m.tp > 1 && m.tp_p2p) ? m.mg_seq[1] : 0u;   // point-to-point exchanges since the group was formed
    int token = m.st->token;  // decode-only launches continue from the token the previous launch picked
    for (int s = 0; s < m.n_steps; ++s) {
        const bool sample = s >= m.first_sample;
        const int pos = pos0 + s;
        AttnStep ageo{1, 1};        // split geometry of this step's attention phases (the same in every layer)
        bool ageo_set = false;
        if (s > 0 && s - 1 >= m.first_sample) {
            grid_wait();
            token = decode_key(s - 1);
            publish(s - 1, token);
        }
        if (s < m.n_prompt) token = m.prompt[s];
        for (int ph = 0; ph < m.nphases; ++ph) {
            const bool is_head = ph == m.nphases - m.n_head;  // the lm_head is the first of the n_head phases at the end of the list
            if (ph >= m.nphases - m.n_head && !sample) continue;
            // debug timeline: thread 0 of CTA 0 (dbg_flags & 4: of every CTA) stamps the SM clock along the phase
            const bool stamp = TL && m.dbg != nullptr && s == 0 && tid == 0 && (blockIdx.x == 0 || (m.dbg_flags & 4));
            long long* ts = m.dbg + ((size_t)blockIdx.x * m.nphases + ph) * kStampsPerPhase;
            if (stamp) ts[0] = clock64();
            // Everything that does not depend on the previous phase's output happens BEFORE the grid barrier:
            // stage the phase descriptor in shared memory, fetch the epilogue's per-column constants.
            const unsigned int slot = phase_seq % 3u;
            ++phase_seq;
            MegaPhase* const sph = reinterpret_cast<MegaPhase*>(sph_base + (size_t)slot * kPhaseSlot);
            const MegaPhase& PG = *sph;   // staged during the previous phase (or before the loop)
            {   // the phase that will execute next (the head phases only on sampling steps; after the last one: phase 0 of the next step)
                int nx = ph + 1;
                if (nx >= m.nphases - m.n_head && !sample) nx = m.nphases;
                if (nx >= m.nphases) nx = 0;
                stage_descriptor(nx, phase_seq % 3u);
            }
            const bool gemv_here = PG.type == PH_GEMV && (int)blockIdx.x < PG.g.L.P;
            const bool to_peers = m.tp > 1 && PG.type == PH_GEMV && PG.mgpu != 0;
            // (a point-to-point phase bumps ex_seq in its tail, after the epilogue: the epilogue tags its stores with the NEXT number)
            const PhaseCtx ctx{true, pos, is_head ? &m.keys[s & 1] : nullptr, to_peers ? m.peer_part[PG.part_sel] : nullptr, to_peers ? m.tp : 0,
                               (size_t)m.tp_rank * (size_t)m.H, (to_peers && m.tp_ll && PG.mgpu == 2) ? ex_seq + 1u : 0u};
            Slab slab{};
            EpiPre pre{};
            XPre xpre;
            ConsumePlan plan{};
            const float* resid = nullptr;
            if (gemv_here) {
                slab = make_slab(PG.g.L, blockIdx.x);
                resid = PG.resid_src == SRC_EMB ? m.emb + (size_t)token * m.H : PG.g.resid;
                pre = gemv_epilogue_prefetch(PG.g, slab, resid, ctx, tid);
                if (!(m.dbg_flags & 1)) gemv_x_prefetch(PG.g, tid, xpre);
                plan = make_consume_plan(PG.g.L, slab, warp, lane);
            }
            if (stamp) ts[12] = clock64();
            if (need_wait) grid_wait(); else bar_sync(1, kConsumerThreads);
            const MegaPhase& P = *sph;
            if (stamp) { ts[1] = clock64(); ts[2] = ts[1]; ts[3] = ts[1]; }
            XStats out_st{0.f, 0.f};
            if (P.type == PH_GEMV) {
                if (gemv_here) {
                    const bool from_emb = P.x_src == SRC_EMB;
                    const float* x = from_emb ? m.emb + (size_t)token * m.H : P.g.x;
                    const GemvArgs& g = P.g;
                    if (stamp) ts[6] = clock64();
                    // conversion scalars: from the barrier warp (it derived them from the statistics of the phase that
                    // produced x), or from the precomputed statistics of the embedding row
                    const XScale xsc = from_emb ? make_xscale(m.emb_stats[token], g.norm_w != nullptr, g.L.K, g.rms_eps)
                                                : XScale{sm.red[0], sm.red[1], sm.red[2]};
                    // the first ring stage of this phase has usually landed long ago: probe its barrier now, so that the
                    // poll's latency hides behind the conversion instead of opening the main loop
                    const bool ready0 = slab.rounds > 0 && mbar_test_wait(&sm.full[it.st], it.par);
                    const float s_x = gemv_stage_x_lean<BITS>(g, x, sm, slab, !from_emb, xsc, xpre, (m.dbg_flags & 1) != 0, tid, lane, stamp ? ts + 13 : nullptr);
                    if (stamp) ts[2] = clock64();
                    if constexpr (TL) {   // every thread runs the same instance (the loop holds warp-collective instructions)
                        gemv_consume<BITS, 3>(g, slab, sm, it, plan, warp, lane, stamp ? ts + 17 : nullptr, ready0, m.dbg_flags & 24);
                    } else {
                        gemv_consume<BITS, 0>(g, slab, sm, it, plan, warp, lane, nullptr, ready0);
                    }
                    if (stamp) ts[3] = clock64();
                    out_st = gemv_epilogue(g, slab, sm, s_x, resid, ctx, pre, tid, lane);
                }
                if (m.tp > 1 && P.mgpu == 2) {
                    // Point-to-point all-reduce of this CTA's column slice.  CTA b of every rank computes the same columns, so
                    // only those P CTAs meet: the partials went to every rank's buffer in the epilogue; one flag per (source
                    // rank, CTA) says "slice written" (system-scope release after the CTA-wide barrier), the CTA waits for
                    // its P flags and sums residual + partials in rank order -- every rank reads the same stored values, so
                    // the replicated residual stream stays bit-identical.
                    ++ex_seq;
                    if (!m.tp_ll) {
                    bar_sync(1, kConsumerThreads);   // every thread's peer stores are ordered before the flags
                    if (warp == 0) {
                        if (lane < m.tp) {
                            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(m.peer_flag[P.part_sel][lane] + (size_t)m.tp_rank * gridDim.x + blockIdx.x), "r"(ex_seq)
                                         : "memory");
                            const unsigned int* fl = m.peer_flag[P.part_sel][m.tp_rank] + (size_t)lane * gridDim.x + blockIdx.x;
                            const long long t0 = clock64();
                            while ((int)(bar_poll_sys(fl) - ex_seq) < 0) {
                                if (clock64() - t0 > 120000000000LL) __trap();
                            }
                        }
                        __syncwarp();
                    }
                    bar_sync(1, kConsumerThreads);
                    }
                    out_st = XStats{0.f, 0.f};
                    if (gemv_here) {
                        const float* part = m.peer_part[P.part_sel][m.tp_rank];
                        const int Hn = P.g.L.N;
                        if (m.tp_ll) {
                            // Every (column, rank) word of this CTA's slice is polled by its OWN thread -- ncols * P <= 512 polls in
                            // flight at once, one L2 round trip whatever P is (one thread walking the P words of a column pays P
                            // round trips: 8 x 0.6 us at TP 8) -- and handed over through shared memory (the attention scratch, idle
                            // in a GEMV phase).  The stores of the slice were tagged with this exchange's number by the thread that
                            // owns the column on each rank.
                            const unsigned long long* part8 = reinterpret_cast<const unsigned long long*>(part);
                            const int nw = slab.ncols * m.tp;
                            for (int idx = tid; idx < nw; idx += kConsumerThreads) {
                                const int c = idx / m.tp, r = idx - c * m.tp, n = slab.col0 + c;
                                float val = 0.f;
                                if (n < Hn) {
                                    const unsigned long long* w = part8 + (size_t)r * Hn + n;
                                    unsigned long long word;
                                    const long long t0 = clock64();
                                    while (true) {
                                        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(w) : "memory");
                                        if ((unsigned int)(word >> 32) == ex_seq) break;
                                        if (clock64() - t0 > 120000000000LL) __trap();
                                    }
                                    val = __uint_as_float((unsigned int)word);
                                }
                                attn_sm[idx] = val;
                            }
                            bar_sync(1, kConsumerThreads);
                        }
                        for (int c = tid; c < slab.ncols; c += kConsumerThreads) {
                            const int n = slab.col0 + c;
                            if (n >= Hn) continue;
                            float v = ld_act(resid + n, P.resid_src != SRC_EMB);
                            if (m.tp_ll) {
                                for (int r = 0; r < m.tp; ++r) v += attn_sm[c * m.tp + r];   // in rank order: the same sums on every rank
                            } else {
                                for (int r = 0; r < m.tp; ++r) v += __ldcg(part + (size_t)r * Hn + n);
                            }
                            P.g.out[n] = v;
                            out_st.ss = fmaf(v, v, out_st.ss);
                            out_st.am = fmaxf(out_st.am, fabsf(P.g.next_norm_w ? v * P.g.next_norm_w[n] : v));
                        }
                        if (m.tp_ll) bar_sync(1, kConsumerThreads);   // the scratch is free again before anybody reuses it
                    }
                }
            } else if (P.type == PH_KEYX) {
                // the lm_head's grid barrier has been passed: this rank's key is final; hand it to every rank (the barrier
                // across the GPUs that ends this phase orders the stores before anybody reads them)
                if (blockIdx.x == 0 && tid == 0) {
                    const unsigned long long key = __ldcg(&m.keys[s & 1]);
                    for (int r = 0; r < m.tp; ++r)
                        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(m.peer_keys[r] + (size_t)(s & 1) * kMaxTp + m.tp_rank), "l"(key) : "memory");
                }
            } else if (P.type == PH_REDUCE) {
                // all-reduce, second half: every rank's partial of the row-parallel GEMV is in this rank's buffer (the
                // barrier across the GPUs has been passed); x <- residual + partials in rank order -- the same sums on
                // every rank, so the replicated residual stream stays bit-identical across the group
                const int n = P.g.L.N;
                const int per = ((n + (int)gridDim.x - 1) / (int)gridDim.x + 3) & ~3;
                const int i = (int)blockIdx.x * per + tid;
                if (tid < per && i < n) {
                    const float* rsd = P.resid_src == SRC_EMB ? m.emb + (size_t)token * m.H : P.g.resid;
                    float v = P.resid_src == SRC_EMB ? rsd[i] : __ldcg(rsd + i);
                    for (int r = 0; r < m.tp; ++r) v += __ldcg(P.g.x + (size_t)r * n + i);
                    P.g.out[i] = v;
                    out_st.ss = v * v;
                    out_st.am = fabsf(P.g.next_norm_w ? v * P.g.next_norm_w[i] : v);
                }
            } else {
                if (P.at.D <= 128 && P.at.page_shift >= 0) {   // lean item
                    if (!ageo_set) { ageo = attn_step_geometry(P.at, pos + 1); ageo_set = true; }
                    out_st.am = mega_attention_lean<kConsumerThreads>(P.at, ageo, spt, pos + 1, m.head_cnt, attn_sm, stamp ? ts : nullptr);
                } else {
                    out_st.am = mega_attention(P.at, pos + 1, m.head_cnt, attn_sm, stamp ? ts : nullptr);
                }
            }
            if (stamp) ts[4] = clock64();
            grid_arrive(out_st, stamp ? ts : nullptr);
            if (stamp) ts[5] = clock64();
        }
    }
    // tail: the last step's token, and the state the host reads back
    grid_wait();
    const int last = m.n_steps - 1;
    if (last >= m.first_sample) {
        token = decode_key(last);
        publish(last, token);
    }
    if (blockIdx.x == 0 && tid == 0) {
        if (m.tp > 1 && m.tp_p2p) m.mg_seq[1] = ex_seq;
        m.st->pos = pos0 + m.n_steps;
        m.st->token = token;
        const int sampled = m.n_steps - m.first_sample;
        if (sampled > 0) m.st->step += sampled;
    }
}

}  // namespace tib
