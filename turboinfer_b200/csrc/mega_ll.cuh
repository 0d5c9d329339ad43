// mega_ll.cuh -- the persistent decode kernel as a DATAFLOW pipeline: one cooperative launch runs whole forward passes
// and whole greedy generations, like mega.cuh, but no phase ends in a grid-wide barrier.
//
// What it replaces: InferenceEngine::forward_pass_incremental (src/model/inference_engine.cpp:1493-1552) ->
// TransformerLayer::forward_incremental (:244-401) and the host-side sampling loop of generate (:752-775).
//
// Why: at batch 1 a decoder layer is five dependent steps (QKV, attention, O, gate/up, down).  With a counter barrier
// each hand-over costs  stores -> fence (wait for acks) -> atomic -> poll -> loads  = 3-4 us measured per phase on a
// B200 (profiles/r01_timeline_barrier.txt), 5x32 times per token for the 7B shape -- more than the 0.5 ms the weights
// need at HBM rate.  Here every activation element travels as one 64-bit LL word {value, epoch} (ptx.cuh): producers
// just store, consumers re-read until the epoch matches, so a hand-over costs one L2 round trip.  The epoch is the
// global phase counter, so a stale word can never be mistaken for a fresh one and buffers need no clearing.
//   * the producer warp streams the CTA's weight slabs of ALL phases back to back through the shared-memory ring
//     (weights never depend on activations), so HBM keeps flowing while the consumers wait for their inputs;
//   * once per step every CTA publishes its best (logit, index) key behind a release fence and reads everyone else's
//     behind an acquire fence: that is the greedy token, and the one point that orders the step's plain stores (KV
//     cache rows, logits) before the next step's reads.
// Write-after-read safety needs no extra synchronisation: a CTA reaches the epilogue of phase p+1 only after it has
// received every output of phase p, i.e. after every CTA has finished reading phase p's inputs; no buffer is read by
// one phase's prologue and written by the same phase's epilogue.
#pragma once
#include "mega.cuh"

namespace tib {

struct MegaLLPhase {
    int type;        // PH_GEMV / PH_ATTN
    int x_src;       // SRC_PTR: x_ll, SRC_EMB: embedding row of the current token (plain floats)
    int resid_src;
    int is_head;
    GemvArgs g;      // g.out / g.resid point at llword arrays (except EPI_LOGITS: plain logits)
    AttnArgs at;
    const llword* x_ll;
    llword* knew_ll;
    llword* vnew_ll;
    // who produced x: only CTAs that had work in that phase publish statistics, and only those are waited for (a CTA
    // without work never blocks and may be phases ahead; it must not touch a slot somebody still waits on)
    int prod_kind;       // 0: a GEMV phase with prod_P CTAs; 1: an attention phase (CTAs = min(heads * nsplit(t), grid))
    int prod_P;
    int prod_heads, prod_splits, prod_minchunk;
};

struct MegaLLArgs {
    const MegaLLPhase* __restrict__ phases;
    const ProdRec* __restrict__ prod;
    int nphases;
    const float* emb;     // [V][H]
    int H, V;
    StepState* st;
    const StepIO* io;
    const int* prompt;
    int n_prompt, n_steps, first_sample;
    llword* keys_ll;      // [gridDim.x][2]: per-CTA best key of the step (hi, lo), the step-boundary exchange
    llword* stats_ll;     // [2][gridDim.x][2]: per-CTA partial XStats (ss, am) of a phase's output, by phase parity
    unsigned int* cnt;    // [0], [32]: CTAs that have finished a phase, by phase parity; [64]: CTAs that have finished a step.
                          // Zeroed by the host before every launch.  Incremented with RELAXED atomics and no fence: they
                          // are a hint that tells one polling thread per CTA when re-reading LL words is worth it; every
                          // LL word is still verified by its epoch, so no ordering is needed between data and counter.
    const XStats* emb_stats;  // [V]
    float* logits;
    int stages;
    int max_kpad, max_units, attn_floats;
    long long* dbg;
    uint32_t ep0;         // epochs used by this launch: ep0 + 1 ... ep0 + n_steps * (nphases + 1); kept monotonic by the host
};

TIB_HD int attn_ll_scratch_floats(int D, int nt) {
    const int groups = D < nt ? nt / D : 1;
    return 3 * D + kAttnTokBlock + 32 + groups * D + 4;
}
TIB_HD size_t mega_ll_smem_bytes(int stages, int max_kpad, int max_units, int attn_floats) {
    return gemv_smem_bytes_for(stages, max_kpad, max_units) + 16 + (size_t)attn_floats * 4 + 16 + 2 * ((sizeof(MegaLLPhase) + 15) & ~size_t(15));
}

// (m, l, o) partial of head h over tokens [t0, t1); the current token's K / V rows come from the LL buffers written by
// the QKV epilogue of this step, older rows from the paged cache.  direct: the item covers the whole context ->
// normalise and publish the head's output; otherwise publish the partial.
template <int NT>
__device__ __forceinline__ float attn_item_ll(const AttnArgs& a, int h, int j, int t0, int t1, int cur_tok, uint32_t in_ep, uint32_t out_ep,
                                              float* sm, bool direct) {
    float out_am = 0.f;  // max |value published to out_ll| by this thread
    float* qs = sm;
    float* ks = qs + a.D;
    float* vs = ks + a.D;
    float* sc = vs + a.D;
    float* red = sc + kAttnTokBlock;
    float* ored = red + 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = a.D, hoff = h * D;
    const bool has_cur = cur_tok >= t0 && cur_tok < t1;
    for (int d = tid; d < D; d += NT) {
        qs[d] = ll_wait_f(a.q_ll + hoff + d, in_ep);
        if (has_cur) {
            ks[d] = ll_wait_f(a.knew_ll + hoff + d, in_ep);
            vs[d] = ll_wait_f(a.vnew_ll + hoff + d, in_ep);
        }
    }
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
            const bool cur = tb + tt == cur_tok;
            const float* kr = cur ? ks : kv_row(a.k_pool, a.page_table, a.page_tokens, a.H, tb + tt) + hoff;
            float s = 0.f;
            for (int d = 4 * lane; d < D; d += 128) {
                const float4 kv = cur ? *reinterpret_cast<const float4*>(kr + d) : __ldcg(reinterpret_cast<const float4*>(kr + d));
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
                        const float v = (tb + tt == cur_tok) ? vs[d] : __ldcg(kv_row(a.v_pool, a.page_table, a.page_tokens, a.H, tb + tt) + hoff + d);
                        acc = fmaf(sc[tt], v, acc);
                    }
                o[i] = acc;
            }
        }
        bar_sync(1, NT);
    }
    llword* po = direct ? a.out_ll + hoff : a.part_ll + ((size_t)h * a.max_splits + j) * (D + 2);
    const float onorm = direct ? 1.0f / l_run : 1.0f;
    if (groups > 1) {
        if (active) ored[grp * D + d0] = o[0];
        bar_sync(1, NT);
        if (grp == 0) {
            float acc = 0.f;
            for (int g = 0; g < groups; ++g) acc += ored[g * D + d0];
            ll_store_f(po + d0, acc * onorm, out_ep);
            if (direct) out_am = fabsf(acc * onorm);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kMaxDims; ++i) {
            const int d = d0 + i * NT;
            if (d < D) {
                ll_store_f(po + d, o[i] * onorm, out_ep);
                if (direct) out_am = fmaxf(out_am, fabsf(o[i] * onorm));
            }
        }
    }
    if (tid == 0 && !direct) {
        ll_store_f(po + D, m_run, out_ep);
        ll_store_f(po + D + 1, l_run, out_ep);
    }
    return out_am;
}

// merge the nsplit partials of head h in split order (deterministic) and publish the head's output
template <int NT>
__device__ __forceinline__ float attn_merge_ll(const AttnArgs& a, int h, int nsplit, uint32_t ep) {
    float out_am = 0.f;
    const int D = a.D;
    const llword* base = a.part_ll + (size_t)h * a.max_splits * (D + 2);
    float M = -INFINITY;
    for (int j = 0; j < nsplit; ++j) M = fmaxf(M, ll_wait_f(base + (size_t)j * (D + 2) + D, ep));
    float Lsum = 0.f;
    for (int j = 0; j < nsplit; ++j)
        Lsum += ll_wait_f(base + (size_t)j * (D + 2) + D + 1, ep) * expf(ll_wait_f(base + (size_t)j * (D + 2) + D, ep) - M);
    for (int d = threadIdx.x; d < D; d += NT) {
        float acc = 0.f;
        for (int j = 0; j < nsplit; ++j)
            acc = fmaf(ll_wait_f(base + (size_t)j * (D + 2) + d, ep), expf(ll_wait_f(base + (size_t)j * (D + 2) + D, ep) - M), acc);
        const float o = acc / Lsum;
        ll_store_f(a.out_ll + h * D + d, o, ep);
        out_am = fmaxf(out_am, fabsf(o));
    }
    return out_am;
}

__device__ __forceinline__ float mega_attention_ll(const AttnArgs& a, int t, uint32_t in_ep, uint32_t out_ep, float* sm) {
    constexpr int NT = kConsumerThreads;
    float out_am = 0.f;
    int nsplit, chunk;
    attn_split_range(t, a.max_splits, a.min_chunk, nsplit, chunk);
    const int items = a.heads * nsplit;
    // all of this CTA's partials first, then its merges: a CTA never waits before it has produced everything it owes
    for (int i = blockIdx.x; i < items; i += gridDim.x) {
        const int h = i / nsplit, j = i - h * nsplit;
        const int t0 = j * chunk, t1 = min(t, t0 + chunk);
        out_am = fmaxf(out_am, attn_item_ll<NT>(a, h, j, t0, t1, t - 1, in_ep, out_ep, sm, nsplit == 1));
        bar_sync(1, NT);
    }
    if (nsplit > 1)
        for (int i = blockIdx.x; i < items; i += gridDim.x)
            if (i % nsplit == 0) out_am = fmaxf(out_am, attn_merge_ll<NT>(a, i / nsplit, nsplit, out_ep));
    return out_am;
}

template <int BITS>
__global__ void __maxnreg__(96) mega_ll_kernel(const __grid_constant__ MegaLLArgs m) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* tail = nullptr;
    const GemvSmem sm = gemv_carve_for(smem_raw, m.stages, m.max_kpad, m.max_units, &tail);
    uintptr_t p = (reinterpret_cast<uintptr_t>(tail) + 15) & ~uintptr_t(15);
    float* attn_sm = reinterpret_cast<float*>(p);
    p += (size_t)m.attn_floats * 4;
    p = (p + 15) & ~uintptr_t(15);
    // two descriptor slots, alternating: a fast warp stages the next phase while slow ones still read this one
    MegaLLPhase* sph2 = reinterpret_cast<MegaLLPhase*>(p);
    constexpr size_t kDescStride = (sizeof(MegaLLPhase) + 15) & ~size_t(15);
    if (tid == 0) gemv_init_barriers(sm, m.stages);
    __syncthreads();

    const int pos0 = m.st->pos;
    uint32_t it = 0;

    if (warp >= kConsumerWarps) {
        // ===== producer warpgroup: warp 16 streams every GEMV phase of every step, back to back =====
        reg_dealloc<40>();
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
        return;
    }

    // ===== consumers =====
    reg_alloc<104>();
    constexpr int kDescWords = (int)(sizeof(MegaLLPhase) / 4);
    static_assert(kDescWords <= kConsumerThreads, "one descriptor word per consumer thread");
    const uint32_t ep0 = m.ep0;
    const uint32_t per_step = (uint32_t)m.nphases + 1;
    // the descriptor of the next phase travels global -> register one phase ahead, register -> shared at the phase start
    uint32_t desc_word = 0;
    auto desc_prefetch = [&](int ph) {
        if (tid < kDescWords) desc_word = __ldg(reinterpret_cast<const uint32_t*>(m.phases + (ph < m.nphases ? ph : 0)) + tid);
    };
    desc_prefetch(0);

    // one thread per CTA waits for a hint counter; everybody else sleeps in the barrier that follows
    auto wait_count = [&](const unsigned int* c, unsigned int target) {
        if (tid == 0) {
            unsigned int v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if (v < target) {
                const long long t0 = clock64();
                do {
                    __nanosleep(20);
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
                    if (clock64() - t0 > 8000000000LL) __trap();
                } while (v < target);
            }
        }
        bar_sync(1, kConsumerThreads);
    };
    auto bump = [&](unsigned int* c) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory"); };

    // step boundary: every CTA publishes its best key of the step (0 when the step does not sample) behind a release
    // fence, then reads all of them behind an acquire fence.  Returns the greedy token (valid when the step sampled).
    auto step_exchange = [&](int s, unsigned long long my_key) -> int {
        const uint32_t ep = ep0 + (uint32_t)s * per_step + per_step;
        bar_sync(1, kConsumerThreads);   // all of this CTA's stores of the step are issued
        if (tid == 0) {
            fence_acq_rel_gpu();
            ll_store(m.keys_ll + 2 * blockIdx.x, (uint32_t)(my_key >> 32), ep);
            ll_store(m.keys_ll + 2 * blockIdx.x + 1, (uint32_t)my_key, ep);
            bump(m.cnt + 64);
        }
        wait_count(m.cnt + 64, (unsigned int)(s + 1) * gridDim.x);
        if (warp == 0) {
            constexpr int kMaxPer = 8;
            llword a[kMaxPer], b[kMaxPer];
#pragma unroll
            for (int j = 0; j < kMaxPer; ++j) {
                const int i = lane + 32 * j;
                if (i < (int)gridDim.x) ll_load2(m.keys_ll + 2 * i, a[j], b[j]);
            }
            unsigned long long best = 0ull;
#pragma unroll
            for (int j = 0; j < kMaxPer; ++j) {
                const int i = lane + 32 * j;
                if (i < (int)gridDim.x) {
                    uint32_t hi = (uint32_t)a[j], lo = (uint32_t)b[j];
                    if (!(ll_ready(a[j], ep) && ll_ready(b[j], ep))) { hi = ll_wait(m.keys_ll + 2 * i, ep); lo = ll_wait(m.keys_ll + 2 * i + 1, ep); }
                    const unsigned long long k = ((unsigned long long)hi << 32) | lo;
                    best = k > best ? k : best;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            fence_acq_rel_gpu();
            if (lane == 0) sm.keyred[0] = best;
        }
        bar_sync(1, kConsumerThreads);
        const unsigned long long best = sm.keyred[0];
        bar_sync(1, kConsumerThreads);   // keyred is reused by the next lm_head epilogue
        return 0x7FFFFFFF - (int)(uint32_t)(best & 0xFFFFFFFFull);
    };
    auto publish = [&](int s, int tok) {
        if (blockIdx.x != 0) return;
        const int k = m.st->step + (s - m.first_sample);
        if (m.io->hist && k < m.io->hist_cap)
            for (int i = tid; i < m.V; i += kConsumerThreads) m.io->hist[(size_t)k * m.V + i] = __ldcg(m.logits + i);
        if (tid == 0 && m.io->out_tokens && k < m.io->out_cap) m.io->out_tokens[k] = tok;
    };

    // per-CTA partial statistics of the phase that just ended (gemv.cuh XStats): every warp leaves its part in shared
    // memory (area alternating with the phase); after the NEXT phase's first barrier warp 0 adds them up, publishes
    // two LL words and bumps the phase counter.
    auto stats_deposit = [&](XStats st, int parity) {
        st.ss = warp_sum(st.ss);
        st.am = warp_max(st.am);
        if (lane == 0) { sm.red[32 + 32 * parity + warp] = st.ss; sm.red[48 + 32 * parity + warp] = st.am; }
    };
    auto stats_publish = [&](uint32_t ep, int parity) {   // warp 0
        float ss = 0.f, am = 0.f;
#pragma unroll
        for (int i = 0; i < kConsumerWarps; ++i) { ss += sm.red[32 + 32 * parity + i]; am = fmaxf(am, sm.red[48 + 32 * parity + i]); }
        llword* slot = m.stats_ll + ((size_t)parity * gridDim.x + blockIdx.x) * 2;
        if (lane == 0) ll_store_f(slot, ss, ep);
        if (lane == 1) ll_store_f(slot + 1, am, ep);
        __syncwarp();
        if (lane == 0) bump(m.cnt + 32 * parity);
    };

    // GEMV prologue of the dataflow engine.  The phase counter says the inputs are (almost certainly) there: every thread
    // issues all its LL loads at once, warp 0 also the producers' statistics; whatever is not there yet is re-read.
    auto prologue_ll = [&](const GemvArgs& g, const Slab& slab, const llword* x_ll, uint32_t ep, int spar, int nprod, long long* dbg) -> float {
        const int K = g.L.K, nvec = layout_kpad(g.L) >> 2, kvec = K >> 2;
        const float* nw = g.norm_w;
        for (int i = tid; i < slab.ncols * 3; i += kConsumerThreads) sm.acc[i] = 0;
        float4 xv[kXCache];
        bool ok[kXCache];
#pragma unroll
        for (int i = 0; i < kXCache; ++i) {
            const int v = tid + i * kConsumerThreads;
            ok[i] = true;
            xv[i] = v < kvec ? ll_try4(x_ll + 4 * v, ep, ok[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (warp == 0) {
            const llword* base = m.stats_ll + (size_t)spar * gridDim.x * 2;
            constexpr int kMaxPer = 8;
            llword a[kMaxPer], b[kMaxPer];
#pragma unroll
            for (int j = 0; j < kMaxPer; ++j) {
                const int i = lane + 32 * j;
                if (i < nprod) ll_load2(base + 2 * i, a[j], b[j]);
            }
            float ss = 0.f, am = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxPer; ++j) {
                const int i = lane + 32 * j;
                if (i < nprod) {
                    float vs = ll_value(a[j]), va = ll_value(b[j]);
                    if (!(ll_ready(a[j], ep) && ll_ready(b[j], ep))) { vs = ll_wait_f(base + 2 * i, ep); va = ll_wait_f(base + 2 * i + 1, ep); }
                    ss += vs;          // lane-strided partial sums, then a fixed shuffle tree: the same order in every CTA
                    am = fmaxf(am, va);
                }
            }
            ss = warp_sum(ss);
            am = warp_max(am);
            if (lane == 0) { sm.red[0] = ss; sm.red[1] = am; }
        }
#pragma unroll
        for (int i = 0; i < kXCache; ++i) {
            const int v = tid + i * kConsumerThreads;
            if (v < kvec && !ok[i]) xv[i] = ll_wait4(x_ll + 4 * v, ep);
        }
        if (dbg) dbg[0] = clock64();
        bar_sync(1, kConsumerThreads);   // B2: statistics in shared memory
        if (dbg) dbg[1] = clock64();
        float inv_rms = 1.f, amax = sm.red[1];
        if (nw != nullptr) {
            inv_rms = rsqrtf(sm.red[0] / (float)K + g.rms_eps);   // :1501
            amax = amax * inv_rms * 1.00001f;                      // a bound of max|y|, y = x * inv_rms * w
        }
        const bool finite = amax > 0.f && amax < INFINITY;
        const float inv_s = finite ? __fdividef(kXQMax, amax) : 0.f;
        const float s_x = finite ? amax * (1.0f / kXQMax) : 0.f;
        long long sxf = 0;
        auto emit = [&](int v, float4 t) {
            if (nw != nullptr) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(nw) + v);
                t.x = (t.x * inv_rms) * w.x;   // the reference divides by rms (:1504-1506); the reciprocal differs by ~1 ulp,
                t.y = (t.y * inv_rms) * w.y;   // far below the 2^-24 max|y| granularity of the fixed-point conversion
                t.z = (t.z * inv_rms) * w.z;
                t.w = (t.w * inv_rms) * w.w;
            }
            x_store_digits<BITS>(sm.xd, v, t, inv_s, sxf);
        };
#pragma unroll
        for (int i = 0; i < kXCache; ++i) {
            const int v = tid + i * kConsumerThreads;
            if (v < nvec) emit(v, xv[i]);
        }
        for (int v = tid + kXCache * kConsumerThreads; v < nvec; v += kConsumerThreads)   // K > 16384
            emit(v, v < kvec ? ll_wait4(x_ll + 4 * v, ep) : make_float4(0.f, 0.f, 0.f, 0.f));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sxf += __shfl_xor_sync(0xffffffffu, sxf, o);
        if (lane == 0) sm.sxf[warp] = sxf;
        if (dbg) dbg[2] = clock64();
        bar_sync(1, kConsumerThreads);   // D: digit planes complete
        if (dbg) dbg[3] = clock64();
        return s_x;
    };

    int token = m.st->token;
    bool have_prev = false;        // this CTA worked in its last phase and has not published that yet
    uint32_t prev_ep = 0;
    int prev_par = 0;
    unsigned int expected[2] = {0u, 0u};   // how far each phase counter will have got once the phases seen so far are done
    for (int s = 0; s < m.n_steps; ++s) {
        const bool sample = s >= m.first_sample;
        const int pos = pos0 + s;
        if (s < m.n_prompt) token = m.prompt[s];
        unsigned long long my_key = 0ull;
        for (int ph = 0; ph < m.nphases; ++ph) {
            const bool is_head = ph == m.nphases - 1;
            if (is_head && !sample) continue;
            const bool stamp = m.dbg != nullptr && s == 0 && blockIdx.x == 0 && tid == 0;
            long long* ts = m.dbg + (size_t)ph * kStampsPerPhase;
            if (stamp) ts[0] = clock64();
            const uint32_t out_ep = ep0 + (uint32_t)s * per_step + (uint32_t)ph + 1, in_ep = out_ep - 1;
            MegaLLPhase* sph = reinterpret_cast<MegaLLPhase*>(reinterpret_cast<uint8_t*>(sph2) + (size_t)(ph & 1) * kDescStride);
            if (tid < kDescWords) reinterpret_cast<uint32_t*>(sph)[tid] = desc_word;
            bar_sync(1, kConsumerThreads);   // A: descriptor staged, previous phase's statistics deposited
            {   // next descriptor: the phase after this one, or the first one of the next step
                int nph = ph + 1;
                if (nph == m.nphases - 1 && !sample) nph = 0;
                if (nph >= m.nphases) nph = 0;
                desc_prefetch(nph);
            }
            if (warp == 0 && have_prev) stats_publish(prev_ep, prev_par);
            have_prev = false;
            const MegaLLPhase& P = *sph;
            // who works in this phase (every CTA keeps the same books, whether it works or not)
            int nsplit = 1, chunk = 0, workers;
            if (P.type == PH_GEMV) {
                workers = P.g.L.P;
            } else {
                attn_split_range(pos + 1, P.at.max_splits, P.at.min_chunk, nsplit, chunk);
                workers = min(P.at.heads * nsplit, (int)gridDim.x);
            }
            const bool worked = (int)blockIdx.x < workers;
            const bool from_emb = P.type == PH_GEMV && P.x_src == SRC_EMB;
            const int ppar = (ph - 1) & 1;
            if (worked && !from_emb) wait_count(m.cnt + 32 * ppar, expected[ppar]);   // B1: the producers of my input are done
            expected[ph & 1] += (unsigned int)workers;
            if (stamp) { ts[1] = clock64(); ts[2] = ts[1]; ts[3] = ts[1]; }
            XStats out_st{0.f, 0.f};
            if (P.type == PH_GEMV) {
                if (worked) {
                    const GemvArgs& g = P.g;
                    const Slab slab = make_slab(g.L, blockIdx.x);
                    const PhaseCtx ctx{true, pos, nullptr, out_ep, P.knew_ll, P.vnew_ll, sm.keyred, P.resid_src == SRC_EMB};
                    const float* resid = P.resid_src == SRC_EMB ? m.emb + (size_t)token * m.H : g.resid;
                    const EpiPre pre = gemv_epilogue_prefetch(g, slab, resid, ctx, tid);
                    float s_x;
                    if (from_emb) {
                        auto emb_fn = [&]() -> XStats { return m.emb_stats[token]; };
                        s_x = gemv_stage_x_known<BITS>(g, m.emb + (size_t)token * m.H, sm, slab, false, emb_fn, tid, lane);
                    } else {
                        int nprod = P.prod_P;
                        if (P.prod_kind == 1) {
                            int ns, ch;
                            attn_split_range(pos + 1, P.prod_splits, P.prod_minchunk, ns, ch);
                            nprod = min(P.prod_heads * ns, (int)gridDim.x);
                        }
                        s_x = prologue_ll(g, slab, P.x_ll, in_ep, ppar, nprod, stamp ? ts + 6 : nullptr);
                    }
                    if (stamp) ts[2] = clock64();
                    gemv_consume<BITS>(g, slab, sm, it, warp, lane);
                    if (stamp) ts[3] = clock64();
                    out_st = gemv_epilogue(g, slab, sm, s_x, resid, ctx, pre, tid, lane);
                    if (is_head) {
                        bar_sync(1, kConsumerThreads);
                        unsigned long long k = 0ull;
#pragma unroll
                        for (int i = 0; i < kConsumerWarps; ++i) k = sm.keyred[i] > k ? sm.keyred[i] : k;
                        my_key = k;
                    }
                }
            } else if (worked) {
                out_st.am = mega_attention_ll(P.at, pos + 1, in_ep, out_ep, attn_sm);
            }
            if (worked) {
                stats_deposit(out_st, ph & 1);
                have_prev = true;
                prev_ep = out_ep;
                prev_par = ph & 1;
            }
            if (stamp) { ts[4] = clock64(); ts[5] = ts[4]; }
        }
        // the last phase's counter bump must not be lost: publish before the step exchange (its barrier orders the deposit)
        bar_sync(1, kConsumerThreads);
        if (warp == 0 && have_prev) stats_publish(prev_ep, prev_par);
        have_prev = false;
        const int tok = step_exchange(s, my_key);
        if (sample) {
            token = tok;
            publish(s, token);
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        m.st->pos = pos0 + m.n_steps;
        m.st->token = token;
        const int sampled = m.n_steps - m.first_sample;
        if (sampled > 0) m.st->step += sampled;
    }
}

}  // namespace tib
