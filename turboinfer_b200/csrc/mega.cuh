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

enum MegaPhaseType : int { PH_GEMV = 0, PH_ATTN = 1 };
enum MegaSrc : int { SRC_PTR = 0, SRC_EMB = 1 };

struct MegaPhase {
    int type;
    int x_src;      // PH_GEMV: where x comes from (pointer in g.x, or the embedding row of the current token)
    int resid_src;  // PH_GEMV + EPI_RESIDUAL: same for the residual input
    int is_head;    // lm_head phase: only run on steps that sample
    GemvArgs g;
    AttnArgs at;
};

struct MegaArgs {
    const MegaPhase* __restrict__ phases;
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
    unsigned int* grid_bar;       // zeroed by the host before every launch
    unsigned int* head_cnt;       // [heads]; zero between phases by construction
    unsigned long long* keys;     // [2] argmax keys, zeroed by the host before every launch
    float* logits;
    int stages;
    int max_kpad, max_items, attn_floats;
    long long* dbg;   // optional: CTA 0 writes 6 clock64 stamps per phase of step 0 (debug timeline)
};

TIB_HD size_t mega_smem_bytes(int stages, int max_kpad, int max_items, int attn_floats) {
    return (size_t)stages * kStageBytes + (size_t)max_kpad * 4 + (size_t)max_items * 4 + 32 * 4 + (size_t)attn_floats * 4 +
           (size_t)2 * kMaxStages * 8 + 16 + 128 + ((sizeof(MegaPhase) + 15) & ~size_t(15));
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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

TIB_HD int attn_scratch_floats(int D, int nt) {
    const int groups = D < nt ? nt / D : 1;
    return D + kAttnTokBlock + 32 + groups * D + 4;
}

// Computes the (m, l, o) partial of head h over tokens [t0, t1) and stores it.  q / K / V are read through L2.
template <int NT>
__device__ __forceinline__ void attn_item(const AttnArgs& a, int h, int j, int t0, int t1, float* sm, bool direct = false) {
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
        }
    } else {
#pragma unroll
        for (int i = 0; i < kMaxDims; ++i) {
            const int d = d0 + i * NT;
            if (d < D) po[d] = o[i];
        }
    }
    if (tid == 0 && !direct) {
        a.part_ml[((size_t)h * a.max_splits + j) * 2 + 0] = m_run;
        a.part_ml[((size_t)h * a.max_splits + j) * 2 + 1] = l_run;
    }
}

// merge the nsplit partials of head h (fixed order, independent of which CTA does it)
template <int NT>
__device__ __forceinline__ void attn_merge_head(const AttnArgs& a, int h, int nsplit) {
    const float* ml = a.part_ml + (size_t)h * a.max_splits * 2;
    float M = -INFINITY;
    for (int j = 0; j < nsplit; ++j) M = fmaxf(M, __ldcg(ml + 2 * j));
    float Lsum = 0.f;
    for (int j = 0; j < nsplit; ++j) Lsum += __ldcg(ml + 2 * j + 1) * expf(__ldcg(ml + 2 * j) - M);
    for (int d = threadIdx.x; d < a.D; d += NT) {
        float acc = 0.f;
        for (int j = 0; j < nsplit; ++j)
            acc = fmaf(__ldcg(a.part_o + ((size_t)h * a.max_splits + j) * a.D + d), expf(__ldcg(ml + 2 * j) - M), acc);
        a.out[h * a.D + d] = acc / Lsum;
    }
}

// stand-alone launches (TensorEngine::attention_fast_incremental / multi_head_attention entry points, per-op engine)
__global__ void __launch_bounds__(kAttnThreads) attn_partial_kernel(const AttnArgs a) {
    extern __shared__ float attn_dyn_smem[];
    const int t = *a.pos_ptr + a.t_bias;
    int nsplit, chunk;
    attn_split_range(t, a.max_splits, a.min_chunk, nsplit, chunk);
    const int j = blockIdx.y;
    if (j >= nsplit) return;
    attn_item<kAttnThreads>(a, blockIdx.x, j, j * chunk, min(t, (j + 1) * chunk), attn_dyn_smem, nsplit == 1);
}
__global__ void __launch_bounds__(kAttnThreads) attn_combine_kernel(const AttnArgs a) {
    const int t = *a.pos_ptr + a.t_bias;
    int nsplit, chunk;
    attn_split_range(t, a.max_splits, a.min_chunk, nsplit, chunk);
    if (nsplit == 1) return;  // attn_partial_kernel wrote the output directly
    attn_merge_head<kAttnThreads>(a, blockIdx.x, nsplit);
}

// the attention phase of the persistent kernel: (head, split) items dealt round-robin to the CTAs; the CTA that
// completes the last split of a head merges that head's partials (no extra grid barrier, deterministic result)
__device__ __forceinline__ void mega_attention(const AttnArgs& a, int t, unsigned int* head_cnt, float* sm) {
    constexpr int NT = kConsumerThreads;
    int nsplit, chunk;
    attn_split_range(t, a.max_splits, a.min_chunk, nsplit, chunk);
    const int items = a.heads * nsplit;
    int* flag = reinterpret_cast<int*>(sm + attn_scratch_floats(a.D, NT) - 4);
    for (int i = blockIdx.x; i < items; i += gridDim.x) {
        const int h = i / nsplit, j = i - h * nsplit;
        const int t0 = j * chunk, t1 = min(t, t0 + chunk);
        if (nsplit == 1) {  // short context: one CTA per head does everything, no partials / counters / merge
            attn_item<NT>(a, h, 0, 0, t, sm, true);
            bar_sync(1, NT);
            continue;
        }
        attn_item<NT>(a, h, j, t0, t1, sm);
        __threadfence();
        bar_sync(1, NT);
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(&head_cnt[h], 1u);
            *flag = (prev == (unsigned int)(nsplit - 1)) ? 1 : 0;
            if (*flag) head_cnt[h] = 0u;  // ready for the next layer
            __threadfence();
        }
        bar_sync(1, NT);
        if (*flag) attn_merge_head<NT>(a, h, nsplit);
        bar_sync(1, NT);
    }
}

// Registers: more than 16 warps put 5 on an SM sub-partition, which caps a uniform allocation at 96 per thread.  The
// kernel is compiled for 96 (__maxnreg__); the producer warpgroup then shrinks to 24 and the four consumer warpgroups
// grow to 112 (the pool is the CTA's launch allocation: 640 x 96 >= 512 x 112 + 128 x 24).  setmaxnreg is a
// warpgroup-wide instruction, so the producer warp comes with three idle siblings (warps 17..19) that only execute
// the shrink and exit: the CTA has 20 warps.
constexpr int kMegaThreads = (kConsumerWarps + 4) * 32;  // 640
template <int BITS>
__global__ void __maxnreg__(96) mega_decode_kernel(const __grid_constant__ MegaArgs m) {
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // carve: ring | xs | part | red | attention scratch | mbarriers
    GemvSmem sm;
    uintptr_t p = (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127);
    sm.ring = reinterpret_cast<uint8_t*>(p);
    p += (size_t)m.stages * kStageBytes;
    sm.xs = reinterpret_cast<float*>(p);
    p += (size_t)m.max_kpad * 4;
    sm.part = reinterpret_cast<float*>(p);
    p += (size_t)m.max_items * 4;
    sm.red = reinterpret_cast<float*>(p);
    p += 32 * 4;
    float* attn_sm = reinterpret_cast<float*>(p);
    p += (size_t)m.attn_floats * 4;
    p = (p + 15) & ~uintptr_t(15);
    sm.full = reinterpret_cast<uint64_t*>(p);
    sm.empty = sm.full + kMaxStages;
    p += (size_t)2 * kMaxStages * 8;
    MegaPhase* sph = reinterpret_cast<MegaPhase*>(p);  // this phase's descriptor, staged by the consumers
    if (tid == 0) gemv_init_barriers(sm, m.stages);
    __syncthreads();

    const int pos0 = m.st->pos;
    uint32_t it = 0;

    if (warp >= kConsumerWarps) {
        // ===== producer warpgroup: warp 16 streams every GEMV phase of every step, back to back =====
        reg_dealloc<24>();
        if (warp == kConsumerWarps && lane == 0) {
            for (int s = 0; s < m.n_steps; ++s) {
                const bool sample = s >= m.first_sample;
                for (int ph = 0; ph < m.nphases; ++ph) {
                    const MegaPhase& P = m.phases[ph];
                    if (P.type != PH_GEMV || (P.is_head && !sample)) continue;
                    if ((int)blockIdx.x >= P.g.L.P) continue;
                    const Slab slab = make_slab(P.g.L, blockIdx.x);
                    gemv_produce(P.g, slab, sm, it);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    reg_alloc<112>();
    unsigned int bar_target = 0;
    bool need_wait = false;
    auto grid_arrive = [&]() {
        bar_sync(1, kConsumerThreads);  // every consumer thread's stores are ordered before thread 0's release
        if (tid == 0) red_release_add(m.grid_bar, 1u);
        bar_target += gridDim.x;
        need_wait = true;
    };
    auto grid_wait = [&]() {
        if (!need_wait) return;
        if (tid == 0) {
            const long long t0 = clock64();
            while (ld_acquire_u32(m.grid_bar) < bar_target) {
                if (clock64() - t0 > 8000000000LL) __trap();
            }
        }
        bar_sync(1, kConsumerThreads);
        need_wait = false;
    };
    auto decode_key = [&](int s) -> int {
        const unsigned long long key = __ldcg(&m.keys[s & 1]);
        return 0x7FFFFFFF - (int)(uint32_t)(key & 0xFFFFFFFFull);
    };
    // CTA 0 publishes the token picked in step s (called by all its consumer threads, after that step's barrier)
    auto publish = [&](int s, int tok) {
        if (blockIdx.x != 0) return;
        const int k = m.st->step + (s - m.first_sample);
        if (m.io->hist && k < m.io->hist_cap)
            for (int i = tid; i < m.V; i += kConsumerThreads) m.io->hist[(size_t)k * m.V + i] = __ldcg(m.logits + i);
        if (tid == 0) {
            if (m.io->out_tokens && k < m.io->out_cap) m.io->out_tokens[k] = tok;
            m.keys[(s + 1) & 1] = 0ull;
        }
    };

    int token = m.st->token;  // decode-only launches continue from the token the previous launch picked
    for (int s = 0; s < m.n_steps; ++s) {
        const bool sample = s >= m.first_sample;
        const int pos = pos0 + s;
        if (s > 0 && s - 1 >= m.first_sample) {
            grid_wait();
            token = decode_key(s - 1);
            publish(s - 1, token);
        }
        if (s < m.n_prompt) token = m.prompt[s];
        for (int ph = 0; ph < m.nphases; ++ph) {
            const bool is_head = ph == m.nphases - 1;  // the lm_head is always the last phase
            if (is_head && !sample) continue;
            const bool stamp = m.dbg != nullptr && s == 0 && blockIdx.x == 0 && tid == 0;
            long long* ts = m.dbg + (size_t)ph * 6;
            if (stamp) ts[0] = clock64();
            // Everything that does not depend on the previous phase's output happens BEFORE the grid barrier:
            // stage the phase descriptor in shared memory, fetch the epilogue's per-column constants.
            const MegaPhase& PG = m.phases[ph];
            {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(&PG);
                uint32_t* dst = reinterpret_cast<uint32_t*>(sph);
                for (int i = tid; i < (int)(sizeof(MegaPhase) / 4); i += kConsumerThreads) dst[i] = src[i];
            }
            const bool gemv_here = PG.type == PH_GEMV && (int)blockIdx.x < PG.g.L.P;
            const PhaseCtx ctx{true, pos, is_head ? &m.keys[s & 1] : nullptr};
            Slab slab{};
            EpiPre pre{};
            const float* resid = nullptr;
            if (gemv_here) {
                slab = make_slab(PG.g.L, blockIdx.x);
                resid = PG.resid_src == SRC_EMB ? m.emb + (size_t)token * m.H : PG.g.resid;
                pre = gemv_epilogue_prefetch(PG.g, slab, resid, ctx, tid);
            }
            if (need_wait) grid_wait(); else bar_sync(1, kConsumerThreads);
            const MegaPhase& P = *sph;
            if (stamp) { ts[1] = clock64(); ts[2] = ts[1]; ts[3] = ts[1]; }
            if (P.type == PH_GEMV) {
                if (gemv_here) {
                    const float* x = P.x_src == SRC_EMB ? m.emb + (size_t)token * m.H : P.g.x;
                    const GemvArgs& g = P.g;
                    const float sumx = gemv_stage_x<BITS>(g, x, sm, P.x_src != SRC_EMB, g.colzterm != nullptr, tid, warp, lane);
                    if (stamp) ts[2] = clock64();
                    gemv_consume<BITS>(g, slab, sm, it, warp, lane);
                    if (stamp) ts[3] = clock64();
                    gemv_epilogue(g, slab, sm, sumx, resid, ctx, pre, tid, lane);
                }
            } else {
                mega_attention(P.at, pos + 1, m.head_cnt, attn_sm);
            }
            if (stamp) ts[4] = clock64();
            grid_arrive();
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
        m.st->pos = pos0 + m.n_steps;
        m.st->token = token;
        const int sampled = m.n_steps - m.first_sample;
        if (sampled > 0) m.st->step += sampled;
    }
}

}  // namespace tib
