// sampling.cuh -- on-device sample_next_token (src/model/inference_engine.cpp:1554-1673): temperature -> top-k -> softmax ->
// top-p -> inverse CDF, plus the log-probabilities of compute_logprobs (:873-954).  The reference does this on the host
// with two std::sort calls over the vocabulary per token; here one CTA per sequence does it next to the logits, and the
// token never leaves the device between steps.
//
// Randomness: the reference draws from a time-seeded std::mt19937 (:472), so its stream cannot be reproduced.  Here the
// uniform of step s is a counter-based hash of (seed, s) (sample_uniform below; the test-side restatement of the reference algorithm takes the same uniform), which
// makes a generation reproducible and testable.
//
// Arithmetic contract (what tests/test_gpu_sampling.py checks against the CPU restatement of the reference algorithm on the same logits):
//   * top_k in [1, 1024] (the reference default is 50): the survivors are gathered in index order into shared memory and
//     every sum the reference computes sequentially over the vocabulary (softmax denominator, top-p cumulative sum in sorted
//     order, renormalisation, inverse CDF) is computed sequentially over the survivors in the same order -- adding the
//     exact zeros of the filtered entries changes nothing, so the result equals the reference's arithmetic up to the last
//     bit of expf (CUDA's expf vs glibc's).  Equal logits / probabilities: std::sort is unstable, i.e. the reference's
//     order among ties is unspecified; here (and in the test-side restatement) ties go by ascending index.
//   * top_k = 0 or > 1024 (no top-k filter, or a very wide one): the same pipeline with block-wide parallel sums and a
//     radix search for the top-p cut -- the same distribution, sums in a different order (~1e-6 relative), ties at the
//     top-p cut kept as a group.
#pragma once
#include "kernels.cuh"

namespace tib {

constexpr int kSampleThreads = 1024;
constexpr int kTopKMax = 1024;

// uniform in [0, 1) with 24 bits, from (seed, step); the same function on the host side of the tests
TIB_HD float sample_uniform(uint64_t seed, uint64_t step) {
    const uint64_t h = splitmix64(splitmix64(seed) ^ (step * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull));
    return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}

struct SampleArgs {
    const float* logits;   // [V] of this sequence (blockIdx.x selects the row: logits + row * ld)
    int V, ld;
    float temperature;
    int top_k;
    float top_p;
    uint64_t seed;
    const int* step_ptr;   // device scalar: index of the token being sampled (the counter of the RNG); may be null -> step
    int step;
    int* token_out;        // [rows]
    float* logprob_out;    // [rows] or null
    int* hist_tokens;      // optional history: hist_tokens[row * hist_stride + step]
    float* hist_logprobs;
    int hist_stride;
    int* feed_token;       // optional: where the next decode step reads its input token (StepState::token)
    int staged;            // 1: the launch carries sample_row_smem_bytes(V) of dynamic shared memory for the scaled row
};
// dynamic shared memory of sample_kernel / beam_expand_kernel for a row of V logits (0: too long, read from global memory)
TIB_HD size_t sample_row_smem_bytes(int V) {
    const int per = (V + kSampleThreads - 1) / kSampleThreads;
    const size_t bytes = (size_t)kSampleThreads * (size_t)(per | 1) * sizeof(float);
    return bytes <= (size_t)176 * 1024 ? bytes : 0;
}
// the <32> instances (a thread's keys in registers) serve staged rows of at most 32 logits per thread
TIB_HD bool sample_keys_in_registers(int V) { return sample_row_smem_bytes(V) != 0 && (V + kSampleThreads - 1) / kSampleThreads <= 32; }

__device__ __forceinline__ uint32_t desc_key(float v) {   // larger value -> larger key
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_exscan(int v, int* wsum, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        wsum[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    const int r = wsum[warp] + inc - v;
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_exscan_f(float v, float* wsum, float* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        float w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0.f;
        float winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        wsum[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    const float r = wsum[warp] + inc - v;
    __syncthreads();
    return r;
}

// Block-wide sums that every thread receives (1 024 threads): warp reduction, one shared word per warp, every warp sums the 32
// words.  `wbuf` holds two sets of 32 words used alternately (one __syncthreads per sum); the order of the additions is fixed,
// so float sums are deterministic.
__device__ __forceinline__ int block_sum_i(int v, int* wbuf, int& parity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    int* buf = wbuf + 32 * parity;
    if (lane == 0) buf[warp] = v;
    __syncthreads();
    parity ^= 1;
    return __reduce_add_sync(0xffffffffu, buf[lane]);
}
__device__ __forceinline__ float block_sum_f(float v, float* wbuf, int& parity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    float* buf = wbuf + 32 * parity;
    if (lane == 0) buf[warp] = v;
    __syncthreads();
    parity ^= 1;
    float t = buf[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// Key of the `need`-th largest element: a bit-by-bit search from the top, each step one block-wide count of the keys >= the
// candidate (32 counts of register-held keys: ~5 us.  The first two versions used radix histograms -- 32 000 shared atomics on a
// few hot buckets, then per-warp histograms with __match_any_sync -- and spent 140-200 us per token in them).
// key_k(k) = key of the thread's k-th own element (0 beyond its range: never counted, a real value's key is never 0).
// Returns the key; *need_eq = how many of the elements EQUAL to it belong to the top `need` (the lowest indices).
template <int P, typename KeyK>
__device__ __forceinline__ uint32_t select_kth_key(KeyK key_k, int per, int need, int* wbuf, int& parity, int* need_eq) {
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = prefix | (1u << bit);
        int c = 0;
#pragma unroll
        for (int k = 0; k < (P ? P : per); ++k) c += key_k(k) >= cand ? 1 : 0;
        if (block_sum_i(c, wbuf, parity) >= need) prefix = cand;
    }
    int c = 0;
#pragma unroll
    for (int k = 0; k < (P ? P : per); ++k) c += key_k(k) > prefix ? 1 : 0;
    *need_eq = need - block_sum_i(c, wbuf, parity);
    return prefix;
}
// The largest key T (>= floor_key) such that the mass of the elements with key >= T reaches `target` (0 when even all of them
// do not, by rounding): the nucleus cut of the wide path.  mass_k(k) = the element's (unnormalised) probability.
template <int P, typename KeyK, typename MassK>
__device__ __forceinline__ uint32_t select_mass_key(KeyK key_k, MassK mass_k, int per, uint32_t floor_key, float target, float* wbuf, int& parity) {
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = prefix | (1u << bit);
        float mass = 0.f;
#pragma unroll
        for (int k = 0; k < (P ? P : per); ++k) {
            const uint32_t key = key_k(k);
            mass += (key >= cand && key >= floor_key) ? mass_k(k) : 0.f;
        }
        if (block_sum_f(mass, wbuf, parity) >= target) prefix = cand;
    }
    return prefix > floor_key ? prefix : floor_key;
}

// P = 32: rows of up to 32 768 logits staged in shared memory, every thread's 32 keys held in registers (fully unrolled loops);
// P = 0: any row, keys recomputed from the row at every use.
template <int P>
__global__ void __launch_bounds__(kSampleThreads) sample_kernel(const SampleArgs a) {
    __shared__ int wbuf_i[64];
    __shared__ float wbuf_f[64];
    int parity_i = 0, parity_f = 0;
    __shared__ int wsum[32];
    __shared__ float wsumf[32];
    __shared__ int s_tot, s_sel, s_token;
    __shared__ float s_ftot, s_max, s_sum, s_lp;
    __shared__ int sidx[kTopKMax];
    __shared__ float sval[kTopKMax];      // scaled logit, then probability
    __shared__ int sord[kTopKMax];        // positions sorted by (probability desc, index asc)
    const int tid = threadIdx.x, V = a.V;
    const float* lg = a.logits + (size_t)blockIdx.x * a.ld;
    const int step = a.step_ptr ? *a.step_ptr : a.step;
    const float u = sample_uniform(a.seed + (uint64_t)blockIdx.x * 0x51ED27ull, (uint64_t)step);
    const bool scale = a.temperature != 1.0f && a.temperature > 0.0f;                      // :1578-1582
    // indices in blocked order: thread t owns [t * per, (t + 1) * per) -- a compaction by thread order is then in index order
    const int per = (V + kSampleThreads - 1) / kSampleThreads;
    const int i0 = min(tid * per, V), i1 = min(i0 + per, V);
    // The row is read many times (4 radix passes, gathers, sums).  In the blocked order a warp's load touches 32 different cache
    // lines -- 32 LSU transactions per instruction, ~16 us per pass over 32 000 logits, 144-200 us per token in the first version
    // of this kernel -- so the scaled row is staged ONCE in shared memory (coalesced global reads), each thread's range at an odd
    // stride (conflict-free).  Rows too long for shared memory are read from global memory as before.
    extern __shared__ float s_row[];
    const int stride = per | 1;
    if (a.staged) {
        for (int i = tid; i < V; i += kSampleThreads) {
            const float x = lg[i];
            s_row[(i / per) * stride + (i % per)] = scale ? x / a.temperature : x;
        }
        __syncthreads();
    }
    const float* const own = s_row + tid * stride - i0;   // own[i] for i in [i0, i1)
    auto val = [&](int i) -> float {   // an element of this thread's own range
        if (a.staged) return own[i];
        const float x = lg[i];
        return scale ? x / a.temperature : x;
    };
    auto val_any = [&](int i) -> float {   // any element
        if (a.staged) return s_row[(i / per) * stride + (i % per)];
        const float x = lg[i];
        return scale ? x / a.temperature : x;
    };
    const bool use_topk = a.top_k > 0 && a.top_k < V;                                        // :1585
    const bool exact = use_topk && a.top_k <= kTopKMax;
    const int cnt = i1 - i0;
    uint32_t kreg[P ? P : 1];
    if (P) {
#pragma unroll
        for (int k = 0; k < (P ? P : 1); ++k) kreg[k] = k < cnt ? desc_key(own[i0 + k]) : 0u;
    }
    auto key_k = [&](int k) -> uint32_t {
        if (P) return kreg[k];
        return k < cnt ? desc_key(val(i0 + k)) : 0u;
    };

    if (exact) {
        // ---- the k-th largest key ----
        int need = 0;   // how many of the elements equal to that key survive (lowest indices)
        const uint32_t tau = select_kth_key<P>(key_k, per, a.top_k, wbuf_i, parity_i, &need);
        // ---- gather the survivors in index order ----
        int cgt = 0, ceq = 0;
        for (int i = i0; i < i1; ++i) {
            const uint32_t key = desc_key(val(i));
            cgt += key > tau ? 1 : 0;
            ceq += key == tau ? 1 : 0;
        }
        const int eq_before = block_exscan(ceq, wsum, &s_tot);
        int keep = cgt, eqr = eq_before;
        for (int i = i0; i < i1; ++i)
            if (desc_key(val(i)) == tau) { keep += eqr < need ? 1 : 0; ++eqr; }
        int off = block_exscan(keep, wsum, &s_tot);   // the total lands in shared memory: every thread reads it
        eqr = eq_before;
        for (int i = i0; i < i1; ++i) {
            const float x = val(i);
            const uint32_t key = desc_key(x);
            bool take = key > tau;
            if (key == tau) { take = eqr < need; ++eqr; }
            if (take) { sidx[off] = i; sval[off] = x; ++off; }
        }
        __syncthreads();
        const int n = s_tot;   // == k
        // ---- softmax over the survivors (:1600-1612): exp in parallel, the sum sequentially in index order ----
        if (tid == 0) {
            float mx = sval[0];
            for (int i = 1; i < n; ++i) mx = fmaxf(mx, sval[i]);
            s_max = mx;
        }
        __syncthreads();
        for (int i = tid; i < n; i += kSampleThreads) sval[i] = expf(sval[i] - s_max);
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int i = 0; i < n; ++i) s += sval[i];
            s_sum = s;
        }
        __syncthreads();
        for (int i = tid; i < n; i += kSampleThreads) sval[i] = sval[i] / s_sum;
        __syncthreads();
        // ---- top-p (:1614-1648): rank by (probability desc, index asc), sequential cumulative sum in that order ----
        if (a.top_p < 1.0f) {
            for (int i = tid; i < n; i += kSampleThreads) {
                const float p = sval[i];
                int rank = 0;
                for (int j = 0; j < n; ++j) {
                    const float q = sval[j];
                    rank += (q > p || (q == p && j < i)) ? 1 : 0;   // survivors are in index order: j < i <=> smaller index
                }
                sord[rank] = i;
            }
            __syncthreads();
            if (tid == 0) {
                float c = 0.f;
                int cutoff = n;
                for (int r = 0; r < n; ++r) {
                    c += sval[sord[r]];
                    if (c >= a.top_p) { cutoff = r + 1; break; }
                }
                s_sel = cutoff;
            }
            __syncthreads();
            for (int r = s_sel + tid; r < n; r += kSampleThreads) sval[sord[r]] = 0.0f;
            __syncthreads();
            if (tid == 0) {
                float s = 0.f;
                for (int i = 0; i < n; ++i) s += sval[i];
                s_sum = s;
            }
            __syncthreads();
            if (s_sum > 0.0f)
                for (int i = tid; i < n; i += kSampleThreads) sval[i] = sval[i] / s_sum;
            __syncthreads();
        }
        // ---- inverse CDF in index order (:1650-1666) ----
        if (tid == 0) {
            float c = 0.f;
            int tok = V - 1;
            float lp = -INFINITY;   // the fall-back of :1668-1672: last token, log of ITS probability (0 unless it survived)
            bool found = false;
            if (u <= 0.0f) {   // the reference loop's `random_value <= cumsum` holds at index 0 for a uniform of exactly 0
                tok = 0;
                lp = (n > 0 && sidx[0] == 0) ? logf(sval[0]) : -INFINITY;
                found = true;
            }
            for (int i = 0; i < n && !found; ++i) {
                c += sval[i];
                if (u <= c) { tok = sidx[i]; lp = logf(sval[i]); found = true; break; }
            }
            if (!found && n > 0 && sidx[n - 1] == V - 1) lp = logf(sval[n - 1]);
            s_token = tok;
            s_lp = lp;
        }
        __syncthreads();
    } else {
        // ---- no top-k filter (or a wider one than the exact path holds): block-wide sums ----
        // (a top_k > 1024 is applied as a key threshold first, found with the same radix select on counts)
        uint32_t kth = 0;   // survivors: key >= kth
        if (use_topk) {
            int need_eq = 0;
            kth = select_kth_key<P>(key_k, per, a.top_k, wbuf_i, parity_i, &need_eq);
        }
        float mx = -INFINITY;
        for (int i = i0; i < i1; ++i) {
            const float x = val(i);
            if (desc_key(x) >= kth) mx = fmaxf(mx, x);
        }
        mx = warp_max(mx);
        if ((tid & 31) == 0) wsumf[tid >> 5] = mx;
        __syncthreads();
        if (tid == 0) {
            float m2 = -INFINITY;
            for (int w = 0; w < kSampleThreads / 32; ++w) m2 = fmaxf(m2, wsumf[w]);
            s_max = m2;
        }
        __syncthreads();
        auto prob_un = [&](int i) -> float {   // unnormalised probability of a survivor, 0 otherwise
            const float x = val(i);
            return desc_key(x) >= kth ? expf(x - s_max) : 0.f;
        };
        float part = 0.f;
        for (int i = i0; i < i1; ++i) part += prob_un(i);
        float Z;
        (void)block_exscan_f(part, wsumf, &s_ftot);
        Z = s_ftot;
        uint32_t pth = kth;   // top-p: survivors are the keys >= pth
        if (a.top_p < 1.0f) {
            // the largest threshold whose mass from the top reaches top_p * Z (bit-by-bit search; the masses are held in registers)
            float ereg[P ? P : 1];
            if (P) {
#pragma unroll
                for (int k = 0; k < (P ? P : 1); ++k) ereg[k] = k < cnt ? expf(own[i0 + k] - s_max) : 0.f;
            }
            auto mass_k = [&](int k) -> float {
                if (P) return ereg[k];
                return k < cnt ? expf(val(i0 + k) - s_max) : 0.f;
            };
            pth = select_mass_key<P>(key_k, mass_k, per, kth, a.top_p * Z, wbuf_f, parity_f);
        }
        // renormalise over the final survivors and walk the CDF: per-thread sequential, block-wide exclusive scan between
        float mine = 0.f;
        for (int i = i0; i < i1; ++i) {
            const float x = val(i);
            if (desc_key(x) >= pth) mine += expf(x - s_max);
        }
        float Z2;
        const float before = block_exscan_f(mine, wsumf, &s_ftot);
        Z2 = s_ftot;
        if (tid == 0) { s_token = -1; s_lp = -INFINITY; }
        __syncthreads();
        const float target = u * Z2;
        // the owner of the target: the first thread (in index order) whose inclusive prefix reaches it
        if (mine > 0.f && before < target && target <= before + mine) {
            float c = before;
            int tok = -1;
            float pr = 0.f;
            for (int i = i0; i < i1; ++i) {
                const float x = val(i);
                if (desc_key(x) >= pth) {
                    const float p = expf(x - s_max);
                    c += p;
                    tok = i;
                    pr = p;
                    if (target <= c) break;
                }
            }
            s_token = tok;
            s_lp = logf(pr / Z2);
        }
        __syncthreads();
        if (tid == 0 && s_token < 0) {
            // u == 0 exactly (target 0: the first survivor), or rounding left the target just above the total: the reference
            // takes the first entry with u <= cumsum / falls back to the last token (:1668-1672)
            int tok = V - 1;
            float lp = -INFINITY;
            if (target <= 0.f) {   // uniform of exactly 0: index 0 (see the exact path)
                tok = 0;
                lp = desc_key(val_any(0)) >= pth ? logf(expf(val_any(0) - s_max) / Z2) : -INFINITY;
            } else if (desc_key(val_any(V - 1)) >= pth) {
                lp = logf(expf(val_any(V - 1) - s_max) / Z2);
            }
            s_token = tok;
            s_lp = lp;
        }
        __syncthreads();
    }
    if (tid == 0) {
        a.token_out[blockIdx.x] = s_token;
        if (a.logprob_out) a.logprob_out[blockIdx.x] = s_lp;
        if (a.hist_tokens && step < a.hist_stride) a.hist_tokens[(size_t)blockIdx.x * a.hist_stride + step] = s_token;
        if (a.hist_logprobs && step < a.hist_stride) a.hist_logprobs[(size_t)blockIdx.x * a.hist_stride + step] = s_lp;
        if (a.feed_token) a.feed_token[blockIdx.x] = s_token;
    }
}

// ---------------------------------------------------------------------------------------------------
// beam_search_decode's expansion of one candidate (:1964-2018): probs = softmax(logits / T) over the vocabulary -> top-k filter
// on the PROBABILITIES + renormalise (:1821-1856) -> top-p filter + renormalise (:1858-1909) -> the beam_size most probable
// tokens with probability > 0, most probable first.  One CTA per candidate (row of the batched step's logits); the host
// only receives [beam] (probability, token) pairs per row.
//   * the softmax denominator over the whole vocabulary is a block-wide sum (the reference adds 32000 terms sequentially in
//     fp32; with a top-k or top-p filter the denominator cancels in the renormalisation, without any filter the reference's
//     own accumulated rounding shows: ~3e-5 relative at V = 32000);
//   * top_k in [1, 1024]: every later sum runs sequentially over the survivors in index order, as the reference's loops do;
//   * wider / no top-k: block-wide sums and a radix search for the nucleus, like sample_kernel's wide path;
//   * equal probabilities: lower token id first (the reference's std::sort leaves the order of ties unspecified).
// ---------------------------------------------------------------------------------------------------
constexpr int kBeamMax = 64;
struct BeamExpandArgs {
    const float* logits;   // [rows][ld]
    int V, ld;
    float temperature;
    int top_k;
    float top_p;
    int beam;
    float* cand_prob;      // [rows][beam]
    int* cand_tok;         // [rows][beam]
    int* cand_cnt;         // [rows]
    int staged;            // as SampleArgs::staged
};
template <int P>
__global__ void __launch_bounds__(kSampleThreads) beam_expand_kernel(const BeamExpandArgs a) {
    __shared__ int wbuf_i[64];
    __shared__ float wbuf_f[64];
    int parity_i = 0, parity_f = 0;
    __shared__ int wsum[32];
    __shared__ float wsumf[32];
    __shared__ int s_tot, s_sel;
    __shared__ float s_ftot, s_max, s_sum;
    __shared__ int sidx[kTopKMax];
    __shared__ float sval[kTopKMax];
    __shared__ int sord[kTopKMax];
    const int tid = threadIdx.x, V = a.V;
    const float* lg = a.logits + (size_t)blockIdx.x * a.ld;
    const bool scale = a.temperature != 1.0f;                                                 // :1972-1976
    const int per = (V + kSampleThreads - 1) / kSampleThreads;
    const int i0 = min(tid * per, V), i1 = min(i0 + per, V);
    extern __shared__ float s_row[];   // the scaled row, staged once (see sample_kernel)
    const int stride = per | 1;
    if (a.staged) {
        for (int i = tid; i < V; i += kSampleThreads) {
            const float x = lg[i];
            s_row[(i / per) * stride + (i % per)] = scale ? x / a.temperature : x;
        }
        __syncthreads();
    }
    const float* const own = s_row + tid * stride - i0;
    auto val = [&](int i) -> float {   // an element of this thread's own range
        if (a.staged) return own[i];
        const float x = lg[i];
        return scale ? x / a.temperature : x;
    };
    auto key_of = [&](int i) -> uint32_t { return desc_key(val(i)); };
    const int cnt = i1 - i0;
    uint32_t kreg[P ? P : 1];
    if (P) {
#pragma unroll
        for (int k = 0; k < (P ? P : 1); ++k) kreg[k] = k < cnt ? desc_key(own[i0 + k]) : 0u;
    }
    auto key_k = [&](int k) -> uint32_t {
        if (P) return kreg[k];
        return k < cnt ? desc_key(val(i0 + k)) : 0u;
    };
    const bool use_topk = a.top_k > 0 && a.top_k < V;                                         // :1982
    const bool exact = use_topk && a.top_k <= kTopKMax;
    // ---- softmax over the vocabulary (:1798-1819) ----
    float mx = -INFINITY;
    for (int i = i0; i < i1; ++i) mx = fmaxf(mx, val(i));
    mx = warp_max(mx);
    if ((tid & 31) == 0) wsumf[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        float m2 = -INFINITY;
        for (int w = 0; w < kSampleThreads / 32; ++w) m2 = fmaxf(m2, wsumf[w]);
        s_max = m2;
    }
    __syncthreads();
    float part = 0.f;
    for (int i = i0; i < i1; ++i) part += expf(val(i) - s_max);
    (void)block_exscan_f(part, wsumf, &s_ftot);
    const float S = s_ftot;
    auto prob = [&](int i) -> float { return expf(val(i) - s_max) / S; };
    int n = 0;   // entries of the candidate list (sidx, sval), in index order
    if (exact) {
        int need = 0;
        const uint32_t tau = select_kth_key<P>(key_k, per, a.top_k, wbuf_i, parity_i, &need);
        int cgt = 0, ceq = 0;
        for (int i = i0; i < i1; ++i) {
            const uint32_t key = key_of(i);
            cgt += key > tau ? 1 : 0;
            ceq += key == tau ? 1 : 0;
        }
        const int eq_before = block_exscan(ceq, wsum, &s_tot);
        int keep = cgt, eqr = eq_before;
        for (int i = i0; i < i1; ++i)
            if (key_of(i) == tau) { keep += eqr < need ? 1 : 0; ++eqr; }
        int off = block_exscan(keep, wsum, &s_tot);
        eqr = eq_before;
        for (int i = i0; i < i1; ++i) {
            const uint32_t key = key_of(i);
            bool take = key > tau;
            if (key == tau) { take = eqr < need; ++eqr; }
            if (take) { sidx[off] = i; sval[off] = prob(i); ++off; }
        }
        __syncthreads();
        n = s_tot;
        // renormalise over the top-k set (:1845-1854), sequentially in index order
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < n; ++i) t += sval[i];
            s_sum = t;
        }
        __syncthreads();
        if (s_sum > 0.0f)
            for (int i = tid; i < n; i += kSampleThreads) sval[i] = sval[i] / s_sum;
        __syncthreads();
        if (a.top_p < 1.0f) {   // (:1858-1909) nucleus in (probability desc, index asc) order, sequential cumulative sum
            for (int i = tid; i < n; i += kSampleThreads) {
                const float p = sval[i];
                int rank = 0;
                for (int j = 0; j < n; ++j) {
                    const float q = sval[j];
                    rank += (q > p || (q == p && j < i)) ? 1 : 0;
                }
                sord[rank] = i;
            }
            __syncthreads();
            if (tid == 0) {
                float c = 0.f;
                int cutoff = n;
                for (int r = 0; r < n; ++r) {
                    c += sval[sord[r]];
                    if (c >= a.top_p) { cutoff = r + 1; break; }
                }
                s_sel = cutoff;
            }
            __syncthreads();
            for (int r = s_sel + tid; r < n; r += kSampleThreads) sval[sord[r]] = 0.0f;
            __syncthreads();
            if (tid == 0) {
                float t = 0.f;
                for (int i = 0; i < n; ++i) t += sval[i];
                s_sum = t;
            }
            __syncthreads();
            if (s_sum > 0.0f)
                for (int i = tid; i < n; i += kSampleThreads) sval[i] = sval[i] / s_sum;
            __syncthreads();
        }
    } else {
        // ---- wide: thresholds on the key instead of a survivor list; only the `beam` best are gathered at the end ----
        uint32_t kth = 0;
        float Zk = 1.0f;   // mass of the top-k set (1: no top-k filter, nothing is renormalised, :1824-1826)
        if (use_topk) {
            int need_eq = 0;
            kth = select_kth_key<P>(key_k, per, a.top_k, wbuf_i, parity_i, &need_eq);
            float mine = 0.f;
            for (int i = i0; i < i1; ++i)
                if (key_of(i) >= kth) mine += prob(i);
            (void)block_exscan_f(mine, wsumf, &s_ftot);
            Zk = s_ftot;
        }
        uint32_t pth = kth;
        float Z2 = Zk;
        if (a.top_p < 1.0f) {
            float ereg[P ? P : 1];   // cumulative RENORMALISED probability >= top_p, i.e. mass >= top_p * Zk
            if (P) {
#pragma unroll
                for (int k = 0; k < (P ? P : 1); ++k) ereg[k] = k < cnt ? prob(i0 + k) : 0.f;
            }
            auto mass_k = [&](int k) -> float {
                if (P) return ereg[k];
                return k < cnt ? prob(i0 + k) : 0.f;
            };
            pth = select_mass_key<P>(key_k, mass_k, per, kth, a.top_p * Zk, wbuf_f, parity_f);
            float mine = 0.f;
            for (int i = i0; i < i1; ++i)
                if (key_of(i) >= pth) mine += prob(i);
            (void)block_exscan_f(mine, wsumf, &s_ftot);
            Z2 = s_ftot;
        }
        // the `beam` largest keys among the survivors (key >= pth), gathered in index order
        int cs = 0;
        for (int i = i0; i < i1; ++i) cs += key_of(i) >= pth ? 1 : 0;
        (void)block_exscan(cs, wsum, &s_tot);
        const int survivors = s_tot;
        const int want = min(a.beam, survivors);
        int need = 0;
        const uint32_t tau = select_kth_key<P>(key_k, per, want, wbuf_i, parity_i, &need);   // want <= survivors: tau >= pth
        int cgt = 0, ceq = 0;
        for (int i = i0; i < i1; ++i) {
            const uint32_t key = key_of(i);
            cgt += key > tau ? 1 : 0;
            ceq += key == tau ? 1 : 0;
        }
        const int eq_before = block_exscan(ceq, wsum, &s_tot);
        int keep = cgt, eqr = eq_before;
        for (int i = i0; i < i1; ++i)
            if (key_of(i) == tau) { keep += eqr < need ? 1 : 0; ++eqr; }
        int off = block_exscan(keep, wsum, &s_tot);
        eqr = eq_before;
        const bool renorm = use_topk || a.top_p < 1.0f;
        for (int i = i0; i < i1; ++i) {
            const uint32_t key = key_of(i);
            bool take = key > tau;
            if (key == tau) { take = eqr < need; ++eqr; }
            if (take) { sidx[off] = i; sval[off] = renorm ? prob(i) / Z2 : prob(i); ++off; }
        }
        __syncthreads();
        n = s_tot;
    }
    // ---- the beam most probable entries with probability > 0, most probable first (:1990-2005) ----
    for (int i = tid; i < n; i += kSampleThreads) {
        const float p = sval[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float q = sval[j];
            rank += (q > p || (q == p && j < i)) ? 1 : 0;
        }
        sord[rank] = i;
    }
    __syncthreads();
    if (tid == 0) {
        int cnt = 0;
        for (int r = 0; r < n && r < a.beam; ++r) {
            const int i = sord[r];
            if (!(sval[i] > 0.0f)) break;
            a.cand_prob[(size_t)blockIdx.x * a.beam + cnt] = sval[i];
            a.cand_tok[(size_t)blockIdx.x * a.beam + cnt] = sidx[i];
            ++cnt;
        }
        a.cand_cnt[blockIdx.x] = cnt;
    }
}

// compute_logprobs (:919-944): one block per position, logprob of tokens[pos] under logits[pos]:
//   logit - max - log(sum exp(logit - max)); the -20 of an out-of-vocabulary token id (:933-936)
__global__ void logprob_rows_kernel(const float* logits, int V, const int* tokens, float* out) {
    __shared__ float red[32];
    const float* row = logits + (size_t)blockIdx.x * V;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, row[i]);
    mx = block_max_256(mx, red);
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(row[i] - mx);
    const float tot = block_sum_256(s, red);
    if (threadIdx.x == 0) {
        const int tok = tokens[blockIdx.x];
        out[blockIdx.x] = (tok < 0 || tok >= V) ? -20.0f : row[tok] - mx - logf(tot);
    }
}

}  // namespace tib
