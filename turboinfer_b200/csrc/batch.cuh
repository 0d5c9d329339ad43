// batch.cuh -- the kernels that turn the tensor-core GEMM path into a batched decode step: B independent sequences
// advance one token each per step, InferenceEngine::generate_batch (src/model/inference_engine.cpp:804-828, a sequential
// loop of generate calls in the reference) run in lockstep.  The weights are then read ONCE per step for all B rows
// (tcgen05 GEMM, gemm_tc.cuh); every sequence keeps its own pages of the KV cache.
//   rope_kv_batch_kernel   apply_rope on q, k (:1510-1624, per head) + KVCache append (:78-160), row b -> sequence b
//   (attention)            attn_partial_kernel / attn_combine_kernel of mega.cuh with blockIdx.z = sequence
//   argmax_rows_kernel     greedy sample_next_token (:1554-1673, top_k = 1: first maximum) per row, token fed back on device
#pragma once
#include "prefill.cuh"

namespace tib {

// qkv[B][3H] (q | k | v) -> RoPE on q (in place) and k; k / v rows appended at position *pos_ptr of sequence b, whose
// page table is tables[b * pages_per_seq ..]
__global__ void rope_kv_batch_kernel(float* qkv, int H, int rope_dim, const float* inv_freq, const int* pos_ptr, float* k_pool, float* v_pool,
                                     const int* tables, int pages_per_seq, int page_tokens) {
    const int b = blockIdx.x;
    const int pos = *pos_ptr;
    float* row = qkv + (size_t)b * 3 * H;
    const int page = tables[(size_t)b * pages_per_seq + pos / page_tokens];
    const size_t kvoff = ((size_t)page * page_tokens + (pos % page_tokens)) * H;
    for (int p = threadIdx.x; p < H / 2; p += blockDim.x) {
        const int d = 2 * p;
        float q0 = row[d], q1 = row[d + 1], k0 = row[H + d], k1 = row[H + d + 1];
        if (rope_dim > 0) {
            float sn, cs;
            sincosf((float)pos * inv_freq[(d % rope_dim) >> 1], &sn, &cs);
            const float a = __fsub_rn(__fmul_rn(q0, cs), __fmul_rn(q1, sn)), c = __fadd_rn(__fmul_rn(q0, sn), __fmul_rn(q1, cs));   // :1584-1585
            const float e = __fsub_rn(__fmul_rn(k0, cs), __fmul_rn(k1, sn)), f = __fadd_rn(__fmul_rn(k0, sn), __fmul_rn(k1, cs));
            q0 = a; q1 = c; k0 = e; k1 = f;
            row[d] = q0;
            row[d + 1] = q1;
        }
        *reinterpret_cast<float2*>(k_pool + kvoff + d) = make_float2(k0, k1);
        *reinterpret_cast<float2*>(v_pool + kvoff + d) = make_float2(row[2 * H + d], row[2 * H + d + 1]);
    }
}

// One block per row: the FIRST maximum of logits[b][0..V) (std::sort descending + [0] of the reference keeps the first of
// equal values only by accident of its sort; the oracle and every engine here define greedy as "first maximum").
// tokens[b] feeds the next step's embedding lookup; out[b * out_stride + *step_ptr] is the history the host reads.
__global__ void argmax_rows_kernel(const float* logits, int V, int* tokens, int* out, int out_stride, const int* step_ptr) {
    __shared__ unsigned long long best[32];
    const int b = blockIdx.x;
    const float* row = logits + (size_t)b * V;
    unsigned long long key = 0ull;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const unsigned long long k = argmax_pack(row[i], i);
        key = k > key ? k : key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) key = best[i] > key ? best[i] : key;
        const int tok = 0x7FFFFFFF - (int)(uint32_t)(key & 0xFFFFFFFFull);
        tokens[b] = tok;
        const int step = *step_ptr;
        if (out && step < out_stride) out[(size_t)b * out_stride + step] = tok;
    }
}

// end of a batched step: every sequence is one token longer; `sampled` steps also advance the output column
__global__ void batch_advance_kernel(int* pos_step, int sampled) {
    pos_step[0] += 1;
    if (sampled) pos_step[1] += 1;
}

}  // namespace tib
