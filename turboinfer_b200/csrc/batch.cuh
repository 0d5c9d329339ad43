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
    pdl_wait_prior_grid();       // (no-op unless launched with programmatic stream serialization: the lockstep step's graph)
    pdl_launch_dependents();
    const int b = blockIdx.x;
    const int pos = *pos_ptr;
    float* row = qkv + (size_t)b * 3 * H;
    const int page = tables[(size_t)b * pages_per_seq + pos / page_tokens];
    const size_t kvoff = ((size_t)page * page_tokens + (pos % page_tokens)) * H;
    for (int p = blockIdx.y * blockDim.x + threadIdx.x; p < H / 2; p += gridDim.y * blockDim.x) {   // grid (B, slices of the row)
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

// rmsnorm_digits_kernel for the batched decode: one block of 1024 threads per row, the row stays in registers between the
// passes (K <= 16384), the digit planes are written four bytes at a time.  gu != nullptr: the row is not read from x but
// computed on the fly as up * silu(gate) from the interleaved (gate_i, up_i) columns of gu[M][2K] (SwiGLU fused in front
// of the down projection, :918, :1729).  Same expressions as rmsnorm_digits_kernel (prefill.cuh).
constexpr int kDigitsThreads = 1024, kDigitsVecs = 4;   // 4 float4 per thread
// tile_layout = 1: planes in the 32-row GEMM's tile images (xtile_offset); 0: [3][m_pad][k_pad] row-major for the 128-row GEMM
__global__ void __launch_bounds__(kDigitsThreads) rmsnorm_digits_small_kernel(const float* x, const float* gu, const float* w, float eps, int K,
                                                                               int m_pad, int k_pad, int8_t* planes, float* sx_out, long long* sxf_out,
                                                                               int tile_layout) {
    pdl_wait_prior_grid();       // (no-op unless launched with programmatic stream serialization: the lockstep step's graph)
    pdl_launch_dependents();
    __shared__ float red[32];
    __shared__ long long redl[32];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float4 v[kDigitsVecs];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kDigitsVecs; ++i) {
        const int k = 4 * (tid + i * kDigitsThreads);
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
            if (gu) {
                const float4 a = *reinterpret_cast<const float4*>(gu + (size_t)row * 2 * K + 2 * k);
                const float4 b = *reinterpret_cast<const float4*>(gu + (size_t)row * 2 * K + 2 * k + 4);
                v[i] = make_float4(a.y * (a.x / (1.0f + expf(-a.x))), a.w * (a.z / (1.0f + expf(-a.z))), b.y * (b.x / (1.0f + expf(-b.x))),
                                   b.w * (b.z / (1.0f + expf(-b.z))));
            } else {
                v[i] = *reinterpret_cast<const float4*>(x + (size_t)row * K + k);
            }
        }
        ss = fmaf(v[i].x, v[i].x, ss); ss = fmaf(v[i].y, v[i].y, ss); ss = fmaf(v[i].z, v[i].z, ss); ss = fmaf(v[i].w, v[i].w, ss);
    }
    float amax = 0.f;
    if (w) {
        ss = warp_sum(ss);
        if (lane == 0) red[wid] = ss;
        __syncthreads();
        float tot = 0.f;
        for (int i = 0; i < kDigitsThreads / 32; ++i) tot += red[i];
        __syncthreads();
        const float rms = sqrtf(tot / (float)K + eps);   // :1501
#pragma unroll
        for (int i = 0; i < kDigitsVecs; ++i) {
            const int k = 4 * (tid + i * kDigitsThreads);
            if (k < K) {
                const float4 ww = *reinterpret_cast<const float4*>(w + k);
                v[i] = make_float4((v[i].x / rms) * ww.x, (v[i].y / rms) * ww.y, (v[i].z / rms) * ww.z, (v[i].w / rms) * ww.w);   // :1504-1506
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kDigitsVecs; ++i) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
    amax = warp_max(amax);
    if (lane == 0) red[wid] = amax;
    __syncthreads();
    amax = 0.f;
    for (int i = 0; i < kDigitsThreads / 32; ++i) amax = fmaxf(amax, red[i]);
    const bool finite = amax > 0.f && amax < INFINITY;
    const float inv_s = finite ? kXQMax / amax : 0.f;
    const float s_x = finite ? amax / kXQMax : 0.f;
    long long sxf = 0;
#pragma unroll
    for (int i = 0; i < kDigitsVecs; ++i) {
        const int k = 4 * (tid + i * kDigitsThreads);
        if (k < k_pad) {   // columns K .. k_pad - 1 are zero
            const int f0 = __float2int_rn(v[i].x * inv_s), f1 = __float2int_rn(v[i].y * inv_s), f2 = __float2int_rn(v[i].z * inv_s),
                      f3 = __float2int_rn(v[i].w * inv_s);
            sxf += (long long)((f0 + f1) + (f2 + f3));
            // signed digits: the bytes of f + 0x808080, xor 0x80 (every digit in [-128, 127])
            const int u0 = f0 + 0x808080, u1 = f1 + 0x808080, u2 = f2 + 0x808080, u3 = f3 + 0x808080;
            const uint32_t lo01 = __byte_perm(u0, u1, 0x5140), lo23 = __byte_perm(u2, u3, 0x5140);
            const uint32_t hi01 = __byte_perm(u0, u1, 0x0062), hi23 = __byte_perm(u2, u3, 0x0062);
            const uint32_t d0 = __byte_perm(lo01, lo23, 0x5410) ^ 0x80808080u, d1 = __byte_perm(lo01, lo23, 0x7632) ^ 0x80808080u,
                           d2 = __byte_perm(hi01, hi23, 0x5410) ^ 0x80808080u;
            if (tile_layout) {   // the swizzled tile images the 32-row GEMM copies in bulk (gemm_tc.cuh xtile_offset)
                *reinterpret_cast<uint32_t*>(planes + xtile_offset(0, row, k)) = d0;
                *reinterpret_cast<uint32_t*>(planes + xtile_offset(1, row, k)) = d1;
                *reinterpret_cast<uint32_t*>(planes + xtile_offset(2, row, k)) = d2;
            } else {
                *reinterpret_cast<uint32_t*>(planes + ((size_t)0 * m_pad + row) * k_pad + k) = d0;
                *reinterpret_cast<uint32_t*>(planes + ((size_t)1 * m_pad + row) * k_pad + k) = d1;
                *reinterpret_cast<uint32_t*>(planes + ((size_t)2 * m_pad + row) * k_pad + k) = d2;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sxf += __shfl_xor_sync(0xffffffffu, sxf, o);
    if (lane == 0) redl[wid] = sxf;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int i = 0; i < kDigitsThreads / 32; ++i) t += redl[i];
        sx_out[row] = s_x;
        sxf_out[row] = t;
    }
}

// One block per row: the FIRST maximum of logits[b][0..V) (std::sort descending + [0] of the reference keeps the first of
// equal values only by accident of its sort; the oracle and every engine here define greedy as "first maximum").
// tokens[b] feeds the next step's embedding lookup.  Ragged batches: sequence b's prompt has lens[b] tokens, all sequences
// are left-aligned and advance in lockstep, so at step s (= tokens already in every cache, *pos_ptr) sequence b has produced
// its output number s - (lens[b] - 1): that is the column of out[b][...] the pick goes to (nothing is written before the
// sequence's prompt has ended or after its n_new-th token).
__global__ void argmax_rows_kernel(const float* logits, int V, int* tokens, int* out, int out_stride, const int* pos_ptr, const int* lens) {
    pdl_wait_prior_grid();       // (no-op unless launched with programmatic stream serialization: the lockstep step's graph)
    pdl_launch_dependents();
    __shared__ unsigned long long best[32];
    const int b = blockIdx.x;
    const float* row = logits + (size_t)b * V;
    unsigned long long key = 0ull;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        // NaN never wins and -inf never beats "nothing yet" (key 0 -> token 0): the policy of the GEMV epilogue (`y > best`)
        const float v = row[i];
        const unsigned long long k = v > -INFINITY ? argmax_pack(v, i) : 0ull;
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
        int tok = key == 0ull ? 0 : 0x7FFFFFFF - (int)(uint32_t)(key & 0xFFFFFFFFull);
        tok = min(max(tok, 0), V - 1);
        const int s = *pos_ptr, col = s - (lens[b] - 1);
        if (col >= 0) tokens[b] = tok;   // still inside the prompt: the next token comes from the prompt (batch_feed_kernel)
        if (out && col >= 0 && col < out_stride) out[(size_t)b * out_stride + col] = tok;
    }
}

// start of a batched step: sequence b is fed its prompt token of this position while its prompt lasts, its own last pick after
// that (prompts: [max_len][B], column b = sequence b, padded)
__global__ void batch_feed_kernel(const int* prompts, const int* lens, const int* pos_ptr, int B, int* tokens) {
    pdl_wait_prior_grid();       // (no-op unless launched with programmatic stream serialization: the lockstep step's graph)
    pdl_launch_dependents();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int s = *pos_ptr;
    if (s < lens[b]) tokens[b] = prompts[(size_t)s * B + b];
}

// Beam search forks a sequence by copying its page TABLE; only the last, partly filled page of a shared prefix has to be copied
// before two beams append to it.  pairs[2 * i] = source page, pairs[2 * i + 1] = destination page; grid (pairs, layers, 2: K / V),
// `elems` = the filled part of the page (tokens * row width), a multiple of 4.
__global__ void kv_pages_copy_kernel(float* const* k_pools, float* const* v_pools, const int* pairs, size_t page_elems, size_t elems) {
    float* pool = (blockIdx.z == 0 ? k_pools : v_pools)[blockIdx.y];
    const float4* src = reinterpret_cast<const float4*>(pool + (size_t)pairs[2 * blockIdx.x] * page_elems);
    float4* dst = reinterpret_cast<float4*>(pool + (size_t)pairs[2 * blockIdx.x + 1] * page_elems);
    for (size_t i = threadIdx.x; i < elems / 4; i += blockDim.x) dst[i] = src[i];
}

// end of a batched step: every sequence is one token longer; `sampled` steps also advance the output column
__global__ void batch_advance_kernel(int* pos_step, int sampled) {
    pdl_wait_prior_grid();       // (no-op unless launched with programmatic stream serialization: the lockstep step's graph)
    pdl_launch_dependents();
    pos_step[0] += 1;
    if (sampled) pos_step[1] += 1;
}

}  // namespace tib
