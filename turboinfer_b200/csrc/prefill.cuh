// prefill.cuh -- the kernels around the tensor-core GEMM (gemm_tc.cuh) that turn it into a batched forward pass over
// M prompt tokens: InferenceEngine::forward_pass (src/model/inference_engine.cpp:1429-1491) -> TransformerLayer::forward
// (:203-233) with the intended dataflow of SURVEY.md 8c level B, but M rows at a time instead of one.
//   rmsnorm_digits_kernel   rms_norm (:1452-1508) fused with the fixed-point conversion that feeds the GEMM
//   rope_kv_kernel          apply_rope on q and k (:1510-1624, per head) + KVCache append (:78-160) for M positions
//   causal_attention_kernel multi_head_attention with a causal mask over the paged cache (:1149-1252, :348-362)
//   swiglu_rows_kernel      multiply(up, silu(gate)) on interleaved (gate, up) columns (:389-391)
// All fp32, full-precision expf / division like the decode path.
#pragma once
#include "gemm_tc.cuh"

namespace tib {

// X[M][K] fp32 -> (optional RMSNorm with weight w) -> digit planes [3][m_pad][k_pad] + per-row s_x, sum(xf).
// One block of 256 threads per row; the row is cached in shared memory between the passes (K <= 16384).
__global__ void rmsnorm_digits_kernel(const float* x, const float* w, float eps, int M, int K, int m_pad, int k_pad, int8_t* planes,
                                      float* sx_out, long long* sxf_out) {
    extern __shared__ float rowbuf[];
    __shared__ float red[32];
    __shared__ long long redl[32];
    const int row = blockIdx.x, tid = threadIdx.x, nw = blockDim.x >> 5;
    const float* xr = x + (size_t)row * K;
    float ss = 0.f;
    for (int k = tid; k < K; k += blockDim.x) {
        const float v = xr[k];
        rowbuf[k] = v;
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < nw; ++i) tot += red[i];
    __syncthreads();
    const float rms = w ? sqrtf(tot / (float)K + eps) : 1.f;   // :1501
    float amax = 0.f;
    for (int k = tid; k < K; k += blockDim.x) {
        float v = rowbuf[k];
        if (w) v = (v / rms) * w[k];                            // :1504-1506
        rowbuf[k] = v;
        amax = fmaxf(amax, fabsf(v));
    }
    amax = warp_max(amax);
    if ((tid & 31) == 0) red[tid >> 5] = amax;
    __syncthreads();
    amax = 0.f;
    for (int i = 0; i < nw; ++i) amax = fmaxf(amax, red[i]);
    const bool finite = amax > 0.f && amax < INFINITY;
    const float inv_s = finite ? kXQMax / amax : 0.f;
    const float s_x = finite ? amax / kXQMax : 0.f;
    long long sxf = 0;
    for (int k = tid; k < k_pad; k += blockDim.x) {
        const int f = k < K ? __float2int_rn(rowbuf[k] * inv_s) : 0;
        sxf += f;
        planes[((size_t)0 * m_pad + row) * k_pad + k] = (int8_t)(f & 0xFF);
        planes[((size_t)1 * m_pad + row) * k_pad + k] = (int8_t)((f >> 8) & 0xFF);
        planes[((size_t)2 * m_pad + row) * k_pad + k] = (int8_t)((f >> 16) & 0xFF);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sxf += __shfl_xor_sync(0xffffffffu, sxf, o);
    if ((tid & 31) == 0) redl[tid >> 5] = sxf;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int i = 0; i < nw; ++i) t += redl[i];
        sx_out[row] = s_x;
        sxf_out[row] = t;
    }
}

// qkv[M][3H] (q | k | v) -> RoPE on q (in place) and k, k / v rows appended to the paged cache at positions pos0 + m
__global__ void rope_kv_kernel(float* qkv, int M, int H, int rope_dim, const float* inv_freq, int pos0, float* k_pool, float* v_pool,
                               const int* page_table, int page_tokens) {
    const int m = blockIdx.x;
    const int pos = pos0 + m;
    float* row = qkv + (size_t)m * 3 * H;
    const int page = page_table[pos / page_tokens];
    const size_t kvoff = ((size_t)page * page_tokens + (pos % page_tokens)) * H;
    for (int p = threadIdx.x; p < H / 2; p += blockDim.x) {
        const int d = 2 * p;
        float q0 = row[d], q1 = row[d + 1], k0 = row[H + d], k1 = row[H + d + 1];
        if (rope_dim > 0) {
            float sn, cs;
            sincosf((float)pos * inv_freq[(d % rope_dim) >> 1], &sn, &cs);
            const float a = __fsub_rn(__fmul_rn(q0, cs), __fmul_rn(q1, sn)), b = __fadd_rn(__fmul_rn(q0, sn), __fmul_rn(q1, cs));   // :1584-1585
            const float c = __fsub_rn(__fmul_rn(k0, cs), __fmul_rn(k1, sn)), e = __fadd_rn(__fmul_rn(k0, sn), __fmul_rn(k1, cs));
            q0 = a; q1 = b; k0 = c; k1 = e;
            row[d] = q0;
            row[d + 1] = q1;
        }
        *reinterpret_cast<float2*>(k_pool + kvoff + d) = make_float2(k0, k1);
        *reinterpret_cast<float2*>(v_pool + kvoff + d) = make_float2(row[2 * H + d], row[2 * H + d + 1]);
    }
}

// Causal multi-head attention for M queries at positions pos0 .. pos0 + M - 1 over the paged cache (which already holds
// their own K / V rows).  grid (heads, ceil(M / 32)); 8 warps, 4 queries per warp; a lane owns 4 dims (D <= 128).
// Key tiles of 32 tokens are staged in shared memory once per CTA and shared by its 32 queries; the softmax is online
// (max-subtracted, expf), one key at a time per query, sums in key order -- the order of attention_fast_incremental.
constexpr int kPfQ = 32, kPfKT = 32, kPfThreads = 256;
__global__ void __launch_bounds__(kPfThreads) causal_attention_kernel(const float* qkv, int M, int H, int D, float scale, int pos0,
                                                                        const float* k_pool, const float* v_pool, const int* page_table,
                                                                        int page_tokens, float* out) {
    __shared__ __align__(16) float ks[kPfKT][128];
    __shared__ __align__(16) float vs[kPfKT][128];
    const int h = blockIdx.x, qb = blockIdx.y * kPfQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hoff = h * D;
    const bool lane_on = 4 * lane < D;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 q[4], o[4];
    float mrun[4], lrun[4];
    int qpos[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = qb + warp * 4 + i;
        qpos[i] = m < M ? pos0 + m : -1;
        q[i] = (m < M && lane_on) ? *reinterpret_cast<const float4*>(qkv + (size_t)m * 3 * H + hoff + 4 * lane) : zero4;
        o[i] = zero4;
        mrun[i] = -INFINITY;
        lrun[i] = 0.f;
    }
    const int last_pos = pos0 + min(M, qb + kPfQ) - 1;   // keys 0 .. last_pos can matter to this CTA
    for (int t0 = 0; t0 <= last_pos; t0 += kPfKT) {
        __syncthreads();
        for (int i = threadIdx.x; i < kPfKT * 32; i += kPfThreads) {
            const int tt = i >> 5, l = i & 31, t = t0 + tt;
            float4 kv = zero4, vv = zero4;
            if (t <= last_pos && 4 * l < D) {
                const size_t off = ((size_t)page_table[t / page_tokens] * page_tokens + (t % page_tokens)) * H + hoff + 4 * l;
                kv = *reinterpret_cast<const float4*>(k_pool + off);
                vv = *reinterpret_cast<const float4*>(v_pool + off);
            }
            *reinterpret_cast<float4*>(&ks[tt][4 * l]) = kv;
            *reinterpret_cast<float4*>(&vs[tt][4 * l]) = vv;
        }
        __syncthreads();
        const int nt = min(kPfKT, last_pos - t0 + 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (qpos[i] < t0) continue;   // warp-uniform: every key of the tile is in this query's future (or the query is padding)
            for (int tt = 0; tt < nt && t0 + tt <= qpos[i]; ++tt) {
                const float4 kv = *reinterpret_cast<const float4*>(&ks[tt][4 * lane]);
                float s = q[i].x * kv.x;
                s = fmaf(q[i].y, kv.y, s);
                s = fmaf(q[i].z, kv.z, s);
                s = fmaf(q[i].w, kv.w, s);
                s = warp_sum(s) * scale;
                const float mn = fmaxf(mrun[i], s);
                const float corr = expf(mrun[i] - mn), p = expf(s - mn);
                const float4 vv = *reinterpret_cast<const float4*>(&vs[tt][4 * lane]);
                lrun[i] = lrun[i] * corr + p;
                o[i].x = fmaf(p, vv.x, o[i].x * corr);
                o[i].y = fmaf(p, vv.y, o[i].y * corr);
                o[i].z = fmaf(p, vv.z, o[i].z * corr);
                o[i].w = fmaf(p, vv.w, o[i].w * corr);
                mrun[i] = mn;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = qb + warp * 4 + i;
        if (m < M && lane_on) {
            const float inv = 1.0f / lrun[i];
            *reinterpret_cast<float4*>(out + (size_t)m * H + hoff + 4 * lane) = make_float4(o[i].x * inv, o[i].y * inv, o[i].z * inv, o[i].w * inv);
        }
    }
}

// act[m][i] = up * silu(gate) from interleaved columns (gate_i, up_i) of gu[M][2I]
__global__ void swiglu_rows_kernel(const float* gu, float* act, size_t M, size_t I) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < M * I; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t m = idx / I, i = idx - m * I;
        const float g = gu[m * 2 * I + 2 * i], u = gu[m * 2 * I + 2 * i + 1];
        act[idx] = u * (g / (1.0f + expf(-g)));   // :918, :1729
    }
}
__global__ void relu_rows_kernel(const float* up, float* act, size_t n) {
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) act[idx] = fmaxf(up[idx], 0.f);
}
// x[m][:] = emb[token_m][:]
__global__ void embed_rows_kernel(const float* emb, const int* tokens, float* x, int H) {
    const int m = blockIdx.x;
    const float* e = emb + (size_t)tokens[m] * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) x[(size_t)m * H + i] = e[i];
}

}  // namespace tib
