// prefill.cuh -- the kernels around the tensor-core GEMM (gemm_tc.cuh) that turn it into a batched forward pass over
// M prompt tokens: InferenceEngine::forward_pass (src/model/inference_engine.cpp:1429-1491) -> TransformerLayer::forward
// (:203-233) with the intended dataflow of SURVEY.md 8c level B, but M rows at a time instead of one.
//   rmsnorm_digits_kernel   rms_norm (:1452-1508) fused with the fixed-point conversion that feeds the GEMM
//   rope_kv_kernel          apply_rope on q and k (:1510-1624, per head) + KVCache append (:78-160) for M positions
//   causal_attention_kernel multi_head_attention with a causal mask over the paged cache (:1149-1252, :348-362)
//   swiglu_rows_kernel      multiply(up, silu(gate)) on interleaved (gate, up) columns (:389-391)
// All fp32, full-precision expf / division like the decode path.
#pragma once
#include <cuda_fp16.h>
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
        // signed digits, f = d2*65536 + d1*256 + d0 with every digit in [-128, 127]: the bytes of f + 0x808080, xor 0x80
        const int u = f + 0x808080;
        planes[((size_t)0 * m_pad + row) * k_pad + k] = (int8_t)((u & 0xFF) ^ 0x80);
        planes[((size_t)1 * m_pad + row) * k_pad + k] = (int8_t)(((u >> 8) & 0xFF) ^ 0x80);
        planes[((size_t)2 * m_pad + row) * k_pad + k] = (int8_t)(((u >> 16) & 0xFF) ^ 0x80);
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
// their own K / V rows).  fp32 on the CUDA cores, organised as two register-tiled GEMMs per key tile (no per-pair
// shuffles): a CTA owns 128 queries of one head; per tile of 64 keys
//   S[128 x 64] = Q K^T   thread (ty, tx) computes 8 queries x 4 keys, reading Q and K TRANSPOSED from shared memory
//                         ([d][query], [d][key]: two broadcast LDS.128 + one LDS.128 per 32 FMAs)
//   online softmax        per query row over the tile (max-subtracted, expf), row statistics shared by the 16 threads of a
//                         row with shuffles once per tile; O and the running sum are rescaled by exp(m_old - m_new)
//   O[128 x 128] += P V   the same thread owns the same 8 query rows x 8 dims (64 accumulators), P read transposed
// Keys in a query's future get probability 0; tiles entirely in the future are skipped.  D <= 128 (padded with zeros).
constexpr int kPfQ = 128, kPfKT = 64, kPfThreads = 256, kPfD = 128;
constexpr size_t kPfSmemBytes = ((size_t)kPfD * kPfQ + (size_t)kPfD * kPfKT + (size_t)kPfKT * kPfD + (size_t)kPfKT * kPfQ) * sizeof(float);   // 160 KiB
__global__ void __launch_bounds__(kPfThreads, 1) causal_attention_kernel(const float* qkv, int M, int H, int D, float scale, int pos0,
                                                                           const float* k_pool, const float* v_pool, const int* page_table,
                                                                           int page_tokens, float* out) {
    extern __shared__ __align__(16) float pf_smem[];
    float* Qt = pf_smem;                       // [128 d][128 q]
    float* Kt = Qt + kPfD * kPfQ;              // [128 d][64 k]
    float* Vs = Kt + kPfD * kPfKT;             // [64 k][128 d]
    float* Pt = Vs + kPfKT * kPfD;             // [64 k][128 q]
    const int h = blockIdx.x, qb = (int)(gridDim.y - 1 - blockIdx.y) * kPfQ;   // latest queries (most key tiles) first: no long tail
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int hoff = h * D;
    const int nq = min(kPfQ, M - qb);          // live queries of this CTA
    // Q tile, transposed.  Consecutive lanes take consecutive queries, so the four scattered stores of a float4 are
    // conflict-free (the 16-byte global reads of neighbouring d meet again in L1).
    for (int i = tid; i < kPfQ * (kPfD / 4); i += kPfThreads) {
        const int q = i & (kPfQ - 1), d4 = (i >> 7) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < nq && d4 < D) v = *reinterpret_cast<const float4*>(qkv + (size_t)(qb + q) * 3 * H + hoff + d4);
        Qt[(d4 + 0) * kPfQ + q] = v.x; Qt[(d4 + 1) * kPfQ + q] = v.y; Qt[(d4 + 2) * kPfQ + q] = v.z; Qt[(d4 + 3) * kPfQ + q] = v.w;
    }
    f32x2 o2[8][4];   // 8 queries x 8 dims, as pairs of dims
    float mrun[8], lrun[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mrun[i] = -INFINITY;
        lrun[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o2[i][j] = pack2(0.f, 0.f);
    }
    const int last_pos = pos0 + qb + nq - 1;   // keys 0 .. last_pos can matter to this CTA
    for (int t0 = 0; t0 <= last_pos; t0 += kPfKT) {
        __syncthreads();                       // the previous tile's Kt / Vs / Pt are no longer read
        for (int i = tid; i < kPfKT * (kPfD / 4); i += kPfThreads) {   // K transposed: consecutive lanes = consecutive keys
            const int kk = i & (kPfKT - 1), d4 = (i >> 6) * 4, t = t0 + kk;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t <= last_pos && d4 < D)
                kv = *reinterpret_cast<const float4*>(k_pool + ((size_t)page_table[t / page_tokens] * page_tokens + (t % page_tokens)) * H + hoff + d4);
            Kt[(d4 + 0) * kPfKT + kk] = kv.x; Kt[(d4 + 1) * kPfKT + kk] = kv.y; Kt[(d4 + 2) * kPfKT + kk] = kv.z; Kt[(d4 + 3) * kPfKT + kk] = kv.w;
        }
        for (int i = tid; i < kPfKT * (kPfD / 4); i += kPfThreads) {   // V as it is: consecutive lanes = consecutive dims
            const int kk = i >> 5, d4 = (i & 31) * 4, t = t0 + kk;
            float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t <= last_pos && d4 < D)
                vv = *reinterpret_cast<const float4*>(v_pool + ((size_t)page_table[t / page_tokens] * page_tokens + (t % page_tokens)) * H + hoff + d4);
            *reinterpret_cast<float4*>(Vs + kk * kPfD + d4) = vv;
        }
        __syncthreads();
        // S = Q K^T for 8 queries x 4 keys
        // (packed fp32x2 FMAs, SASS FFMA2: two fma.rn per instruction, bit-identical to the scalar form)
        f32x2 sacc2[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) sacc2[i][0] = sacc2[i][1] = pack2(0.f, 0.f);
#pragma unroll 4
        for (int d = 0; d < kPfD; ++d) {
            const float4 qa = *reinterpret_cast<const float4*>(Qt + d * kPfQ + 8 * ty), qb4 = *reinterpret_cast<const float4*>(Qt + d * kPfQ + 8 * ty + 4);
            const float4 kb = *reinterpret_cast<const float4*>(Kt + d * kPfKT + 4 * tx);
            const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb4.x, qb4.y, qb4.z, qb4.w};
            const f32x2 k01 = pack2(kb.x, kb.y), k23 = pack2(kb.z, kb.w);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const f32x2 qq = pack2(qv[i], qv[i]);
                sacc2[i][0] = fma2(qq, k01, sacc2[i][0]);
                sacc2[i][1] = fma2(qq, k23, sacc2[i][1]);
            }
        }
        float sacc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { unpack2(sacc2[i][0], sacc[i][0], sacc[i][1]); unpack2(sacc2[i][1], sacc[i][2], sacc[i][3]); }
        // online softmax per query row; the 16 threads tx = 0..15 of a row are the 16 lanes of a half-warp
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int qpos = pos0 + qb + 8 * ty + i;
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = t0 + 4 * tx + j;
                sacc[i][j] = (t <= qpos && 8 * ty + i < nq) ? sacc[i][j] * scale : -INFINITY;
                mx = fmaxf(mx, sacc[i][j]);
            }
#pragma unroll
            for (int o2 = 8; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
            const float mn = fmaxf(mrun[i], mx);
            float rs = 0.f, p[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                p[j] = mn == -INFINITY ? 0.f : expf(sacc[i][j] - mn);   // expf(-inf) = 0 for masked keys
                rs += p[j];
            }
#pragma unroll
            for (int o2 = 8; o2 > 0; o2 >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o2);
            const float corr = mn == -INFINITY ? 1.f : expf(mrun[i] - mn);
            lrun[i] = lrun[i] * corr + rs;
            mrun[i] = mn;
            const f32x2 cc = pack2(corr, corr);
#pragma unroll
            for (int j = 0; j < 4; ++j) o2[i][j] = mul2(o2[i][j], cc);
#pragma unroll
            for (int j = 0; j < 4; ++j) Pt[(4 * tx + j) * kPfQ + 8 * ty + i] = p[j];
        }
        __syncthreads();
        // O += P V for the same 8 queries x dims 8 tx .. 8 tx + 7
        const int nk = min(kPfKT, last_pos - t0 + 1);
        for (int kk = 0; kk < nk; ++kk) {
            const float4 pa = *reinterpret_cast<const float4*>(Pt + kk * kPfQ + 8 * ty), pb = *reinterpret_cast<const float4*>(Pt + kk * kPfQ + 8 * ty + 4);
            const float4 va = *reinterpret_cast<const float4*>(Vs + kk * kPfD + 8 * tx), vb = *reinterpret_cast<const float4*>(Vs + kk * kPfD + 8 * tx + 4);
            const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
            const f32x2 v2[4] = {pack2(va.x, va.y), pack2(va.z, va.w), pack2(vb.x, vb.y), pack2(vb.z, vb.w)};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const f32x2 pp = pack2(pv[i], pv[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) o2[i][j] = fma2(pp, v2[j], o2[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int q = 8 * ty + i;
        if (q < nq && 8 * tx < D) {
            const float inv = 1.0f / lrun[i];
            float o[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) unpack2(o2[i][j], o[2 * j], o[2 * j + 1]);
            float* dst = out + (size_t)(qb + q) * H + hoff + 8 * tx;
            *reinterpret_cast<float4*>(dst) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
            if (8 * tx + 4 < D) *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4] * inv, o[5] * inv, o[6] * inv, o[7] * inv);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// The same attention on the tensor cores (head dimension 64 or 128): S = Q K^T and O += P V as mma.sync m16n8k8 TF32 MMAs
// with every fp32 operand split into two TF32 terms (x = hi + lo, hi = tf32(x), lo = tf32(x - hi)) and the three products
// hi*hi + hi*lo + lo*hi accumulated in fp32 -- about 21 mantissa bits, i.e. the fp32 result to a few 1e-7 relative, at a
// third of the TF32 rate instead of the CUDA cores' fp32 rate.  Softmax (max-subtracted, expf, running rescale) stays fp32.
//   CTA = 8 warps = 128 queries of one head, a warp owns 16 query rows; per tile of 64 keys the CTA loads K and V once from the
//   paged cache, splits them and keeps hi / lo images in shared memory ([key][D + 4]: conflict-free fragment reads);
//   Q stays in shared memory as fp32 (A fragments read and split per k-step); P never leaves registers: the C fragment of S is used as the
//   A fragment of P V with the keys of a k-step permuted (k slot t <-> key 2t, slot t + 4 <-> key 2t + 1), V rows read to match.
// ---------------------------------------------------------------------------------------------------
constexpr int kTcQ = 128, kTcKT = 64, kTcThreads = 256;
template <int DH> constexpr size_t attn_tc_smem_bytes() { return ((size_t)4 * kTcKT + kTcQ) * (DH + 4) * sizeof(uint32_t); }
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int DH>
__global__ void __launch_bounds__(kTcThreads, 1) causal_attention_tc_kernel(const float* qkv, int M, int H, float scale, int pos0, const float* k_pool,
                                                                              const float* v_pool, const int* page_table, int page_tokens, float* out) {
    constexpr int LD = DH + 4, KS = DH / 8;   // row stride of the shared tiles; k-steps of Q K^T = n-tiles of P V
    extern __shared__ __align__(16) uint32_t tc_smem[];
    uint32_t* Khi = tc_smem;
    uint32_t* Klo = Khi + kTcKT * LD;
    uint32_t* Vhi = Klo + kTcKT * LD;
    uint32_t* Vlo = Vhi + kTcKT * LD;
    float* Qs = reinterpret_cast<float*>(Vlo + kTcKT * LD);                    // [128 q][LD]
    const int h = blockIdx.x, qb = (int)(gridDim.y - 1 - blockIdx.y) * kTcQ;   // latest queries (most key tiles) first
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int hoff = h * DH;
    const int r0 = qb + 16 * warp + g, r1 = r0 + 8;   // this thread's two query rows (within the M prompt rows)
    for (int i = tid; i < kTcQ * (DH / 4); i += kTcThreads) {
        const int q = i / (DH / 4), d4 = (i % (DH / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qb + q < M) v = *reinterpret_cast<const float4*>(qkv + (size_t)(qb + q) * 3 * H + hoff + d4);
        *reinterpret_cast<float4*>(Qs + q * LD + d4) = v;
    }
    const float* qrow = Qs + (16 * warp + g) * LD + t;   // A fragment of k-step ks: rows g / g + 8, dims 8 ks + t / + 4
    float o[KS][4];
#pragma unroll
    for (int n = 0; n < KS; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int last_pos = pos0 + min(qb + kTcQ, M) - 1;        // keys 0 .. last_pos can matter to this CTA
    const int warp_last = pos0 + min(qb + 16 * warp + 15, M - 1);   // ... and to this warp
    for (int t0 = 0; t0 <= last_pos; t0 += kTcKT) {
        __syncthreads();   // the previous tile is no longer read
        for (int i = tid; i < kTcKT * (DH / 4); i += kTcThreads) {
            const int kk = i / (DH / 4), d4 = (i % (DH / 4)) * 4, tk = t0 + kk;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (tk <= last_pos) {
                const size_t row = ((size_t)page_table[tk / page_tokens] * page_tokens + (tk % page_tokens)) * H + hoff + d4;
                kv = *reinterpret_cast<const float4*>(k_pool + row);
                vv = *reinterpret_cast<const float4*>(v_pool + row);
            }
            uint4 hi, lo;
            split_tf32(kv.x, hi.x, lo.x); split_tf32(kv.y, hi.y, lo.y); split_tf32(kv.z, hi.z, lo.z); split_tf32(kv.w, hi.w, lo.w);
            *reinterpret_cast<uint4*>(Khi + kk * LD + d4) = hi;
            *reinterpret_cast<uint4*>(Klo + kk * LD + d4) = lo;
            split_tf32(vv.x, hi.x, lo.x); split_tf32(vv.y, hi.y, lo.y); split_tf32(vv.z, hi.z, lo.z); split_tf32(vv.w, hi.w, lo.w);
            *reinterpret_cast<uint4*>(Vhi + kk * LD + d4) = hi;
            *reinterpret_cast<uint4*>(Vlo + kk * LD + d4) = lo;
        }
        __syncthreads();
        if (t0 > warp_last) continue;   // every key of the tile lies in this warp's future (the barriers above stay CTA-wide)
        // ---- S = Q K^T: 16 rows x 64 keys per warp ----
        float sc[kTcKT / 8][4];
#pragma unroll
        for (int j = 0; j < kTcKT / 8; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t ah[4], al[4];
            split_tf32(qrow[8 * ks], ah[0], al[0]);
            split_tf32(qrow[8 * LD + 8 * ks], ah[1], al[1]);
            split_tf32(qrow[8 * ks + 4], ah[2], al[2]);
            split_tf32(qrow[8 * LD + 8 * ks + 4], ah[3], al[3]);
            // the three products of an accumulator are issued a whole pass apart (8 independent MMAs between dependent ones)
            uint32_t bh0[kTcKT / 8], bh1[kTcKT / 8], bl0[kTcKT / 8], bl1[kTcKT / 8];
#pragma unroll
            for (int j = 0; j < kTcKT / 8; ++j) {
                const int off = (8 * j + g) * LD + 8 * ks + t;   // B fragment: (k = dim, n = key)
                bh0[j] = Khi[off]; bh1[j] = Khi[off + 4]; bl0[j] = Klo[off]; bl1[j] = Klo[off + 4];
            }
#pragma unroll
            for (int j = 0; j < kTcKT / 8; ++j) mma_tf32(sc[j], al, bh0[j], bh1[j]);
#pragma unroll
            for (int j = 0; j < kTcKT / 8; ++j) mma_tf32(sc[j], ah, bl0[j], bl1[j]);
#pragma unroll
            for (int j = 0; j < kTcKT / 8; ++j) mma_tf32(sc[j], ah, bh0[j], bh1[j]);
        }
        // ---- online softmax over the tile (rows r0, r1; this thread holds keys 8 j + 2 t, + 1 of every j) ----
        const int qp0 = pos0 + r0, qp1 = pos0 + r1;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < kTcKT / 8; ++j) {
            const int key = t0 + 8 * j + 2 * t;
            sc[j][0] = (key <= qp0 && r0 < M) ? sc[j][0] * scale : -INFINITY;
            sc[j][1] = (key + 1 <= qp0 && r0 < M) ? sc[j][1] * scale : -INFINITY;
            sc[j][2] = (key <= qp1 && r1 < M) ? sc[j][2] * scale : -INFINITY;
            sc[j][3] = (key + 1 <= qp1 && r1 < M) ? sc[j][3] * scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int j = 0; j < kTcKT / 8; ++j) {
            sc[j][0] = mn0 == -INFINITY ? 0.f : expf(sc[j][0] - mn0);   // expf(-inf) = 0 for masked keys
            sc[j][1] = mn0 == -INFINITY ? 0.f : expf(sc[j][1] - mn0);
            sc[j][2] = mn1 == -INFINITY ? 0.f : expf(sc[j][2] - mn1);
            sc[j][3] = mn1 == -INFINITY ? 0.f : expf(sc[j][3] - mn1);
            rs0 += sc[j][0] + sc[j][1];
            rs1 += sc[j][2] + sc[j][3];
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        const float c0 = mn0 == -INFINITY ? 1.f : expf(m0 - mn0), c1 = mn1 == -INFINITY ? 1.f : expf(m1 - mn1);
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
        m0 = mn0;
        m1 = mn1;
#pragma unroll
        for (int n = 0; n < KS; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
        // ---- O += P V: k-step j = keys 8 j .. 8 j + 7, with k slot t <-> key 8 j + 2 t and slot t + 4 <-> key 8 j + 2 t + 1 ----
#pragma unroll
        for (int j = 0; j < kTcKT / 8; ++j) {
            uint32_t ph[4], pl[4];
            split_tf32(sc[j][0], ph[0], pl[0]);   // a0: (row g,     slot t)
            split_tf32(sc[j][2], ph[1], pl[1]);   // a1: (row g + 8, slot t)
            split_tf32(sc[j][1], ph[2], pl[2]);   // a2: (row g,     slot t + 4)
            split_tf32(sc[j][3], ph[3], pl[3]);   // a3: (row g + 8, slot t + 4)
            const int vrow = (8 * j + 2 * t) * LD + g;
#pragma unroll
            for (int n0 = 0; n0 < KS; n0 += 8) {   // 8 output tiles at a time: dependent MMAs a pass apart, 32 fragment registers
                uint32_t bh0[8], bh1[8], bl0[8], bl1[8];
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int off = vrow + 8 * (n0 + n);
                    bh0[n] = Vhi[off]; bh1[n] = Vhi[off + LD]; bl0[n] = Vlo[off]; bl1[n] = Vlo[off + LD];
                }
#pragma unroll
                for (int n = 0; n < 8; ++n) mma_tf32(o[n0 + n], pl, bh0[n], bh1[n]);
#pragma unroll
                for (int n = 0; n < 8; ++n) mma_tf32(o[n0 + n], ph, bl0[n], bl1[n]);
#pragma unroll
                for (int n = 0; n < 8; ++n) mma_tf32(o[n0 + n], ph, bh0[n], bh1[n]);
            }
        }
    }
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int n = 0; n < KS; ++n) {
        if (r0 < M) *reinterpret_cast<float2*>(out + (size_t)r0 * H + hoff + 8 * n + 2 * t) = make_float2(o[n][0] * i0, o[n][1] * i0);
        if (r1 < M) *reinterpret_cast<float2*>(out + (size_t)r1 * H + hoff + 8 * n + 2 * t) = make_float2(o[n][2] * i1, o[n][3] * i1);
    }
}

// ---------------------------------------------------------------------------------------------------
// The same attention with FP16 MMAs (mma.sync m16n8k16, twice the work per instruction of the TF32 form): every fp32 operand
// is split into two halves, x = hi + lo (hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits), and hi*hi + hi*lo + lo*hi is
// accumulated in fp32 -- the accuracy of the split-TF32 kernel above at half its instruction count.  Values are activations
// (|x| well inside fp16's range).  Q, K, V hi / lo images live in shared memory as halves ([row][D + 8]: conflict-free
// ldmatrix / LDS.32); K fragments come from ldmatrix, V fragments from ldmatrix.trans (the k index of P V is the key), and the
// C fragments of two adjacent S tiles ARE the A fragment of P V for 16 keys, so P never leaves registers.
// ---------------------------------------------------------------------------------------------------
template <int DH> constexpr size_t attn_h3_smem_bytes() { return ((size_t)2 * kTcQ + 4 * kTcKT) * (DH + 8) * sizeof(__half); }
__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
template <int DH>
__global__ void __launch_bounds__(kTcThreads, 1) causal_attention_h3_kernel(const float* qkv, int M, int H, float scale, int pos0, const float* k_pool,
                                                                              const float* v_pool, const int* page_table, int page_tokens, float* out) {
    constexpr int LD = DH + 8, KS = DH / 16, NT = DH / 8;   // row stride in halves; k-steps of Q K^T; n-tiles of P V
    extern __shared__ __align__(16) __half h3_smem[];
    __half* Qh = h3_smem;                 // [128][LD]
    __half* Ql = Qh + kTcQ * LD;
    __half* Kh = Ql + kTcQ * LD;          // [64][LD]
    __half* Kl = Kh + kTcKT * LD;
    __half* Vh = Kl + kTcKT * LD;
    __half* Vl = Vh + kTcKT * LD;
    const int h = blockIdx.x, qb = (int)(gridDim.y - 1 - blockIdx.y) * kTcQ;   // latest queries (most key tiles) first
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int hoff = h * DH;
    const int r0 = qb + 16 * warp + g, r1 = r0 + 8;
    auto store_split4 = [](const float4& v, __half* hi, __half* lo) {
        __half a[4], b[4];
        split_h(v.x, a[0], b[0]); split_h(v.y, a[1], b[1]); split_h(v.z, a[2], b[2]); split_h(v.w, a[3], b[3]);
        *reinterpret_cast<uint2*>(hi) = make_uint2(pack_h2(a[0], a[1]), pack_h2(a[2], a[3]));
        *reinterpret_cast<uint2*>(lo) = make_uint2(pack_h2(b[0], b[1]), pack_h2(b[2], b[3]));
    };
    for (int i = tid; i < kTcQ * (DH / 4); i += kTcThreads) {
        const int q = i / (DH / 4), d4 = (i % (DH / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qb + q < M) v = *reinterpret_cast<const float4*>(qkv + (size_t)(qb + q) * 3 * H + hoff + d4);
        store_split4(v, Qh + q * LD + d4, Ql + q * LD + d4);
    }
    float o[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int last_pos = pos0 + min(qb + kTcQ, M) - 1;
    const int warp_last = pos0 + min(qb + 16 * warp + 15, M - 1);
    // A fragments of Q: rows g / g + 8 of the warp's 16, halves (2t, 2t + 1) and (2t + 8, 2t + 9) of a 16-dim k-step
    const uint32_t* qh_row = reinterpret_cast<const uint32_t*>(Qh + (16 * warp + g) * LD) + t;
    const uint32_t* ql_row = reinterpret_cast<const uint32_t*>(Ql + (16 * warp + g) * LD) + t;
    // ldmatrix row addresses: lane -> (matrix = lane / 8, row = lane % 8)
    const int lm = lane >> 3, lr = lane & 7;
    const uint32_t k_ld = (uint32_t)(((lr + 8 * (lm >> 1)) * LD + 8 * (lm & 1)) * 2);   // K: matrices (keys 0-7 | dims 0-7, 8-15), (keys 8-15 | ...)
    const uint32_t v_ld = (uint32_t)(((lr + 8 * (lm & 1)) * LD + 8 * (lm >> 1)) * 2);   // V^T: matrices (keys 0-7, 8-15 | dims 0-7), (... | dims 8-15)
    const uint32_t kh_s = smem_u32(Kh), kl_s = smem_u32(Kl), vh_s = smem_u32(Vh), vl_s = smem_u32(Vl);
    for (int t0 = 0; t0 <= last_pos; t0 += kTcKT) {
        __syncthreads();
        for (int i = tid; i < kTcKT * (DH / 4); i += kTcThreads) {
            const int kk = i / (DH / 4), d4 = (i % (DH / 4)) * 4, tk = t0 + kk;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (tk <= last_pos) {
                const size_t row = ((size_t)page_table[tk / page_tokens] * page_tokens + (tk % page_tokens)) * H + hoff + d4;
                kv = *reinterpret_cast<const float4*>(k_pool + row);
                vv = *reinterpret_cast<const float4*>(v_pool + row);
            }
            store_split4(kv, Kh + kk * LD + d4, Kl + kk * LD + d4);
            store_split4(vv, Vh + kk * LD + d4, Vl + kk * LD + d4);
        }
        __syncthreads();
        if (t0 > warp_last) continue;
        // ---- S = Q K^T ----
        float sc[kTcKT / 8][4];
#pragma unroll
        for (int j = 0; j < kTcKT / 8; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const uint32_t ah[4] = {qh_row[8 * ks], qh_row[8 * ks + 4 * LD], qh_row[8 * ks + 4], qh_row[8 * ks + 4 * LD + 4]};   // (LD halves = LD / 2 words; 8 rows = 4 LD words)
            const uint32_t al[4] = {ql_row[8 * ks], ql_row[8 * ks + 4 * LD], ql_row[8 * ks + 4], ql_row[8 * ks + 4 * LD + 4]};
#pragma unroll
            for (int jp = 0; jp < kTcKT / 16; ++jp) {   // two key tiles per ldmatrix.x4: {b0, b1} of tile 2 jp, {b0, b1} of tile 2 jp + 1
                uint32_t bh[4], bl[4];
                const uint32_t off = (uint32_t)((16 * jp * LD + 16 * ks) * 2) + k_ld;
                ldsm_x4(bh, kh_s + off);
                ldsm_x4(bl, kl_s + off);
                mma_f16(sc[2 * jp], al, bh[0], bh[1]);
                mma_f16(sc[2 * jp + 1], al, bh[2], bh[3]);
                mma_f16(sc[2 * jp], ah, bl[0], bl[1]);
                mma_f16(sc[2 * jp + 1], ah, bl[2], bl[3]);
                mma_f16(sc[2 * jp], ah, bh[0], bh[1]);
                mma_f16(sc[2 * jp + 1], ah, bh[2], bh[3]);
            }
        }
        // ---- online softmax (as in the TF32 kernel) ----
        const int qp0 = pos0 + r0, qp1 = pos0 + r1;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < kTcKT / 8; ++j) {
            const int key = t0 + 8 * j + 2 * t;
            sc[j][0] = (key <= qp0 && r0 < M) ? sc[j][0] * scale : -INFINITY;
            sc[j][1] = (key + 1 <= qp0 && r0 < M) ? sc[j][1] * scale : -INFINITY;
            sc[j][2] = (key <= qp1 && r1 < M) ? sc[j][2] * scale : -INFINITY;
            sc[j][3] = (key + 1 <= qp1 && r1 < M) ? sc[j][3] * scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int j = 0; j < kTcKT / 8; ++j) {
            sc[j][0] = mn0 == -INFINITY ? 0.f : expf(sc[j][0] - mn0);
            sc[j][1] = mn0 == -INFINITY ? 0.f : expf(sc[j][1] - mn0);
            sc[j][2] = mn1 == -INFINITY ? 0.f : expf(sc[j][2] - mn1);
            sc[j][3] = mn1 == -INFINITY ? 0.f : expf(sc[j][3] - mn1);
            rs0 += sc[j][0] + sc[j][1];
            rs1 += sc[j][2] + sc[j][3];
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        const float c0 = mn0 == -INFINITY ? 1.f : expf(m0 - mn0), c1 = mn1 == -INFINITY ? 1.f : expf(m1 - mn1);
        l0 = l0 * c0 + rs0;
        l1 = l1 * c1 + rs1;
        m0 = mn0;
        m1 = mn1;
#pragma unroll
        for (int n = 0; n < NT; ++n) { o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1; }
        // ---- O += P V: k-step jp = keys 16 jp .. 16 jp + 15 = S tiles 2 jp and 2 jp + 1 ----
#pragma unroll
        for (int jp = 0; jp < kTcKT / 16; ++jp) {
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {   // a0: (row g, keys 2t, 2t+1) a1: (row g+8, ...) of tile 2 jp; a2, a3: the same of tile 2 jp + 1
                const float x = sc[2 * jp + (i >> 1)][2 * (i & 1)], y = sc[2 * jp + (i >> 1)][2 * (i & 1) + 1];
                __half xh, xl, yh, yl;
                split_h(x, xh, xl);
                split_h(y, yh, yl);
                ph[i] = pack_h2(xh, yh);
                pl[i] = pack_h2(xl, yl);
            }
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {   // two dim tiles per ldmatrix.x4.trans
                uint32_t bh[4], bl[4];
                const uint32_t off = (uint32_t)((16 * jp * LD + 16 * np) * 2) + v_ld;
                ldsm_x4_trans(bh, vh_s + off);
                ldsm_x4_trans(bl, vl_s + off);
                mma_f16(o[2 * np], pl, bh[0], bh[1]);
                mma_f16(o[2 * np + 1], pl, bh[2], bh[3]);
                mma_f16(o[2 * np], ph, bl[0], bl[1]);
                mma_f16(o[2 * np + 1], ph, bl[2], bl[3]);
                mma_f16(o[2 * np], ph, bh[0], bh[1]);
                mma_f16(o[2 * np + 1], ph, bh[2], bh[3]);
            }
        }
    }
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (r0 < M) *reinterpret_cast<float2*>(out + (size_t)r0 * H + hoff + 8 * n + 2 * t) = make_float2(o[n][0] * i0, o[n][1] * i0);
        if (r1 < M) *reinterpret_cast<float2*>(out + (size_t)r1 * H + hoff + 8 * n + 2 * t) = make_float2(o[n][2] * i1, o[n][3] * i1);
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
    pdl_wait_prior_grid();       // (no-op unless launched with programmatic stream serialization: the lockstep step's graph)
    pdl_launch_dependents();
    const int m = blockIdx.x;
    const float* e = emb + (size_t)tokens[m] * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) x[(size_t)m * H + i] = e[i];
}

}  // namespace tib
