// kernels.cuh -- the small kernels around the streaming GEMV: quantize / pack / unpack, the reference-order fp32
// matmul, element-wise ops, RMSNorm, RoPE, softmax, split-K flash-decoding attention over the paged KV cache,
// embedding lookup and the greedy-token bookkeeping.  All fp32, no fast-math: divisions, sqrt, roundf, expf are
// the IEEE / full-precision versions so that integer results are bit-exact with the reference.
#pragma once
#include <cfloat>
#include "gemv.cuh"

namespace tib {

// ---------------------------------------------------------------------------------------------------
// synthetic weights: uniform(-amp, amp) from a counter-based hash (splitmix64) of (seed, index)
// ---------------------------------------------------------------------------------------------------
TIB_HD uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
TIB_HD float synth_uniform(uint64_t seed, uint64_t idx, float amp) {
    const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + idx);
    const float u = (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);  // [0,1), 24 bits
    return (2.0f * u - 1.0f) * amp;
}
__global__ void synth_fill_kernel(float* out, size_t n, uint64_t seed, float amp) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = synth_uniform(seed, i, amp);
}

// ---------------------------------------------------------------------------------------------------
// Quantizer::calculate_quantization_info (src/optimize/quantization.cpp:335-394): min / max scan
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order(float f) {  // order-preserving float -> uint
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_unorder(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// mm[0] = ordered min (init 0xFFFFFFFF), mm[1] = ordered max (init 0)
__global__ void minmax_kernel(const float* x, size_t n, uint32_t* mm) {
    float mn = INFINITY, mx = -INFINITY;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], float_order(mn));
        atomicMax(&mm[1], float_order(mx));
    }
}
// sz[0] = scale, sz[1] = zero_point, with the reference's exact expressions (:352-388)
__global__ void quant_params_kernel(const uint32_t* mm, int qtype, int symmetric, float* sz) {
    const float mn = float_unorder(mm[0]), mx = float_unorder(mm[1]);
    float scale, zp;
    if (symmetric) {
        const float amax = fmaxf(fabsf(mn), fabsf(mx));
        scale = __fdiv_rn(amax, qtype == 0 ? 127.0f : 7.0f);
        zp = 0.0f;
    } else {
        scale = __fdiv_rn(__fsub_rn(mx, mn), qtype == 0 ? 255.0f : 15.0f);
        zp = __fdiv_rn(-mn, scale);
    }
    sz[0] = scale;
    sz[1] = zp;
}

// quantize_to_int8 (:662-674) / quantize_to_int4 (:676-693): the integer the reference stores
__device__ __forceinline__ int quantize_one(float x, int qtype, float scale, float zp) {
    if (qtype == 0) {
        float v = roundf(__fadd_rn(__fdiv_rn(x, scale), zp));
        v = fmaxf(-128.0f, fminf(127.0f, v));
        return (int)v;
    }
    float v = roundf(__fsub_rn(__fdiv_rn(x, scale), zp));
    if (zp == 0.0f) v = fmaxf(-7.0f, fminf(7.0f, v));
    else v = fmaxf(0.0f, fminf(15.0f, v));
    return (int)v;
}
__global__ void quantize_flat_kernel(const float* x, size_t n, int qtype, float scale, float zp, int8_t* q8, int32_t* q32) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int q = quantize_one(x[i], qtype, scale, zp);
        if (qtype == 0) q8[i] = (int8_t)q;
        else q32[i] = q;
    }
}
// dequantize_from_int8 (:695-703): scale*(q - zp);  dequantize_from_int4 (:705-713): scale*(q + zp)
__global__ void dequantize_flat_kernel(const int8_t* q8, const int32_t* q32, size_t n, int qtype, float scale, float zp, float* x) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (qtype == 0) x[i] = __fmul_rn(scale, __fsub_rn((float)q8[i], zp));
        else x[i] = __fmul_rn(scale, __fadd_rn((float)q32[i], zp));
    }
}

// ---------------------------------------------------------------------------------------------------
// pack: fp32 [K][N_src] sources -> quantize -> the streaming layout of qlayout.cuh
// A packed matrix may fuse several reference tensors that share x:
//   mode 0  concatenate along N (q | k | v), mode 1  interleave two sources (gate_i, up_i)
// Each source keeps its own per-tensor scale / zero-point (written per column into colscale / colzterm).
// ---------------------------------------------------------------------------------------------------
struct PackSrc {
    const void* w;    // on the device: element (k, c) at w[k * ld + c] -- a whole tensor or a tensor-parallel shard view
    int n;            // columns of the view
    int ld;           // row stride of the underlying tensor (= n when the view is the whole tensor)
    int kind;         // 0: fp32, quantized here; 1 / 2: int8 / int32 integers of an already quantized tensor (.tinq), stored as they are
    int shift4;       // INT4 integers in the asymmetric range [0, 15] under the signed nibble code (off4 = 8): stored as q - 8, and
                      // 8 joins the zero-point term
    const float* sz;  // device: scale, zero_point (of the WHOLE tensor: quantize first, then shard -- SURVEY.md 8e)
};
struct PackArgs {
    PackSrc src[3];
    int nsrc;
    int mode;
    int qtype;       // 0 int8, 1 int4
    int off4;        // INT4: value added before storing the nibble (8 symmetric, 0 asymmetric)
    int unit_scale;  // 1: colscale = 1 (compat_literal: unscaled integers)
    QLayout L;
    uint8_t* out;
    float* colscale;
    float* colzterm;
};
__device__ __forceinline__ void pack_locate(const PackArgs& a, int n, int& si, int& sc) {
    if (a.mode == 1) { si = n & 1; sc = n >> 1; return; }
    si = 0; sc = n;
    while (si < a.nsrc - 1 && sc >= a.src[si].n) { sc -= a.src[si].n; ++si; }
}
// grid = P slabs, block = 512
__global__ void pack_kernel(const PackArgs a) {
    const QLayout& L = a.L;
    const Slab slab = make_slab(L, blockIdx.x);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kik = kitem_k(L.bits), nel = L.bits == 4 ? 8 : 4;
    for_each_quad_word(L, slab, warp, lane, [&](size_t off, int nl, int grp, int chunk, int ki, int widx) {
        uint32_t word = 0;
        for (int i = 0; i < nel; ++i) {
            int row, kk;
            kitem_word_elem(L.bits, nl, widx, i, row, kk);
            const int n = slab.col0 + 16 * grp + row, k = chunk * L.kc + ki * kik + kk;
            int qv = 0;
            const bool live = n < L.N && k < L.K;
            if (live) {
                int si, sc;
                pack_locate(a, n, si, sc);
                const PackSrc& ps = a.src[si];
                const size_t e = (size_t)k * ps.ld + sc;
                if (ps.kind == 0) qv = quantize_one(static_cast<const float*>(ps.w)[e], a.qtype, ps.sz[0], ps.sz[1]);
                else if (ps.kind == 1) qv = static_cast<const int8_t*>(ps.w)[e];
                else qv = static_cast<const int32_t*>(ps.w)[e] - ps.shift4;
            }
            if (L.bits == 4) word |= ((uint32_t)(live ? qv + a.off4 : a.off4) & 0xFu) << (4 * i);   // padding stores q = 0
            else word |= ((uint32_t)qv & 0xFFu) << (8 * i);
        }
        *reinterpret_cast<uint32_t*>(a.out + off) = word;
    });
    // per-column scale and zero-point term:  y = scale * (sum x*(u - off) + zterm * sum x)
    //   INT8 dequant = scale*(q - zp), stored q          -> zterm = -zp
    //   INT4 dequant = scale*(q + zp), stored u = q + off -> zterm = +zp
    for (int c = threadIdx.x; c < slab.ncols; c += blockDim.x) {
        const int n = slab.col0 + c;
        const bool col_live = n < L.N;
        float scale = 1.f, zp = 0.f;
        if (col_live) {
            int si, sc;
            pack_locate(a, n, si, sc);
            scale = a.src[si].sz[0];
            zp = a.src[si].sz[1] + (float)a.src[si].shift4;
        }
        a.colscale[n] = (col_live && !a.unit_scale) ? scale : (col_live ? 1.0f : 0.0f);
        float zt = 0.f;
        if (col_live && zp != 0.0f) zt = a.qtype == 0 ? -zp : zp;
        if (a.colzterm) a.colzterm[n] = a.unit_scale ? 0.f : zt;
    }
}
// inverse of pack for a single-source matrix: q_out[k][n] in the reference's [K,N] int32 order
__global__ void unpack_kernel(const uint8_t* packed, QLayout L, int off4, int32_t* q_out) {
    const Slab slab = make_slab(L, blockIdx.x);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kik = kitem_k(L.bits), nel = L.bits == 4 ? 8 : 4;
    for_each_quad_word(L, slab, warp, lane, [&](size_t off, int nl, int grp, int chunk, int ki, int widx) {
        const uint32_t word = *reinterpret_cast<const uint32_t*>(packed + off);
        for (int i = 0; i < nel; ++i) {
            int row, kk;
            kitem_word_elem(L.bits, nl, widx, i, row, kk);
            const int n = slab.col0 + 16 * grp + row, k = chunk * L.kc + ki * kik + kk;
            if (n >= L.N || k >= L.K) continue;
            q_out[(size_t)k * L.N + n] = L.bits == 4 ? (int)((word >> (4 * i)) & 0xFu) - off4 : (int)(int8_t)((word >> (8 * i)) & 0xFFu);
        }
    });
}

// ---------------------------------------------------------------------------------------------------
// TensorEngine::matmul on fp32 (src/core/tensor_engine.cpp:490-640) in the reference build's order of roundings
// (its order is pinned by the test-side restatement, function tio_matmul): thread = output column, k sequential.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float chain_unfused4(const float* a, const float* b, size_t ldb, int k0, int k1, float s) {
    const int kv = k0 + ((k1 - k0) / 4) * 4;
    for (int k = k0; k < kv; ++k) s = __fadd_rn(s, __fmul_rn(a[k], b[(size_t)k * ldb]));
    for (int k = kv; k < k1; ++k) s = fmaf(a[k], b[(size_t)k * ldb], s);
    return s;
}
__global__ void matmul_f32_exact_kernel(const float* A, const float* B, float* C, int M, int K, int N, int relu) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N) return;
    const float* a = A + (size_t)i * K;
    const float* b = B + j;
    const bool tiled = M >= 32 && N >= 32 && K >= 32;
    float s = 0.f;
    if (tiled && j < (N / 8) * 8) {
        for (int k = 0; k < K; ++k) s = fmaf(a[k], b[(size_t)k * N], s);
    } else if (tiled) {
        for (int k0 = 0; k0 < K; k0 += 256) s = chain_unfused4(a, b, N, k0, k0 + 256 < K ? k0 + 256 : K, s);
    } else {
        s = chain_unfused4(a, b, N, 0, K, s);
    }
    C[(size_t)i * N + j] = relu ? fmaxf(s, 0.f) : s;
}

// ---------------------------------------------------------------------------------------------------
// element-wise ops and row ops (stand-alone TensorEngine entry points; the decode path uses the fused forms)
// ---------------------------------------------------------------------------------------------------
enum EwOp : int { EW_SILU = 0, EW_RELU = 1, EW_ADD = 2, EW_MUL = 3, EW_SILU_MUL = 4 };
__global__ void elementwise_kernel(const float* a, const float* b, float* y, size_t n, int op) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = a[i];
        float r;
        switch (op) {
            case EW_SILU: r = x / (1.0f + expf(-x)); break;              // :918
            case EW_RELU: r = fmaxf(0.0f, x); break;                     // :856
            case EW_ADD: r = x + b[i]; break;                            // :1664
            case EW_MUL: r = x * b[i]; break;                            // :1729
            default: r = b[i] * (x / (1.0f + expf(-x))); break;          // multiply(up, silu(gate)), a = gate, b = up
        }
        y[i] = r;
    }
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    __syncthreads();
    return s;
}
__device__ __forceinline__ float block_max_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = -INFINITY;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s = fmaxf(s, red[i]);
    __syncthreads();
    return s;
}

// rms_norm (:1452-1508), one block per row
__global__ void rms_norm_kernel(const float* x, const float* w, float* y, int H, float eps) {
    __shared__ float red[32];
    const float* xr = x + (size_t)blockIdx.x * H;
    float ss = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) ss = fmaf(xr[i], xr[i], ss);
    const float tot = block_sum_256(ss, red);
    const float rms = sqrtf(tot / (float)H + eps);
    for (int i = threadIdx.x; i < H; i += blockDim.x) y[(size_t)blockIdx.x * H + i] = (xr[i] / rms) * w[i];
}

// softmax, scalar branch (:1017-1033), one block per row
__global__ void softmax_kernel(const float* x, float* y, int n, float temperature) {
    __shared__ float red[32];
    const float* xr = x + (size_t)blockIdx.x * n;
    float* yr = y + (size_t)blockIdx.x * n;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, xr[i]);
    mx = block_max_256(mx, red);
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = expf((xr[i] - mx) / temperature);
        yr[i] = v;
        s += v;
    }
    const float tot = block_sum_256(s, red);
    for (int i = threadIdx.x; i < n; i += blockDim.x) yr[i] = yr[i] / tot;
}

// apply_rope (:1510-1624): rows = B*heads*T vectors of D floats; pos index = (b*T + s) or s
__global__ void rope_kernel(const float* x, const float* pos, const float* inv_freq, float* y, int heads, int T, int D, int pos_2d, size_t rows) {
    const int half = D / 2;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < rows * half; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = idx / half;
        const int i = (int)(idx - row * half);
        const size_t s = row % T, b = row / ((size_t)heads * T);
        const float p = pos_2d ? pos[b * T + s] : pos[s];
        float sn, cs;
        sincosf(p * inv_freq[i], &sn, &cs);
        const float xe = x[row * D + 2 * i], xo = x[row * D + 2 * i + 1];
        y[row * D + 2 * i] = __fsub_rn(__fmul_rn(xe, cs), __fmul_rn(xo, sn));
        y[row * D + 2 * i + 1] = __fadd_rn(__fmul_rn(xe, sn), __fmul_rn(xo, cs));
    }
}

// ---------------------------------------------------------------------------------------------------
// Split-K flash-decoding attention over the paged KV cache.
// Replaces attention_fast_incremental (:1254-1388) / multi_head_attention q_len 1 (:1149-1252) and the
// copy-out of KVCache::update_incremental (inference_engine.cpp:147-157): K/V are read in place, once.
//   grid (heads, splits): CTA (h, j) covers tokens [j*chunk, (j+1)*chunk) of head h:
//     scores s_t = scale * q_h . K_t  (warp per token, float4 lanes, shuffle reduce) -> shared memory
//     block max / exp / sum, o_d = sum_t p_t V_t[d] (threads over d: coalesced rows, no shuffles),
//     online rescale across sub-blocks of 256 tokens; writes (m, l, o) partials
//   attn_combine_kernel merges the splits: o = sum_j e^{m_j - M} o_j / sum_j e^{m_j - M} l_j
// ---------------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 256;
constexpr int kAttnTokBlock = 256;

struct AttnArgs {
    const float* q;          // [heads*D]
    const float* k_pool;     // [page][page_tokens][H]
    const float* v_pool;
    const int* page_table;
    int page_tokens;
    const int* pos_ptr;      // device scalar; tokens in cache = *pos_ptr + t_bias
    int t_bias;              // 1 inside the decode step (the current token was just appended)
    int H, D, heads;
    int max_splits, min_chunk;
    float scale;
    float* part_o;           // [heads][max_splits][D]
    float* part_ml;          // [heads][max_splits][2]
    float* out;              // [heads*D]
    // batched decode (stand-alone kernels only): blockIdx.z = sequence; element strides per sequence of q / out / the page
    // table / the partial buffers (0 for a single sequence)
    int zq, zout, ztable;
    size_t zpart_o, zpart_ml;
    int page_shift;          // log2(page_tokens) when it is a power of two (the fast item then needs no division), else -1
};
TIB_HD int log2_if_pow2(int v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

__device__ __forceinline__ const float* kv_row(const float* pool, const int* table, int page_tokens, int H, int t) {
    const int page = table[t / page_tokens];
    return pool + ((size_t)page * page_tokens + (t % page_tokens)) * H;
}

__device__ __forceinline__ void attn_split_range(int t, int max_splits, int min_chunk, int& nsplit, int& chunk) {
    nsplit = (t + min_chunk - 1) / min_chunk;
    if (nsplit > max_splits) nsplit = max_splits;
    if (nsplit < 1) nsplit = 1;
    chunk = (t + nsplit - 1) / nsplit;
}

// KVCache::update_incremental (inference_engine.cpp:78-160) on the paged pools: new keys / values arrive in the reference's
// [heads, new_tokens, head_dim] order and are appended at token positions pos0 .. pos0 + new_tokens - 1 (one pool row per
// token, heads side by side: the layout the attention kernels read)
__global__ void kv_append_kernel(const float* k_new, const float* v_new, int heads, int new_tokens, int head_dim, int pos0, float* k_pool,
                                 float* v_pool, const int* page_table, int page_tokens) {
    const int H = heads * head_dim;
    const size_t n = (size_t)new_tokens * H;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / H), c = (int)(i - (size_t)t * H), h = c / head_dim, d = c - h * head_dim;
        const size_t src = ((size_t)h * new_tokens + t) * head_dim + d;
        const int pos = pos0 + t;
        const size_t dst = ((size_t)page_table[pos / page_tokens] * page_tokens + (pos % page_tokens)) * H + c;
        k_pool[dst] = k_new[src];
        v_pool[dst] = v_new[src];
    }
}
// the copy-out of update_incremental (:132-157): the first `len` tokens as [heads, len, head_dim]
__global__ void kv_read_kernel(const float* k_pool, const float* v_pool, const int* page_table, int page_tokens, int heads, int len, int head_dim,
                               float* k_out, float* v_out) {
    const int H = heads * head_dim;
    const size_t n = (size_t)len * H;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / H), c = (int)(i - (size_t)t * H), h = c / head_dim, d = c - h * head_dim;
        const size_t src = ((size_t)page_table[t / page_tokens] * page_tokens + (t % page_tokens)) * H + c;
        const size_t dst = ((size_t)h * len + t) * head_dim + d;
        k_out[dst] = k_pool[src];
        v_out[dst] = v_pool[src];
    }
}

// (attn_partial_kernel / attn_combine_kernel are defined in mega.cuh on top of attn_item / attn_merge_head)

// ---------------------------------------------------------------------------------------------------
// decode-step bookkeeping
// ---------------------------------------------------------------------------------------------------
struct StepState {
    int pos;        // tokens in the KV cache
    int token;      // token to feed next
    int step;       // tokens written to out_tokens
    int pad;
    unsigned long long argmax_key;
};

// x = token_embeddings[token]  (the lookup of the dead InferenceEngineImpl::forward_pass, inference_engine.cpp:594-612);
// compat_literal: x[i] = 0.1f * (i % 100)  (:1508-1512)
__global__ void embed_kernel(const float* emb, const StepState* st, float* x, int H, int literal) {
    const int tok = st->token;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H; i += gridDim.x * blockDim.x)
        x[i] = literal ? 0.1f * (float)(i % 100) : emb[(size_t)tok * H + i];
}

// compat_literal (BASELINE.json configs[0], the literal benchmark_inference path):
//   placeholder embeddings over the flattened [1,T,H] index (inference_engine.cpp:1444-1448, :1508-1512)
__global__ void literal_embed_kernel(float* x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = 0.1f * (float)(i % 100);
}
//   Quantizer::quantize_model (:89-118) followed by convert_dtype's cast WITHOUT the scale (tensor_engine.cpp:2218-2253,
//   SURVEY R8): w <- float(q(w)), in place; sz = {scale, zero_point} of the whole tensor
__global__ void literal_quant_kernel(float* w, size_t n, int qtype, const float* sz) {
    const float scale = sz[0], zp = sz[1];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        w[i] = (float)quantize_one(w[i], qtype, scale, zp);
}

// XStats of every embedding row against the norm weight of the first GEMV (gemv.cuh XStats): one block per row
__global__ void emb_stats_kernel(const float* emb, const float* norm_w, XStats* out, int H) {
    __shared__ float red[64];
    const float* e = emb + (size_t)blockIdx.x * H;
    float ss = 0.f, am = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        const float v = e[i];
        ss = fmaf(v, v, ss);
        am = fmaxf(am, fabsf(norm_w ? v * norm_w[i] : v));
    }
    ss = warp_sum(ss);
    am = warp_max(am);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[warp] = ss; red[32 + warp] = am; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f, a = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s += red[i]; a = fmaxf(a, red[32 + i]); }
        out[blockIdx.x] = XStats{s, a};
    }
}

struct StepIO {
    int* out_tokens;  // generated ids, [out_cap]
    int out_cap;
    float* hist;      // optional logits history [hist_cap][V]
    int hist_cap;
};

// after the lm_head (or after the last layer during prefill): optionally keep the logits, decode the argmax key
// (greedy = the top_k 1 branch of sample_next_token, inference_engine.cpp:1585-1598), publish the token and
// advance the cache position
__global__ void step_finish_kernel(StepState* st, const StepIO* io, const float* logits, int V, int sample) {
    const int step = st->step;
    if (sample && io->hist && step < io->hist_cap)
        for (int i = threadIdx.x; i < V; i += blockDim.x) io->hist[(size_t)step * V + i] = logits[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sample) {
            // key 0 = no logit compared greater than -inf (all NaN / -inf): defined as token 0, like the persistent kernel
            int tok = st->argmax_key == 0ull ? 0 : 0x7FFFFFFF - (int)(uint32_t)(st->argmax_key & 0xFFFFFFFFull);
            tok = min(max(tok, 0), V - 1);
            st->token = tok;
            if (io->out_tokens && step < io->out_cap) io->out_tokens[step] = tok;
            st->step = step + 1;
        }
        st->argmax_key = 0ull;
        st->pos += 1;
    }
}
__global__ void set_token_kernel(StepState* st, const int* prompt, int i) { st->token = prompt[i]; }

}  // namespace tib
