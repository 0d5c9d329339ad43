/*
 * ti_oracle.c -- plain-C restatement of TurboInfer's token-generation hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ti_oracle.h): the CUDA product never links or calls this file.
 * Parity pin: every function below is checked bit-for-bit (or to the stated ulp bound) against the
 * compiled reference (oracle/_ref/libti_ref.so) by tests/test_oracle_vs_ref.py in the build
 * container, and against the golden vectors in tests/golden/ (generated from the compiled reference
 * by tests/golden/make_golden.py) everywhere else.
 *
 * All file:line citations are relative to /root/reference.  The reference's Release build (g++ 13.3,
 * -O3 -mavx2 -mfma) decides where `s += a*b` becomes one FMA and where it stays two roundings; those
 * decisions were read off the disassembly of oracle/_ref and are written out explicitly here (fmaf()
 * = fused, `p = a*b; s = s + p` = un-fused), and this file is compiled with -ffp-contract=off so the
 * rounding does not depend on compiler flags.
 */
#define _POSIX_C_SOURCE 200809L
#include "ti_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

int tio_kind(void) { return 0; }

/* ------------------------------------------------------------------------------------------
 * Quantizer -- src/optimize/quantization.cpp
 * ---------------------------------------------------------------------------------------- */

/* calculate_quantization_info, :335-394.  Per-tensor min/max scan; symmetric: scale = absmax/127
 * (INT8) or absmax/7 (INT4), zero_point = 0; asymmetric: scale = (max-min)/255 or /15,
 * zero_point = -min/scale. */
int tio_quant_info(const float* x, size_t n, int qtype, int symmetric, float* scale, float* zero_point) {
    if (n == 0 || (qtype != TIO_QINT8 && qtype != TIO_QINT4)) return -1;
    float mn = x[0], mx = x[0];
    for (size_t i = 1; i < n; ++i) {
        mn = x[i] < mn ? x[i] : mn; /* std::min(min_val, data[i]) :346 */
        mx = mx < x[i] ? x[i] : mx; /* std::max(max_val, data[i]) :347 */
    }
    const float levels_sym = qtype == TIO_QINT8 ? 127.0f : 7.0f;   /* :359, :378 */
    const float levels_asym = qtype == TIO_QINT8 ? 255.0f : 15.0f; /* :363, :381 */
    if (symmetric) {
        float a = fabsf(mn), b = fabsf(mx);
        float absmax = a < b ? b : a;
        *scale = absmax / levels_sym;
        *zero_point = 0.0f;
    } else {
        *scale = (mx - mn) / levels_asym;
        *zero_point = -mn / *scale;
    }
    return 0;
}

/* quantize_to_int8, :662-674: q = clamp(round(x/scale + zp), -128, 127); std::round is
 * half-away-from-zero (roundf), the division is a true fp32 division. */
void tio_quantize_int8(const float* x, int8_t* q, size_t n, float scale, float zp) {
    for (size_t i = 0; i < n; ++i) {
        float v = roundf(x[i] / scale + zp);
        v = fmaxf(-128.0f, fminf(127.0f, v));
        q[i] = (int8_t)v;
    }
}

/* quantize_to_int4, :676-693: q = round(x/scale - zp), clamped to [-7,7] when zp == 0 (symmetric)
 * else [0,15]; stored one value per int32. */
void tio_quantize_int4(const float* x, int32_t* q, size_t n, float scale, float zp) {
    for (size_t i = 0; i < n; ++i) {
        float v = roundf(x[i] / scale - zp);
        if (zp == 0.0f) v = fmaxf(-7.0f, fminf(7.0f, v));
        else v = fmaxf(0.0f, fminf(15.0f, v));
        q[i] = (int32_t)v;
    }
}

/* dequantize_from_int8, :695-703: x = scale * (q - zp) */
void tio_dequantize_int8(const int8_t* q, float* x, size_t n, float scale, float zp) {
    for (size_t i = 0; i < n; ++i) x[i] = scale * ((float)q[i] - zp);
}

/* dequantize_from_int4, :705-713: x = scale * (q + zp) */
void tio_dequantize_int4(const int32_t* q, float* x, size_t n, float scale, float zp) {
    for (size_t i = 0; i < n; ++i) x[i] = scale * ((float)q[i] + zp);
}

/* ------------------------------------------------------------------------------------------
 * TensorEngine -- src/core/tensor_engine.cpp
 * ---------------------------------------------------------------------------------------- */

/* matmul_2d, :538-592 (and matmul_3d_2d :594-640, same inner loops).
 *
 * What the reference COMPUTES per output element is pinned by its Release build (g++ 13.3, -O3 -mavx2
 * -mfma), observed in the disassembly of oracle/_ref and verified bit-for-bit by tests/test_oracle_vs_ref.py:
 *
 *  - scalar i-j-k loop (:565-573, :625-633; taken when M < 32 or N < 32 or K < 32, i.e. every decode GEMV):
 *    GCC vectorises the k loop 4-wide as an IN-ORDER reduction: the four products are rounded on their
 *    own (vmulps) and then added one after the other (4 x vaddss); the K % 4 leftover iterations are
 *    contracted to FMAs (vfmadd231ss).
 *  - AVX2 tiled path simd_gemm_float (:191-255; M, N, K >= 32): full 8-column blocks are a sequential-k
 *    FMA chain (_mm256_fmadd_ps, partial sums reloaded between K tiles of 256, which keeps the order);
 *    the N % 8 remainder columns run the scalar loop per K tile, compiled as above.
 *
 * The loops are run k-outer / j-inner here so the oracle finishes in seconds; the per-element sequence
 * of roundings is unchanged.  (This file is compiled with -ffp-contract=off: `c + a*b` is two roundings,
 * fmaf() is one.) */
static void matmul_cols_unfused4(const float* ai, const float* b, float* ci, size_t k0, size_t k1, size_t N,
                                 size_t j0, size_t j1) {
    size_t kv = k0 + ((k1 - k0) / 4) * 4;
    for (size_t k = k0; k < kv; ++k) {
        const float av = ai[k];
        const float* bk = b + k * N;
        for (size_t j = j0; j < j1; ++j) {
            float p = av * bk[j];
            ci[j] = ci[j] + p;
        }
    }
    for (size_t k = kv; k < k1; ++k) {
        const float av = ai[k];
        const float* bk = b + k * N;
        for (size_t j = j0; j < j1; ++j) ci[j] = fmaf(av, bk[j], ci[j]);
    }
}

void tio_matmul(const float* a, const float* b, float* c, size_t M, size_t K, size_t N) {
    const int tiled = M >= 32 && N >= 32 && K >= 32; /* :561 */
    const size_t n8 = tiled ? (N / 8) * 8 : 0;
    for (size_t i = 0; i < M; ++i) {
        float* ci = c + i * N;
        const float* ai = a + i * K;
#pragma omp parallel for schedule(static) if (N * K > (1u << 20))
        for (size_t jb = 0; jb < N; jb += 256) {
            size_t je = jb + 256 < N ? jb + 256 : N;
            for (size_t j = jb; j < je; ++j) ci[j] = 0.0f;
            /* columns [jb, jf) belong to full 8-wide AVX2 blocks, [jf, je) to the scalar loop */
            size_t jf = je < n8 ? je : (jb < n8 ? n8 : jb);
            if (jf > jb)
                for (size_t k = 0; k < K; ++k) {
                    const float av = ai[k];
                    const float* bk = b + k * N;
                    for (size_t j = jb; j < jf; ++j) ci[j] = fmaf(av, bk[j], ci[j]);
                }
            if (je > jf) {
                if (tiled) for (size_t k0 = 0; k0 < K; k0 += 256)
                    matmul_cols_unfused4(ai, b, ci, k0, k0 + 256 < K ? k0 + 256 : K, N, jf, je);
                else matmul_cols_unfused4(ai, b, ci, 0, K, N, jf, je);
            }
        }
    }
}

/* rms_norm, :1452-1508: y = (x / sqrt(sum(x^2)/H + eps)) * w, sequential fp32 sum of squares */
void tio_rms_norm(const float* x, const float* w, float* y, size_t rows, size_t H, float eps) {
    for (size_t r = 0; r < rows; ++r) {
        const float* xr = x + r * H;
        /* :1497-1500.  The build vectorises this in-order: squares rounded on their own and added one
         * by one for the 8-wide body and one 4-wide block, FMAs for the last H % 4 elements. */
        float ss = 0.0f;
        size_t i = 0, hv = (H / 4) * 4;
        for (; i < hv; ++i) {
            float p = xr[i] * xr[i];
            ss = ss + p;
        }
        for (; i < H; ++i) ss = fmaf(xr[i], xr[i], ss);
        float rms = sqrtf(ss / (float)H + eps);                      /* :1501 */
        for (i = 0; i < H; ++i) y[r * H + i] = (xr[i] / rms) * w[i]; /* :1504-1506 */
    }
}

/* apply_rope, :1510-1624: interleaved pairs (2i, 2i+1), freq_i = 1/pow(theta, 2i/D),
 * angle = pos * freq_i, (x,y) -> (x cos - y sin, x sin + y cos).  3-D input rotates over the whole
 * last dim, 4-D [B,nh,T,D] over head_dim; positions are fp32. */
void tio_rope(const float* x, const float* pos, float* y, size_t B, size_t nh, size_t T, size_t D,
              int ndim, int pos_2d, float theta) {
    size_t half = D / 2;
    float* freqs = (float*)malloc(sizeof(float) * (half ? half : 1));
    for (size_t i = 0; i < half; ++i) freqs[i] = 1.0f / powf(theta, (float)(2 * i) / (float)D); /* :1562-1565 */
    size_t heads = ndim == 4 ? nh : 1;
    for (size_t b = 0; b < B; ++b)
        for (size_t h = 0; h < heads; ++h)
            for (size_t s = 0; s < T; ++s) {
                float p = pos_2d ? pos[b * T + s] : pos[s];
                size_t base = ((b * heads + h) * T + s) * D;
                for (size_t i = 0; i < half; ++i) {
                    float xe = x[base + 2 * i], xo = x[base + 2 * i + 1];
                    float c = cosf(p * freqs[i]), sn = sinf(p * freqs[i]);
                    /* the build keeps the four products un-fused (vmulps + vaddsubps), :1584-1585, :1611-1612 */
                    float a0 = xe * c, a1 = xo * sn, b0 = xe * sn, b1 = xo * c;
                    y[base + 2 * i] = a0 - a1;
                    y[base + 2 * i + 1] = b0 + b1;
                }
            }
    free(freqs);
}

/* silu, :900-923: x / (1 + exp(-x)) */
void tio_silu(const float* x, float* y, size_t n) {
    for (size_t i = 0; i < n; ++i) y[i] = x[i] / (1.0f + expf(-x[i]));
}
/* relu, :828-869 */
void tio_relu(const float* x, float* y, size_t n) {
    for (size_t i = 0; i < n; ++i) y[i] = x[i] > 0.0f ? x[i] : 0.0f;
}
/* add, :1626-1678; multiply, :1680-1743 */
void tio_add(const float* a, const float* b, float* y, size_t n) {
    for (size_t i = 0; i < n; ++i) y[i] = a[i] + b[i];
}
void tio_mul(const float* a, const float* b, float* y, size_t n) {
    for (size_t i = 0; i < n; ++i) y[i] = a[i] * b[i];
}

/* softmax, scalar branch :1017-1033: exp((x - max)/T) / sum, sequential sum.  (The AVX2 branch
 * :949-1013 is ~12 % wrong, SURVEY R10, and is deliberately not restated.) */
void tio_softmax(const float* x, float* y, size_t rows, size_t n, float temperature) {
    for (size_t r = 0; r < rows; ++r) {
        const float* xr = x + r * n;
        float* yr = y + r * n;
        float mx = xr[0];
        for (size_t i = 1; i < n; ++i) mx = mx < xr[i] ? xr[i] : mx;
        float sum = 0.0f;
        for (size_t i = 0; i < n; ++i) {
            float v = expf((xr[i] - mx) / temperature);
            yr[i] = v;
            sum += v;
        }
        for (size_t i = 0; i < n; ++i) yr[i] /= sum;
    }
}

/* attention_fast_incremental, :1254-1388, SIMD build.  Per batch row:
 *   score[t] = scale * hsum8( fma-accumulated 8-lane partial dots over h ) (+ scalar tail)   :1295-1327
 *   p = exp(score - max) / sum                                                               :1330-1340
 *   out[h]  = hsum8( fma-accumulated 8-lane partials over t blocks of 8 ) (+ scalar tail)    :1343-1384
 * hsum8 adds lanes 0..7 in order into a float that starts at 0. */
/* remainder loops (:1320-1323, :1377-1379), n < 8: the build runs one 4-wide block with un-fused
 * products added in order when n >= 4, then FMAs for what is left */
static float tail_dot(const float* a, size_t sa, const float* b, size_t sb, size_t n, float s) {
    size_t i = 0;
    if (n >= 4) {
        for (; i < 4; ++i) {
            float p = a[i * sa] * b[i * sb];
            s = s + p;
        }
    }
    for (; i < n; ++i) s = fmaf(a[i * sa], b[i * sb], s);
    return s;
}

void tio_attention_fast_incremental(const float* q, const float* k, const float* v, float* out,
                                    size_t B, size_t t, size_t H) {
    const float scale = 1.0f / sqrtf((float)H); /* :1288 */
    float* sc = (float*)malloc(sizeof(float) * (t ? t : 1));
    for (size_t b = 0; b < B; ++b) {
        const float* qb = q + b * H;
        const float* kb = k + b * t * H;
        const float* vb = v + b * t * H;
        size_t h8 = (H / 8) * 8;
        for (size_t p = 0; p < t; ++p) {
            float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (size_t h = 0; h < h8; h += 8)
                for (int j = 0; j < 8; ++j) lane[j] = fmaf(qb[h + j], kb[p * H + h + j], lane[j]);
            float s = 0.0f;
            for (int j = 0; j < 8; ++j) s += lane[j];
            s = tail_dot(qb + h8, 1, kb + p * H + h8, 1, H - h8, s);
            sc[p] = s * scale;
        }
        float mx = sc[0];
        for (size_t p = 1; p < t; ++p) mx = mx < sc[p] ? sc[p] : mx;
        float sum = 0.0f;
        for (size_t p = 0; p < t; ++p) {
            sc[p] = expf(sc[p] - mx);
            sum += sc[p];
        }
        for (size_t p = 0; p < t; ++p) sc[p] /= sum;
        size_t t8 = (t / 8) * 8;
        for (size_t h = 0; h < H; ++h) {
            float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (size_t p = 0; p < t8; p += 8)
                for (int j = 0; j < 8; ++j) lane[j] = fmaf(sc[p + j], vb[(p + j) * H + h], lane[j]);
            float o = 0.0f;
            for (int j = 0; j < 8; ++j) o += lane[j];
            o = tail_dot(sc + t8, 1, vb + t8 * H + h, H, t - t8, o);
            out[b * H + h] = o;
        }
    }
    free(sc);
}

/* multi_head_attention with q_len == 1, :1149-1252: slice head h out of q/k/v, run `attention`
 * (which dispatches q_len == 1 to attention_fast_incremental, :1077-1081, with hidden = head_dim),
 * concatenate. */
void tio_multi_head_attention(const float* q, const float* k, const float* v, float* out,
                              size_t B, size_t t, size_t H, size_t nh) {
    size_t hd = H / nh;
    float* qh = (float*)malloc(sizeof(float) * B * hd);
    float* kh = (float*)malloc(sizeof(float) * B * t * hd);
    float* vh = (float*)malloc(sizeof(float) * B * t * hd);
    float* oh = (float*)malloc(sizeof(float) * B * hd);
    for (size_t h = 0; h < nh; ++h) {
        for (size_t b = 0; b < B; ++b) {
            memcpy(qh + b * hd, q + b * H + h * hd, hd * sizeof(float));
            for (size_t p = 0; p < t; ++p) {
                memcpy(kh + (b * t + p) * hd, k + (b * t + p) * H + h * hd, hd * sizeof(float));
                memcpy(vh + (b * t + p) * hd, v + (b * t + p) * H + h * hd, hd * sizeof(float));
            }
        }
        tio_attention_fast_incremental(qh, kh, vh, oh, B, t, hd);
        for (size_t b = 0; b < B; ++b) memcpy(out + b * H + h * hd, oh + b * hd, hd * sizeof(float));
    }
    free(qh); free(kh); free(vh); free(oh);
}

/* ------------------------------------------------------------------------------------------
 * Level B: decode loop, intended dataflow.
 * TransformerLayer::forward_incremental :244-279, compute_attention :291-368, compute_ffn :376-401,
 * InferenceEngine::forward_pass_incremental :1493-1552, generate :734-802, with 2-D [1,H]
 * activations and the real embedding lookup of the dead InferenceEngineImpl::forward_pass :594-612.
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    float *x, *n, *q, *k, *v, *a, *o, *pa, *f, *up, *gate, *act, *ffn, *tmp;
    float **kc, **vc; /* per layer flat [t, H] */
    size_t cap;
} tio_scratch;

static void step_B(const tio_model* m, tio_scratch* s, size_t t, int token, float* logits) {
    const size_t H = m->hidden, V = m->vocab, I = m->inter, L = m->layers, nh = m->heads;
    const size_t hd = H / nh;
    const float posf = (float)t;
    memcpy(s->x, m->tok_emb + (size_t)token * H, H * sizeof(float));
    for (size_t l = 0; l < L; ++l) {
        const float* an = m->attn_norm ? m->attn_norm[l] : NULL;
        const float* wq = m->wq ? m->wq[l] : NULL;
        const float* wk = m->wk ? m->wk[l] : NULL;
        const float* wv = m->wv ? m->wv[l] : NULL;
        const float* wo = m->wo ? m->wo[l] : NULL;
        const float* fn = m->ffn_norm ? m->ffn_norm[l] : NULL;
        const float* wu = m->w_up ? m->w_up[l] : NULL;
        const float* wg = m->w_gate ? m->w_gate[l] : NULL;
        const float* wd = m->w_down ? m->w_down[l] : NULL;

        if (an) tio_rms_norm(s->x, an, s->n, 1, H, m->rms_eps); /* :257 */
        else memcpy(s->n, s->x, H * sizeof(float));
        const float* attn_out = s->n; /* fall-back returns its input, :293-296 */
        if (wq && wk && wv && wo) {
            tio_matmul(s->n, wq, s->q, 1, H, H); /* :299-301 */
            tio_matmul(s->n, wk, s->k, 1, H, H);
            tio_matmul(s->n, wv, s->v, 1, H, H);
            if (m->rope_mode == 1) {
                tio_rope(s->q, &posf, s->tmp, 1, nh, 1, hd, 4, 0, m->rope_theta);
                memcpy(s->q, s->tmp, H * sizeof(float));
                tio_rope(s->k, &posf, s->tmp, 1, nh, 1, hd, 4, 0, m->rope_theta);
                memcpy(s->k, s->tmp, H * sizeof(float));
            } else if (m->rope_mode == 2) {
                tio_rope(s->q, &posf, s->tmp, 1, 1, 1, H, 3, 0, m->rope_theta);
                memcpy(s->q, s->tmp, H * sizeof(float));
                tio_rope(s->k, &posf, s->tmp, 1, 1, 1, H, 3, 0, m->rope_theta);
                memcpy(s->k, s->tmp, H * sizeof(float));
            }
            memcpy(s->kc[l] + t * H, s->k, H * sizeof(float)); /* KVCache append, :78-160 */
            memcpy(s->vc[l] + t * H, s->v, H * sizeof(float));
            if (m->attn_mode == 1) tio_multi_head_attention(s->q, s->kc[l], s->vc[l], s->a, 1, t + 1, H, nh);
            else tio_attention_fast_incremental(s->q, s->kc[l], s->vc[l], s->a, 1, t + 1, H); /* :323-343 */
            tio_matmul(s->a, wo, s->o, 1, H, H); /* :341 */
            attn_out = s->o;
        }
        tio_add(s->x, attn_out, s->pa, H); /* :264 */
        if (fn) tio_rms_norm(s->pa, fn, s->f, 1, H, m->rms_eps); /* :269 */
        else memcpy(s->f, s->pa, H * sizeof(float));
        const float* ffn_out = s->f; /* fall-back, :377-380 */
        if (wu && wd) {
            tio_matmul(s->f, wu, s->up, 1, H, I); /* :383 */
            if (wg) {
                tio_matmul(s->f, wg, s->gate, 1, H, I); /* :389 */
                tio_silu(s->gate, s->gate, I);           /* :390 */
                tio_mul(s->up, s->gate, s->act, I);      /* :391 */
            } else {
                tio_relu(s->up, s->act, I); /* :394 */
            }
            tio_matmul(s->act, wd, s->ffn, 1, I, H); /* :398 */
            ffn_out = s->ffn;
        }
        tio_add(s->pa, ffn_out, s->x, H); /* :276 */
    }
    if (!logits) return;
    if (m->out_norm) tio_rms_norm(s->x, m->out_norm, s->n, 1, H, m->rms_eps); /* :1533-1535 */
    else memcpy(s->n, s->x, H * sizeof(float));
    tio_matmul(s->n, m->lm_head, logits, 1, H, V); /* :1541-1542 */
}

static int argmax_first(const float* x, size_t n) {
    size_t best = 0;
    for (size_t i = 1; i < n; ++i)
        if (x[i] > x[best]) best = i;
    return (int)best;
}

/* bench.py's CPU arm: per-pass wall-clock times of the next tio_decode_greedy call (set by tio_decode_greedy_timed) */
static double* g_step_times = NULL;
static int g_step_cap = 0, g_step_n = 0;
static void step_B_timed(const tio_model* m, tio_scratch* s, size_t t, int token, float* logits) {
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    step_B(m, s, t, token, logits);
    clock_gettime(CLOCK_MONOTONIC, &b);
    if (g_step_times && g_step_n < g_step_cap) g_step_times[g_step_n++] = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

int tio_decode_greedy(const tio_model* m, const int32_t* prompt, int n_prompt, int n_new,
                      int stop_on_eos, int32_t* out_tokens, float* logits_out) {
    if (!m || !m->lm_head || !m->tok_emb || n_prompt <= 0 || n_new < 0) return -1;
    const size_t H = m->hidden, V = m->vocab, I = m->inter, L = m->layers;
    tio_scratch s;
    memset(&s, 0, sizeof(s));
    s.cap = (size_t)n_prompt + (size_t)n_new;
    float** bufs[] = {&s.x, &s.n, &s.q, &s.k, &s.v, &s.a, &s.o, &s.pa, &s.f, &s.ffn, &s.tmp};
    for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); ++i) *bufs[i] = (float*)calloc(H, sizeof(float));
    s.up = (float*)calloc(I ? I : 1, sizeof(float));
    s.gate = (float*)calloc(I ? I : 1, sizeof(float));
    s.act = (float*)calloc(I ? I : 1, sizeof(float));
    s.kc = (float**)calloc(L ? L : 1, sizeof(float*));
    s.vc = (float**)calloc(L ? L : 1, sizeof(float*));
    for (size_t l = 0; l < L; ++l) {
        s.kc[l] = (float*)calloc(s.cap * H, sizeof(float));
        s.vc[l] = (float*)calloc(s.cap * H, sizeof(float));
    }
    float* logits = (float*)calloc(V, sizeof(float));
    size_t t = 0;
    for (int i = 0; i < n_prompt; ++i, ++t) step_B_timed(m, &s, t, prompt[i], i == n_prompt - 1 ? logits : NULL);
    int produced = 0;
    for (int i = 0; i < n_new; ++i) {
        int best = argmax_first(logits, V); /* top_k = 1: :1585-1598 */
        if (logits_out) memcpy(logits_out + (size_t)i * V, logits, V * sizeof(float));
        out_tokens[produced++] = best;
        if (stop_on_eos && best == 2) break; /* :760 */
        if (i + 1 < n_new) { step_B_timed(m, &s, t, best, logits); ++t; }
    }
    for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); ++i) free(*bufs[i]);
    free(s.up); free(s.gate); free(s.act);
    for (size_t l = 0; l < L; ++l) { free(s.kc[l]); free(s.vc[l]); }
    free(s.kc); free(s.vc); free(logits);
    return produced;
}

int tio_decode_greedy_timed(const tio_model* m, const int32_t* prompt, int n_prompt, int n_new,
                            int32_t* out_tokens, double* step_seconds, int cap) {
    g_step_times = step_seconds;
    g_step_cap = cap;
    g_step_n = 0;
    const int rc = tio_decode_greedy(m, prompt, n_prompt, n_new, 0, out_tokens, NULL);
    g_step_times = NULL;
    return rc < 0 ? rc : g_step_n;
}

/* ------------------------------------------------------------------------------------------
 * Sampling: sample_next_token (src/model/inference_engine.cpp:1554-1673) with the uniform passed in.
 * ---------------------------------------------------------------------------------------- */
static uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* the engine's counter-based uniform (turboinfer_b200/csrc/sampling.cuh sample_uniform): 24 bits in [0, 1) */
float tio_uniform(uint64_t seed, uint64_t step) {
    const uint64_t h = splitmix64(splitmix64(seed) ^ (step * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull));
    return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}

typedef struct { float v; int idx; } spair;
static int spair_desc(const void* a, const void* b) {   /* value descending, index ascending (see ti_oracle.h) */
    const spair* x = (const spair*)a;
    const spair* y = (const spair*)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

int tio_sample(const float* logits, size_t vocab, float temperature, int top_k, float top_p, float u, float* logprob) {
    const size_t V = vocab;
    float* l = (float*)malloc(V * sizeof(float));
    float* p = (float*)malloc(V * sizeof(float));
    spair* pairs = (spair*)malloc(V * sizeof(spair));
    memcpy(l, logits, V * sizeof(float));
    if (temperature != 1.0f && temperature > 0.0f)                       /* :1577-1582 */
        for (size_t i = 0; i < V; ++i) l[i] = l[i] / temperature;
    if (top_k > 0 && (size_t)top_k < V) {                                /* :1584-1598 */
        for (size_t i = 0; i < V; ++i) { pairs[i].v = l[i]; pairs[i].idx = (int)i; }
        qsort(pairs, V, sizeof(spair), spair_desc);
        for (size_t i = (size_t)top_k; i < V; ++i) l[pairs[i].idx] = -INFINITY;
    }
    float mx = l[0];                                                     /* :1600-1612 */
    for (size_t i = 1; i < V; ++i) if (l[i] > mx) mx = l[i];
    float sum = 0.0f;
    for (size_t i = 0; i < V; ++i) { p[i] = expf(l[i] - mx); sum += p[i]; }
    for (size_t i = 0; i < V; ++i) p[i] = p[i] / sum;
    if (top_p < 1.0f) {                                                  /* :1614-1648 */
        for (size_t i = 0; i < V; ++i) { pairs[i].v = p[i]; pairs[i].idx = (int)i; }
        qsort(pairs, V, sizeof(spair), spair_desc);
        float cum = 0.0f;
        size_t cutoff = V;
        for (size_t i = 0; i < V; ++i) {
            cum += pairs[i].v;
            if (cum >= top_p) { cutoff = i + 1; break; }
        }
        for (size_t i = cutoff; i < V; ++i) p[pairs[i].idx] = 0.0f;
        float ns = 0.0f;
        for (size_t i = 0; i < V; ++i) ns += p[i];
        if (ns > 0.0f) for (size_t i = 0; i < V; ++i) p[i] = p[i] / ns;
    }
    float cum = 0.0f;                                                    /* :1650-1672 */
    int tok = (int)V - 1;
    int found = 0;
    for (size_t i = 0; i < V; ++i) {
        cum += p[i];
        if (u <= cum) { tok = (int)i; found = 1; break; }
    }
    (void)found;
    if (logprob) *logprob = logf(p[tok]);
    free(l); free(p); free(pairs);
    return tok;
}

/* ---- beam search: the expansion of one candidate (:1964-2005) ---- */
int tio_beam_expand(const float* logits, size_t vocab, float temperature, int top_k, float top_p, int beam_size, float* probs, int32_t* tokens) {
    const size_t V = vocab;
    float* p = (float*)malloc(V * sizeof(float));
    spair* pairs = (spair*)malloc(V * sizeof(spair));
    for (size_t i = 0; i < V; ++i) p[i] = temperature != 1.0f ? logits[i] / temperature : logits[i];   /* :1972-1976 */
    float mx = p[0];                                                                                  /* softmax, :1798-1819 */
    for (size_t i = 1; i < V; ++i) if (p[i] > mx) mx = p[i];
    float sum = 0.0f;
    for (size_t i = 0; i < V; ++i) { p[i] = expf(p[i] - mx); sum += p[i]; }
    if (sum > 0.0f) for (size_t i = 0; i < V; ++i) p[i] = p[i] / sum;
    if (top_k > 0 && (size_t)top_k < V) {                                                             /* :1982-1984, :1821-1856 */
        for (size_t i = 0; i < V; ++i) { pairs[i].v = p[i]; pairs[i].idx = (int)i; }
        qsort(pairs, V, sizeof(spair), spair_desc);
        for (size_t i = (size_t)top_k; i < V; ++i) p[pairs[i].idx] = 0.0f;
        float s = 0.0f;
        for (size_t i = 0; i < V; ++i) s += p[i];
        if (s > 0.0f) for (size_t i = 0; i < V; ++i) p[i] = p[i] / s;
    }
    if (top_p < 1.0f) {                                                                               /* :1987-1989, :1858-1909 */
        for (size_t i = 0; i < V; ++i) { pairs[i].v = p[i]; pairs[i].idx = (int)i; }
        qsort(pairs, V, sizeof(spair), spair_desc);
        float cum = 0.0f;
        size_t cutoff = V;
        for (size_t i = 0; i < V; ++i) {
            cum += pairs[i].v;
            if (cum >= top_p) { cutoff = i + 1; break; }
        }
        for (size_t i = cutoff; i < V; ++i) p[pairs[i].idx] = 0.0f;
        float s = 0.0f;
        for (size_t i = 0; i < V; ++i) s += p[i];
        if (s > 0.0f) for (size_t i = 0; i < V; ++i) p[i] = p[i] / s;
    }
    size_t n = 0;                                                                                     /* :1990-2005 */
    for (size_t i = 0; i < V; ++i) if (p[i] > 0.0f) { pairs[n].v = p[i]; pairs[n].idx = (int)i; ++n; }
    qsort(pairs, n, sizeof(spair), spair_desc);
    int cnt = 0;
    for (size_t i = 0; i < n && cnt < beam_size; ++i) { probs[cnt] = pairs[i].v; tokens[cnt] = pairs[i].idx; ++cnt; }
    free(p); free(pairs);
    return cnt;
}

typedef struct { int32_t* toks; int n; float log_prob, score; int finished; long order; } bcand;
static int bcand_by_score(const void* a, const void* b) {      /* normalised score descending; stable through `order` */
    const bcand* x = (const bcand*)a;
    const bcand* y = (const bcand*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);
}
static int bcand_by_logprob(const void* a, const void* b) {    /* the heap of :1921-1925: most probable first */
    const bcand* x = (const bcand*)a;
    const bcand* y = (const bcand*)b;
    if (x->log_prob > y->log_prob) return -1;
    if (x->log_prob < y->log_prob) return 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);
}
static void scratch_init(const tio_model* m, tio_scratch* s, size_t cap) {
    const size_t H = m->hidden, I = m->inter, L = m->layers;
    memset(s, 0, sizeof(*s));
    s->cap = cap;
    float** bufs[] = {&s->x, &s->n, &s->q, &s->k, &s->v, &s->a, &s->o, &s->pa, &s->f, &s->ffn, &s->tmp};
    for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); ++i) *bufs[i] = (float*)calloc(H, sizeof(float));
    s->up = (float*)calloc(I ? I : 1, sizeof(float));
    s->gate = (float*)calloc(I ? I : 1, sizeof(float));
    s->act = (float*)calloc(I ? I : 1, sizeof(float));
    s->kc = (float**)calloc(L ? L : 1, sizeof(float*));
    s->vc = (float**)calloc(L ? L : 1, sizeof(float*));
    for (size_t l = 0; l < L; ++l) {
        s->kc[l] = (float*)calloc(cap * H, sizeof(float));
        s->vc[l] = (float*)calloc(cap * H, sizeof(float));
    }
}
static void scratch_free(const tio_model* m, tio_scratch* s) {
    float** bufs[] = {&s->x, &s->n, &s->q, &s->k, &s->v, &s->a, &s->o, &s->pa, &s->f, &s->ffn, &s->tmp};
    for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); ++i) free(*bufs[i]);
    free(s->up); free(s->gate); free(s->act);
    for (size_t l = 0; l < (size_t)m->layers; ++l) { free(s->kc[l]); free(s->vc[l]); }
    free(s->kc); free(s->vc);
}

/* logits of the last position of a whole sequence (what forward_pass(candidate.tokens) returns, :1961) */
/* returns how many logits it wrote: the vocabulary -- or, on the literal path, what the reference TAKES for the vocabulary (below) */
typedef size_t (*seq_logits_fn)(void* ctx, const int32_t* prompt, int n_prompt, const int32_t* toks, int n, float* logits);

static int beam_search_core(seq_logits_fn fwd, void* ctx, size_t max_logits, const int32_t* prompt, int n_prompt, int max_new, int beam_size,
                            float temperature, int top_k, float top_p, float length_penalty, int eos_token, int32_t* out_tokens,
                            int32_t* out_lens, float* out_logprob, float* out_score, int32_t* out_finished) {
    const int cap_tok = max_new > 0 ? max_new : 1;
    float* logits = (float*)calloc(max_logits, sizeof(float));
    float* eprob = (float*)malloc((size_t)beam_size * sizeof(float));
    int32_t* etok = (int32_t*)malloc((size_t)beam_size * sizeof(int32_t));
    const size_t maxc = (size_t)beam_size * (size_t)beam_size + 1;
    bcand* beam = (bcand*)calloc(maxc, sizeof(bcand));
    bcand* next = (bcand*)calloc(maxc, sizeof(bcand));
    bcand* done = (bcand*)calloc(maxc * (size_t)(cap_tok + 1), sizeof(bcand));
    size_t nbeam = 0, ndone = 0;
    long order = 0;
    beam[0].toks = (int32_t*)calloc((size_t)cap_tok, sizeof(int32_t));
    beam[0].order = order++;
    nbeam = 1;
    for (int step = 0; step < max_new && nbeam > 0; ++step) {
        qsort(beam, nbeam, sizeof(bcand), bcand_by_logprob);
        size_t nnext = 0;
        for (size_t c = 0; c < nbeam; ++c) {
            const size_t V = fwd(ctx, prompt, n_prompt, beam[c].toks, beam[c].n, logits);
            const int cnt = tio_beam_expand(logits, V, temperature, top_k, top_p, beam_size, eprob, etok);
            for (int i = 0; i < cnt; ++i) {
                bcand* nc = &next[nnext++];
                *nc = beam[c];
                nc->toks = (int32_t*)calloc((size_t)cap_tok, sizeof(int32_t));
                memcpy(nc->toks, beam[c].toks, (size_t)beam[c].n * sizeof(int32_t));
                nc->toks[nc->n++] = etok[i];
                nc->log_prob = beam[c].log_prob + logf(eprob[i]);
                nc->finished = etok[i] == eos_token || nc->n >= max_new;                                      /* :2015-2016 */
                nc->score = nc->log_prob / powf((float)(n_prompt + nc->n), length_penalty);                  /* :2024-2026 */
                nc->order = order++;
            }
        }
        for (size_t c = 0; c < nbeam; ++c) free(beam[c].toks);
        qsort(next, nnext, sizeof(bcand), bcand_by_score);                                                   /* :2030-2033 */
        nbeam = 0;
        for (size_t i = 0; i < nnext; ++i) {
            if (i >= (size_t)beam_size) { free(next[i].toks); continue; }
            next[i].order = order++;   /* ties from here on: the order of insertion (what a stable sort keeps) */
            if (next[i].finished) done[ndone++] = next[i];
            else beam[nbeam++] = next[i];
        }
        if (ndone >= (size_t)beam_size) break;                                                               /* :2046-2048 */
    }
    qsort(beam, nbeam, sizeof(bcand), bcand_by_logprob);
    for (size_t c = 0; c < nbeam; ++c) { beam[c].finished = 1; beam[c].order = order++; done[ndone++] = beam[c]; }   /* :2051-2057 */
    qsort(done, ndone, sizeof(bcand), bcand_by_score);                                                       /* :2060-2063 */
    const int nres = (int)(ndone < (size_t)beam_size ? ndone : (size_t)beam_size);
    for (int i = 0; i < nres; ++i) {
        out_lens[i] = done[i].n;
        for (int t = 0; t < done[i].n; ++t) out_tokens[(size_t)i * (size_t)cap_tok + (size_t)t] = done[i].toks[t];
        if (out_logprob) out_logprob[i] = done[i].log_prob;
        if (out_score) out_score[i] = done[i].score;
        if (out_finished) out_finished[i] = done[i].finished;
    }
    for (size_t i = 0; i < ndone; ++i) free(done[i].toks);
    free(beam); free(next); free(done); free(logits); free(eprob); free(etok);
    return nres;
}

typedef struct { const tio_model* m; tio_scratch* s; } levelB_ctx;
static size_t levelB_seq_logits(void* ctx, const int32_t* prompt, int n_prompt, const int32_t* toks, int n, float* logits) {
    levelB_ctx* c = (levelB_ctx*)ctx;   /* the whole sequence again, from an empty cache */
    const int len = n_prompt + n;
    for (int i = 0; i < len; ++i) step_B(c->m, c->s, (size_t)i, i < n_prompt ? prompt[i] : toks[i - n_prompt], i == len - 1 ? logits : NULL);
    return c->m->vocab;
}
int tio_beam_search(const tio_model* m, const int32_t* prompt, int n_prompt, int max_new, int beam_size, float temperature, int top_k,
                    float top_p, float length_penalty, int eos_token, int32_t* out_tokens, int32_t* out_lens, float* out_logprob,
                    float* out_score, int32_t* out_finished) {
    if (!m || !m->lm_head || !m->tok_emb || n_prompt <= 0 || max_new < 0 || beam_size <= 0) return -1;
    tio_scratch s;
    scratch_init(m, &s, (size_t)n_prompt + (size_t)(max_new > 0 ? max_new : 1));
    levelB_ctx ctx = {m, &s};
    const int n = beam_search_core(levelB_seq_logits, &ctx, m->vocab, prompt, n_prompt, max_new, beam_size, temperature, top_k, top_p,
                                   length_penalty, eos_token, out_tokens, out_lens, out_logprob, out_score, out_finished);
    scratch_free(m, &s);
    return n;
}

void tio_logprobs(const float* logits, size_t n, size_t vocab, const int32_t* tokens, float* out) {   /* :919-944 */
    for (size_t pos = 0; pos < n; ++pos) {
        const float* row = logits + pos * vocab;
        float mx = row[0];
        for (size_t v = 1; v < vocab; ++v) mx = row[v] > mx ? row[v] : mx;
        float se = 0.0f;
        for (size_t v = 0; v < vocab; ++v) se += expf(row[v] - mx);
        const int tok = tokens[pos];
        out[pos] = (tok < 0 || (size_t)tok >= vocab) ? -20.0f : row[tok] - mx - logf(se);
    }
}

/* ------------------------------------------------------------------------------------------
 * Level C: the literal path that benchmarks/benchmark_inference runs (SURVEY 8c oracle-C).
 *   model      create_test_model, benchmarks/benchmark_inference.cpp:145-225 (no o_proj, no gate, no norms)
 *   quantize   Quantizer::quantize_model :89-118 -> integer tensors; matmul casts them WITHOUT scale
 *              (convert_dtype, src/core/tensor_engine.cpp:2218-2253)
 *   forward    placeholder embeddings 0.1f*(i%100) over the flat [1,T,H] index (:1444-1448, :1508-1512);
 *              attention fall-back returns its input so x <- x + x (:293-296, :264);
 *              x <- x + relu(x . Wup) . Wdown (:376-401, :276); logits = x . lm_head (:1472, :1541)
 *   generate   :734-802 with top_k = 1
 * ---------------------------------------------------------------------------------------- */

/* Greedy pick of sample_next_token with top_k = 1 (:1585-1598): the logits are paired with their
 * index, std::sort'ed descending by value, and element 0 of the sorted array survives.  With exact
 * ties (the ramp-filled lm_head of the benchmark model has identical columns v and v+500) WHICH of
 * the tied maxima lands at position 0 is decided by libstdc++'s introsort, so its algorithm
 * (bits/stl_algo.h: __introsort_loop / __move_median_to_first / __unguarded_partition /
 * __final_insertion_sort, threshold 16, and the heap-sort fall-back once the depth limit
 * 2*floor(log2 n) is exhausted -- which the ramp-shaped benchmark logits DO trigger) is restated here
 * move for move. */
typedef struct { float v; int idx; } lpair;
#define LGT(a, b) ((a).v > (b).v) /* the comparator: a.first > b.first */

static void lp_swap(lpair* a, lpair* b) { lpair t = *a; *a = *b; *b = t; }

static void lp_unguarded_linear_insert(lpair* last) {
    lpair val = *last;
    lpair* next = last - 1;
    while (LGT(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}

static void lp_insertion_sort(lpair* first, lpair* last) {
    if (first == last) return;
    for (lpair* i = first + 1; i != last; ++i) {
        if (LGT(*i, *first)) {
            lpair val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(lpair));
            *first = val;
        } else {
            lp_unguarded_linear_insert(i);
        }
    }
}

/* heap-sort fall-back: std::__partial_sort(first, last, last) = __make_heap + __sort_heap
 * (bits/stl_heap.h: __adjust_heap / __push_heap / __pop_heap) */
static void lp_adjust_heap(lpair* first, long hole, long len, lpair value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LGT(first[child], first[child - 1])) --child;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    long parent = (hole - 1) / 2;
    while (hole > top && LGT(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void lp_heap_sort(lpair* first, lpair* last) {
    long len = last - first;
    if (len >= 2) {
        for (long parent = (len - 2) / 2;; --parent) {
            lp_adjust_heap(first, parent, len, first[parent]);
            if (parent == 0) break;
        }
    }
    while (last - first > 1) {
        --last;
        lpair value = *last;
        *last = *first;
        lp_adjust_heap(first, 0, last - first, value);
    }
}

static int lp_introsort_loop(lpair* first, lpair* last, long depth) {
    while (last - first > 16) {
        if (depth == 0) { lp_heap_sort(first, last); return 0; }
        --depth;
        lpair* mid = first + (last - first) / 2;
        lpair *a = first + 1, *b = mid, *c = last - 1, *result = first;
        if (LGT(*a, *b)) {
            if (LGT(*b, *c)) lp_swap(result, b);
            else if (LGT(*a, *c)) lp_swap(result, c);
            else lp_swap(result, a);
        } else if (LGT(*a, *c)) lp_swap(result, a);
        else if (LGT(*b, *c)) lp_swap(result, c);
        else lp_swap(result, b);
        lpair *lo = first + 1, *hi = last, *pivot = first;
        for (;;) {
            while (LGT(*lo, *pivot)) ++lo;
            --hi;
            while (LGT(*pivot, *hi)) --hi;
            if (!(lo < hi)) break;
            lp_swap(lo, hi);
            ++lo;
        }
        if (lp_introsort_loop(lo, last, depth) != 0) return -1;
        last = lo;
    }
    return 0;
}

/* returns the index std::sort leaves at position 0, or -1 if the heap-sort fall-back would run */
static int greedy_pick_stdsort(const float* logits, size_t n) {
    lpair* p = (lpair*)malloc(n * sizeof(lpair));
    for (size_t i = 0; i < n; ++i) { p[i].v = logits[i]; p[i].idx = (int)i; }
    long lg = 0;
    for (size_t m = n; m > 1; m >>= 1) ++lg;
    int rc = lp_introsort_loop(p, p + n, 2 * lg);
    if (rc == 0) {
        if (n > 16) {
            lp_insertion_sort(p, p + 16);
            for (lpair* i = p + 16; i != p + n; ++i) lp_unguarded_linear_insert(i);
        } else {
            lp_insertion_sort(p, p + n);
        }
    }
    int best = rc == 0 ? p[0].idx : -1;
    free(p);
    return best;
}

static float* ramp(size_t n, size_t shift, size_t mod, float amp) {
    float* d = (float*)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; ++i) d[i] = ((float)((i + shift) % mod) / (float)mod - 0.5f) * amp;
    return d;
}

/* quantize_tensor :36-64 followed by the unscaled cast of convert_dtype: W <- float(q(W)) */
static void literal_quant_inplace(float* w, size_t n, int qtype) {
    if (qtype != TIO_QINT8 && qtype != TIO_QINT4) return;
    float scale, zp;
    tio_quant_info(w, n, qtype, 1, &scale, &zp);
    if (qtype == TIO_QINT8) {
        int8_t* q = (int8_t*)malloc(n);
        tio_quantize_int8(w, q, n, scale, zp);
        for (size_t i = 0; i < n; ++i) w[i] = (float)q[i];
        free(q);
    } else {
        int32_t* q = (int32_t*)malloc(n * sizeof(int32_t));
        tio_quantize_int4(w, q, n, scale, zp);
        for (size_t i = 0; i < n; ++i) w[i] = (float)q[i];
        free(q);
    }
}

static void literal_forward_rows(size_t T, size_t H, size_t I, size_t V, size_t L, const float* up, const float* down,
                                 const float* lm, float* logits, int all_rows) {
    float* x = (float*)malloc(T * H * sizeof(float));
    float* pa = (float*)malloc(T * H * sizeof(float));
    float* u = (float*)malloc(T * I * sizeof(float));
    float* f = (float*)malloc(T * H * sizeof(float));
    for (size_t i = 0; i < T * H; ++i) x[i] = 0.1f * (float)(i % 100);
    for (size_t l = 0; l < L; ++l) {
        tio_add(x, x, pa, T * H);
        tio_matmul(pa, up, u, T, H, I);
        tio_relu(u, u, T * I);
        tio_matmul(u, down, f, T, I, H);
        tio_add(pa, f, x, T * H);
    }
    if (all_rows) tio_matmul(x, lm, logits, T, H, V);      /* forward_pass returns [1, T, V] */
    else tio_matmul(x + (T - 1) * H, lm, logits, 1, H, V); /* only the last row is sampled, :1571-1576 */
    free(x); free(pa); free(u); free(f);
}
static void literal_forward(size_t T, size_t H, size_t I, size_t V, size_t L, const float* up, const float* down,
                            const float* lm, float* last_logits) {
    literal_forward_rows(T, H, I, V, L, up, down, lm, last_logits, 0);
}

int tio_generate_literal(int vocab, int hidden, int layers, int qtype, const int32_t* prompt, int n_prompt,
                         int n_new, int32_t* out_tokens, float* last_logits) {
    (void)prompt; /* token ids never reach the arithmetic */
    if (n_prompt <= 0 || n_new < 0) return -1;
    const size_t V = vocab, H = hidden, L = layers, I = H * 4;
    /* every layer holds the same fill, so one copy of each matrix serves all layers */
    float* up = ramp(H * I, 0, 200, 0.02f);
    float* down = ramp(I * H, 0, 200, 0.02f);
    float* lm = ramp(H * V, 0, 500, 0.01f);
    literal_quant_inplace(up, H * I, qtype);
    literal_quant_inplace(down, I * H, qtype);
    literal_quant_inplace(lm, H * V, qtype);
    float* logits = (float*)malloc(V * sizeof(float));
    literal_forward((size_t)n_prompt, H, I, V, L, up, down, lm, logits); /* prefill, :749 */
    int produced = 0;
    size_t total = (size_t)n_prompt;
    for (int i = 0; i < n_new; ++i) {
        int best = greedy_pick_stdsort(logits, V);
        if (best < 0) { produced = -3; break; }
        out_tokens[produced++] = best;
        ++total;
        if (best == 2) break;            /* :760 */
        if (total >= 2048) break;        /* max_sequence_length, :767 */
        literal_forward(1, H, I, V, L, up, down, lm, logits); /* decode, :774 */
    }
    if (last_logits) memcpy(last_logits, logits, V * sizeof(float));
    free(up); free(down); free(lm); free(logits);
    return produced;
}

/* generate_beam_search on the literal benchmark model (level C): forward_pass over the whole candidate, whose logits depend on
 * the sequence LENGTH only (token ids never reach the arithmetic).  QUIRK restated here and nowhere else:
 * beam_search_decode takes `total_size / shape[0]` for the vocabulary (:1964-1968); forward_pass returns [1, T, V], so the
 * reference soft-maxes over the T * V logits of ALL positions and can emit "token ids" up to T * V - 1.  The level-B
 * restatement (tio_beam_search) and the GPU engine use the last position's V logits, which is what the comment at :1963 intends. */
typedef struct { size_t H, I, V, L; const float *up, *down, *lm; } literal_ctx;
static size_t literal_seq_logits(void* ctx, const int32_t* prompt, int n_prompt, const int32_t* toks, int n, float* logits) {
    (void)prompt; (void)toks;
    const literal_ctx* c = (const literal_ctx*)ctx;
    const size_t T = (size_t)(n_prompt + n);
    literal_forward_rows(T, c->H, c->I, c->V, c->L, c->up, c->down, c->lm, logits, 1);
    return T * c->V;
}
int tio_beam_search_literal(int vocab, int hidden, int layers, int qtype, const int32_t* prompt, int n_prompt, int max_new, int beam_size,
                            float temperature, int top_k, float top_p, float length_penalty, int32_t* out_tokens, int32_t* out_lens,
                            float* out_avg_logprob, int32_t* out_finished) {
    if (n_prompt <= 0 || max_new < 0 || beam_size <= 0) return -1;
    const size_t V = vocab, H = hidden, L = layers, I = H * 4;
    float* up = ramp(H * I, 0, 200, 0.02f);
    float* down = ramp(I * H, 0, 200, 0.02f);
    float* lm = ramp(H * V, 0, 500, 0.01f);
    literal_quant_inplace(up, H * I, qtype);
    literal_quant_inplace(down, I * H, qtype);
    literal_quant_inplace(lm, H * V, qtype);
    literal_ctx ctx = {H, I, V, L, up, down, lm};
    float* lp = (float*)calloc((size_t)beam_size, sizeof(float));
    const int n = beam_search_core(literal_seq_logits, &ctx, (size_t)(n_prompt + (max_new > 0 ? max_new : 1)) * V, prompt, n_prompt, max_new, beam_size, temperature, top_k, top_p, length_penalty,
                                   2 /* InferenceConfig::eos_token_id default */, out_tokens, out_lens, lp, NULL, out_finished);
    for (int i = 0; i < n; ++i) out_avg_logprob[i] = lp[i] / (float)out_lens[i];   /* what GenerationResult::logprobs holds, :862-865 */
    free(lp); free(up); free(down); free(lm);
    return n;
}

/* compute_logprobs on the literal benchmark model (level C): forward_pass over the tokens ([1, T, V] logits, which depend on T
 * only) and the log-softmax of every position (:919-944) -- in the compiled reference InferenceEngine::compute_logprobs itself */
int tio_logprobs_literal(int vocab, int hidden, int layers, int qtype, const int32_t* tokens, int n, float* out) {
    if (n <= 0) return -1;
    const size_t V = vocab, H = hidden, L = layers, I = H * 4;
    float* up = ramp(H * I, 0, 200, 0.02f);
    float* down = ramp(I * H, 0, 200, 0.02f);
    float* lm = ramp(H * V, 0, 500, 0.01f);
    literal_quant_inplace(up, H * I, qtype);
    literal_quant_inplace(down, I * H, qtype);
    literal_quant_inplace(lm, H * V, qtype);
    float* logits = (float*)malloc((size_t)n * V * sizeof(float));
    literal_forward_rows((size_t)n, H, I, V, L, up, down, lm, logits, 1);
    tio_logprobs(logits, (size_t)n, V, tokens, out);
    free(logits); free(up); free(down); free(lm);
    return n;
}

/* generate() on the literal benchmark model with the sampling pipeline switched on (temperature, top_k, top_p as given): every
 * token goes through tio_sample with the uniform u.  With a top_p so small that the nucleus is ONE token the draw does not depend
 * on the generator, which is how the temperature / top-k / softmax / top-p stages are pinned against the compiled reference
 * (whose std::mt19937 is time-seeded, :472). */
int tio_generate_literal_sampled(int vocab, int hidden, int layers, int qtype, const int32_t* prompt, int n_prompt, int n_new,
                                 float temperature, int top_k, float top_p, float u, int32_t* out_tokens) {
    (void)prompt;
    if (n_prompt <= 0 || n_new < 0) return -1;
    const size_t V = vocab, H = hidden, L = layers, I = H * 4;
    float* up = ramp(H * I, 0, 200, 0.02f);
    float* down = ramp(I * H, 0, 200, 0.02f);
    float* lm = ramp(H * V, 0, 500, 0.01f);
    literal_quant_inplace(up, H * I, qtype);
    literal_quant_inplace(down, I * H, qtype);
    literal_quant_inplace(lm, H * V, qtype);
    float* logits = (float*)malloc(V * sizeof(float));
    literal_forward((size_t)n_prompt, H, I, V, L, up, down, lm, logits);
    int produced = 0;
    size_t total = (size_t)n_prompt;
    for (int i = 0; i < n_new; ++i) {
        const int tok = tio_sample(logits, V, temperature, top_k, top_p, u, NULL);
        out_tokens[produced++] = tok;
        ++total;
        if (tok == 2 || total >= 2048) break;
        literal_forward(1, H, I, V, L, up, down, lm, logits);
    }
    free(up); free(down); free(lm); free(logits);
    return produced;
}
