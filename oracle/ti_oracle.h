/*
 * ti_oracle.h -- C ABI shared by the two CPU oracles of the decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * these libraries, and only as the checker or the reported CPU baseline.
 *
 * Two shared objects export exactly this interface:
 *   oracle/libti_oracle.so          plain-C restatement (ti_oracle.c), file:line cited per fn
 *   oracle/_ref/libti_ref.so        the UNMODIFIED reference sources compiled where they lie
 *                                   (/root/reference/src) + ref_harness.cpp calling its C++ API
 * so a test can run the same call against either and compare them bit for bit.
 *
 * All tensors are dense row-major fp32 unless stated.  Weights are [in, out] (y = x . W), the
 * reference's convention (src/core/tensor_engine.cpp:565-573).
 */
#ifndef TI_ORACLE_H
#define TI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* quantization types, numerically equal to turboinfer::optimize::QuantizationType
 * (include/turboinfer/optimize/quantization.hpp:24-29) */
enum { TIO_QINT8 = 0, TIO_QINT4 = 1, TIO_QNONE = 3 };

/* 0 = restatement, 1 = compiled reference */
int tio_kind(void);

/* ---- level A: single ops ------------------------------------------------------------- */

/* Quantizer::calculate_quantization_info  (src/optimize/quantization.cpp:335-394) */
int tio_quant_info(const float* x, size_t n, int qtype, int symmetric, float* scale, float* zero_point);
/* quantize_to_int8 (:662-674) / quantize_to_int4 (:676-693); INT4 is one value per int32 */
void tio_quantize_int8(const float* x, int8_t* q, size_t n, float scale, float zero_point);
void tio_quantize_int4(const float* x, int32_t* q, size_t n, float scale, float zero_point);
/* dequantize_from_int8 (:695-703) / dequantize_from_int4 (:705-713) */
void tio_dequantize_int8(const int8_t* q, float* x, size_t n, float scale, float zero_point);
void tio_dequantize_int4(const int32_t* q, float* x, size_t n, float scale, float zero_point);

/* TensorEngine::matmul on 2-D fp32 (src/core/tensor_engine.cpp:490-592) */
void tio_matmul(const float* a, const float* b, float* c, size_t M, size_t K, size_t N);
/* TensorEngine::rms_norm (:1452-1508) */
void tio_rms_norm(const float* x, const float* w, float* y, size_t rows, size_t H, float eps);
/* TensorEngine::apply_rope (:1510-1624).  ndim 3: x is [B,T,D]; ndim 4: x is [B,nh,T,D].
 * pos is [T] (pos_2d = 0) or [B,T] (pos_2d = 1). */
void tio_rope(const float* x, const float* pos, float* y, size_t B, size_t nh, size_t T, size_t D,
              int ndim, int pos_2d, float theta);
/* silu (:900-923), relu (:828-869), add (:1626-1678), multiply (:1680-1743) */
void tio_silu(const float* x, float* y, size_t n);
void tio_relu(const float* x, float* y, size_t n);
void tio_add(const float* a, const float* b, float* y, size_t n);
void tio_mul(const float* a, const float* b, float* y, size_t n);
/* TensorEngine::softmax (:925-1043).  The restatement follows the SCALAR branch (:1017-1033);
 * the compiled reference takes its AVX2 branch when n >= 16 (12 % off). */
void tio_softmax(const float* x, float* y, size_t rows, size_t n, float temperature);
/* TensorEngine::attention_fast_incremental (:1254-1388): q [B,1,H], k/v [B,t,H] -> out [B,1,H] */
void tio_attention_fast_incremental(const float* q, const float* k, const float* v, float* out,
                                    size_t B, size_t t, size_t H);
/* TensorEngine::multi_head_attention with q_len == 1 (:1149-1252 -> :1077-1081 -> :1254) */
void tio_multi_head_attention(const float* q, const float* k, const float* v, float* out,
                              size_t B, size_t t, size_t H, size_t nh);

/* ---- level B: the decode loop, intended dataflow (SURVEY 8c oracle-B) ------------------ */

typedef struct tio_model {
    int32_t vocab, hidden, layers, heads, inter;
    float rope_theta;      /* 10000 */
    float rms_eps;         /* 1e-5, TensorEngine::rms_norm default */
    int32_t attn_mode;     /* 0: one head over the whole hidden dim (literal)
                              1: heads heads via multi_head_attention */
    int32_t rope_mode;     /* 0: none (literal); 1: per head (4-D apply_rope, d = head_dim);
                              2: whole hidden (3-D apply_rope, d = hidden) */
    const float* tok_emb;  /* [vocab, hidden] */
    const float* out_norm; /* [hidden] or NULL */
    const float* lm_head;  /* [hidden, vocab] */
    /* per layer, arrays of `layers` pointers; a NULL entry means "tensor absent" and triggers the
     * reference's null-weight fall-backs (src/model/inference_engine.cpp:293-296, :377-380, :392-395) */
    const float* const* attn_norm;
    const float* const* wq;
    const float* const* wk;
    const float* const* wv;
    const float* const* wo;
    const float* const* ffn_norm;
    const float* const* w_up;
    const float* const* w_gate;
    const float* const* w_down;
} tio_model;

/* Greedy generation: prompt tokens are fed one per step through the incremental path
 * (forward_incremental, src/model/inference_engine.cpp:244-279), then n_new tokens are produced
 * with argmax (= top_k 1 branch of sample_next_token, :1585-1598).  out_tokens receives n_new ids;
 * logits_out (may be NULL) receives [n_new, vocab] -- the logits each token was picked from.
 * stop_on_eos != 0 reproduces the hard-coded `== 2` stop (:760).  Returns tokens produced, <0 on error. */
int tio_decode_greedy(const tio_model* m, const int32_t* prompt, int n_prompt, int n_new,
                      int stop_on_eos, int32_t* out_tokens, float* logits_out);

/* Same generation, with the wall-clock time of every forward pass (steady clock around one step: n_prompt + n_new - 1
 * passes) written to step_seconds[0 .. cap) -- what bench.py's CPU arm reports; weight set-up (the deep copies of
 * initialize_model) stays outside the timed passes.  Returns the number of passes timed, < 0 on error. */
int tio_decode_greedy_timed(const tio_model* m, const int32_t* prompt, int n_prompt, int n_new,
                            int32_t* out_tokens, double* step_seconds, int cap);

/* ---- sampling ----------------------------------------------------------------
 * InferenceEngine::sample_next_token (src/model/inference_engine.cpp:1554-1673) on one row of logits: temperature
 * (:1577-1582), top-k (:1584-1598), softmax (:1600-1612), top-p (:1614-1648), inverse CDF (:1650-1672), every sum
 * sequential in the reference's order.  The reference draws its uniform from a time-seeded std::mt19937 (:472), which
 * cannot be reproduced: the uniform is an ARGUMENT here (tio_uniform = the counter-based generator the GPU engine uses),
 * so this part of the oracle is pinned against the compiled reference only where no randomness enters (top_k = 1).
 * std::sort is unstable: among equal logits / probabilities the reference's order is unspecified; restated as value
 * descending, index ascending.  Returns the token; *logprob (may be NULL) = log of its final probability. */
int tio_sample(const float* logits, size_t vocab, float temperature, int top_k, float top_p, float u, float* logprob);
float tio_uniform(uint64_t seed, uint64_t step);
/* ---- beam search ---------------------------------------------------------------
 * beam_search_decode's expansion of one candidate (:1964-2005) on one row of logits: logits / temperature (:1972-1976), softmax
 * over the vocabulary (:1798-1819), top-k on the probabilities + renormalise (:1821-1856), top-p + renormalise (:1858-1909), the
 * beam_size most probable tokens with probability > 0 (:1990-2005); every sum sequential in the reference's order, ties restated
 * as value descending, index ascending.  Returns the count; probs / tokens [beam_size].  (Port only.) */
int tio_beam_expand(const float* logits, size_t vocab, float temperature, int top_k, float top_p, int beam_size, float* probs, int32_t* tokens);
/* beam_search_decode (:1912-2069) + the result conversion of generate_beam_search (:849-857) over the level-B forward pass: every
 * candidate's whole sequence is run through the model at every step, as the reference does (:1961).  The candidates leave the heap
 * most probable first (:1941-1944; ties: earlier candidate), the sorts by normalised score are stable.  out_tokens
 * [beam_size][max_new] (new tokens only), best first; returns the number of results.  (Port only.) */
int tio_beam_search(const tio_model* m, const int32_t* prompt, int n_prompt, int max_new, int beam_size, float temperature, int top_k,
                    float top_p, float length_penalty, int eos_token, int32_t* out_tokens, int32_t* out_lens, float* out_logprob,
                    float* out_score, int32_t* out_finished);
/* generate() on the literal benchmark model through the SAMPLING pipeline (both oracles; in the compiled reference the engine's
 * own generate with these InferenceConfig values; u is ignored there).  Deterministic -- and therefore comparable -- when top_p
 * leaves a one-token nucleus. */
int tio_generate_literal_sampled(int vocab, int hidden, int layers, int qtype, const int32_t* prompt, int n_prompt, int n_new,
                                 float temperature, int top_k, float top_p, float u, int32_t* out_tokens);
/* compute_logprobs (:873-954) on the literal benchmark model (level C): pins tio_logprobs against the compiled reference's own
 * InferenceEngine::compute_logprobs.  Returns n (or < 0).  Both oracles. */
int tio_logprobs_literal(int vocab, int hidden, int layers, int qtype, const int32_t* tokens, int n, float* out);
/* generate_beam_search on the literal benchmark model (level C, as tio_generate_literal): in the compiled reference this is
 * InferenceEngine::generate_beam_search itself, which pins the expansion arithmetic and the bookkeeping of the restatement.
 * out_avg_logprob = cumulative log-probability / new tokens (GenerationResult::logprobs, :862-865).  Both oracles. */
int tio_beam_search_literal(int vocab, int hidden, int layers, int qtype, const int32_t* prompt, int n_prompt, int max_new, int beam_size,
                            float temperature, int top_k, float top_p, float length_penalty, int32_t* out_tokens, int32_t* out_lens,
                            float* out_avg_logprob, int32_t* out_finished);
/* compute_logprobs (:873-954) on precomputed logits [n, vocab]: out[pos] = logit[token] - max - log(sum exp(logit - max)),
 * -20 for a token id outside the vocabulary */
void tio_logprobs(const float* logits, size_t n, size_t vocab, const int32_t* tokens, float* out);

/* ---- level C: the literal path of benchmarks/benchmark_inference (SURVEY 8c oracle-C) --- */
/* create_test_model(vocab, hidden, layers) (benchmarks/benchmark_inference.cpp:145-225) run through
 * InferenceEngine::generate with top_k = 1; qtype TIO_QNONE / TIO_QINT8 / TIO_QINT4 goes through
 * Quantizer::quantize_model first (unscaled integer weights). */
int tio_generate_literal(int vocab, int hidden, int layers, int qtype, const int32_t* prompt, int n_prompt,
                         int n_new, int32_t* out_tokens, float* last_logits);

#ifdef __cplusplus
}
#endif
#endif
