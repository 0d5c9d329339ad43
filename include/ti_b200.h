/*
 * ti_b200.h -- C ABI of libturboinfer_b200.so, the B200 (sm_100a) implementation of TurboInfer's
 * token-generation hot path.
 *
 * This is the drop-in boundary.  The reference has no C ABI of its own: its seam is
 * the C++ class surface (TensorEngine / Quantizer / InferenceEngine).  The host-side C++ classes in
 * include/turboinfer/ keep that surface and are implemented ONLY in terms of the entry points below;
 * each entry point cites the reference interface it replaces (file:line relative to /root/reference).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; ti_b200_last_error() returns the
 *    thread-local message.  Nothing throws across this boundary.  The host C++ wrappers turn a
 *    non-zero status into std::runtime_error like the reference does (src/core/tensor_engine.cpp:492-494).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails loudly.
 *  - "host" pointers are ordinary host memory (H2D / D2H copies happen inside the call, synchronous
 *    at return, like the reference's value-in / value-out ops); "_dev" variants take device pointers
 *    obtained from ti_b200_malloc and only enqueue work on the library stream.
 *  - weights are [in, out] row-major (y = x . W), the reference convention (tensor_engine.cpp:565-573).
 *  - handles are opaque 64-bit values; sizes are size_t; no C++ / torch types cross.
 */
#ifndef TI_B200_H
#define TI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TI_B200_ABI_VERSION 1

/* numerically equal to turboinfer::optimize::QuantizationType (include/turboinfer/optimize/quantization.hpp:24-29) */
enum { TI_Q_INT8 = 0, TI_Q_INT4 = 1, TI_Q_NONE = 3 };

typedef uint64_t ti_qweight_t; /* packed INT4/INT8 weight resident in HBM */
typedef uint64_t ti_model_t;   /* device-resident decoder (weights + paged KV cache + decode graph) */
typedef uint64_t ti_kv_t;      /* stand-alone paged KV cache (the reference's KVCache as an object of its own) */

/* ---- lifecycle / errors ---------------------------------------------------------------------
 * replaces TensorEngine::TensorEngine(ComputeDevice) / initialize / gpu_available / device_info
 * (include/turboinfer/core/tensor_engine.hpp:42-60, src/core/tensor_engine.cpp:323-469) */
int ti_b200_abi_version(void);
int ti_b200_init(int device);
int ti_b200_shutdown(void);
int ti_b200_device_count(int* n);
int ti_b200_device_info(char* buf, size_t cap);
const char* ti_b200_last_error(void);
int ti_b200_sync(void);

/* ---- tensor parallelism (the reference has none) --------------------------------------------------
 * One process per GPU.  Rank 0 obtains an id with ti_b200_tp_unique_id and hands it to the other ranks by any means
 * (the tests and the bench use torch.distributed); every rank then calls ti_b200_tp_init.  A model created with
 * cfg.reserved[1] = nranks is sharded: column-parallel q/k/v (whole heads per rank) and gate/up, row-parallel o and
 * down, one NCCL all-reduce (fp32 sum over NVLink) after each row-parallel GEMV; embeddings, norms and lm_head are
 * replicated.  Every rank is given the WHOLE fp32 tensors: quantization parameters come from the whole tensor
 * (quantize first, then shard), so the integers are the reference's. */
int ti_b200_tp_unique_id(uint8_t* out, size_t cap);   /* cap >= 128 */
int ti_b200_tp_init(int nranks, int rank, const uint8_t* unique_id, size_t id_bytes);

/* ---- raw device memory (for callers that keep activations resident) --------------------------- */
int ti_b200_malloc(void** dev, size_t bytes);
int ti_b200_free(void* dev);
int ti_b200_memcpy_h2d(void* dev, const void* host, size_t bytes);
int ti_b200_memcpy_d2h(void* host, const void* dev, size_t bytes);

/* ---- quantizer ---------------------------------------------------------------------------------
 * ti_b200_quant_info    <- Quantizer::calculate_quantization_info (src/optimize/quantization.cpp:335-394)
 * ti_b200_quantize      <- quantize_to_int8 / quantize_to_int4 (:662-693); q_out is int8[n] or int32[n]
 *                          exactly like the reference's tensors (INT4 = one value per int32)
 * ti_b200_dequantize    <- dequantize_from_int8 / dequantize_from_int4 (:695-713)
 * ti_b200_quantize_pack <- Quantizer::quantize_tensor (:36-64) followed by the device re-tiling:
 *                          min/max scan, scale/zero-point, quantize, pack 2 nibbles (INT4) or 1 byte
 *                          (INT8) per element into the GEMV streaming layout, all on the GPU.
 * ti_b200_qweight_unpack -> the integers back in the reference's [K,N] int32 order, for the
 *                          unpack(pack(q)) == reference-ints check.
 * All integer results and the scale / zero-point bits are bit-exact with the reference. */
int ti_b200_quant_info(const float* x_host, size_t n, int qtype, int symmetric, float* scale, float* zero_point);
int ti_b200_quantize(const float* x_host, size_t n, int qtype, float scale, float zero_point, void* q_out_host);
int ti_b200_dequantize(const void* q_host, size_t n, int qtype, float scale, float zero_point, float* x_out_host);
int ti_b200_quantize_pack(const float* w_host, size_t K, size_t N, int qtype, int symmetric, ti_qweight_t* out);
int ti_b200_qweight_info(ti_qweight_t w, size_t* K, size_t* N, int* qtype, float* scale, float* zero_point,
                         size_t* packed_bytes);
int ti_b200_qweight_unpack(ti_qweight_t w, int32_t* q_out_host);
int ti_b200_qweight_free(ti_qweight_t w);

/* ---- tensor-engine ops (host tensors in, host tensors out) --------------------------------------
 * ti_b200_gemv_q        <- TensorEngine::matmul(x, dequantize_tensor(quantize_tensor(W)))  (tensor_engine.cpp:490-640)
 *                          y[b,:] = scale * (x[b,:] . q) (+ zero-point term); rows = batch of independent GEMVs
 * ti_b200_matmul_f32    <- TensorEngine::matmul on fp32 2-D tensors, reproducing the reference build's
 *                          per-element order of roundings (bit-exact, see oracle/ti_oracle.c tio_matmul)
 * ti_b200_rms_norm      <- TensorEngine::rms_norm (:1452-1508)
 * ti_b200_rope          <- TensorEngine::apply_rope (:1510-1624); ndim 3: [B,T,D], ndim 4: [B,nh,T,D]
 * ti_b200_silu / relu / add / mul <- :900-923, :828-869, :1626-1678, :1680-1743
 * ti_b200_silu_mul      <- multiply(up, silu(gate)), the SwiGLU of compute_ffn (inference_engine.cpp:389-391)
 * ti_b200_softmax       <- TensorEngine::softmax, scalar branch (:1017-1033)
 * ti_b200_attention_decode <- attention_fast_incremental (:1254-1388) when num_heads == 1,
 *                          multi_head_attention with q_len 1 (:1149-1252) otherwise; q [B,1,H], k/v [B,t,H] */
int ti_b200_gemv_q(ti_qweight_t w, const float* x_host, float* y_host, size_t rows);
/* ti_b200_gemm_q        <- TensorEngine::matmul with M >= 32 rows, i.e. simd_gemm_float (tensor_engine.cpp:191-255), on the
 *                          same quantized weight: the prefill / batched-decode contraction on the tcgen05 tensor cores
 *                          (INT8 digit planes x INT8 / unpacked INT4, exact int32 accumulation in TMEM).  Every output row
 *                          is bit-identical to ti_b200_gemv_q of that row. */
int ti_b200_gemm_q(ti_qweight_t w, const float* x_host, float* y_host, size_t rows);
int ti_b200_matmul_f32(const float* a_host, const float* b_host, float* c_host, size_t M, size_t K, size_t N);
int ti_b200_rms_norm(const float* x_host, const float* w_host, float* y_host, size_t rows, size_t H, float eps);
int ti_b200_rope(const float* x_host, const float* pos_host, float* y_host, size_t B, size_t nh, size_t T, size_t D,
                 int ndim, int pos_2d, float theta);
int ti_b200_silu(const float* x_host, float* y_host, size_t n);
int ti_b200_relu(const float* x_host, float* y_host, size_t n);
int ti_b200_add(const float* a_host, const float* b_host, float* y_host, size_t n);
int ti_b200_mul(const float* a_host, const float* b_host, float* y_host, size_t n);
int ti_b200_silu_mul(const float* gate_host, const float* up_host, float* y_host, size_t n);
int ti_b200_softmax(const float* x_host, float* y_host, size_t rows, size_t n, float temperature);
int ti_b200_attention_decode(const float* q_host, const float* k_host, const float* v_host, float* out_host,
                             size_t B, size_t t, size_t H, size_t num_heads);

/* ---- fused GEMV (north_star item 4) -----------------------------------------------------------------
 * ti_b200_quantize_pack_fused: up to three [K, n_i] matrices sharing x packed as ONE streaming weight, each keeping its own
 *                           per-tensor scale: concatenated (q | k | v) or, interleave = 1, two sources column-interleaved
 *                           (gate_i, up_i) for the SwiGLU epilogue -- what initialize_model's separate tensors become here.
 * ti_b200_gemv_q_ex:        the decode GEMV with its fused prologue / epilogues: norm_w_host != NULL applies rms_norm(x, norm_w, eps)
 *                           (tensor_engine.cpp:1452-1508) in front; epilogue = TI_EPI_STORE, TI_EPI_RESIDUAL (y + resid, :1626-1678),
 *                           TI_EPI_SWIGLU (up_i * silu(gate_i), N / 2 outputs, :900-923 + :1680-1743) or TI_EPI_RELU (:828-869). */
enum { TI_EPI_STORE = 0, TI_EPI_RESIDUAL = 1, TI_EPI_SWIGLU = 2, TI_EPI_RELU = 3 };
int ti_b200_quantize_pack_fused(const float* const* w_host, const size_t* n_cols, int32_t n_src, int32_t interleave, size_t K,
                                int qtype, ti_qweight_t* out);
int ti_b200_gemv_q_ex(ti_qweight_t w, const float* x_host, float* y_host, int32_t epilogue, const float* resid_host,
                      const float* norm_w_host, float eps);

/* ---- KV cache manager ---------------------------------------------------------------------------------
 * The reference's KVCache (src/model/inference_engine.cpp:25-172) as a device-resident, paged object: per layer two pools
 * [page][page_tokens][heads * head_dim] fp32 behind one page table.
 *   ti_b200_kv_create    <- KVCache::initialize (:40-55), batch 1
 *   ti_b200_kv_reset     <- KVCache::reset (:57-69): lengths to 0 (no zero-fill needed: nothing reads beyond the length)
 *   ti_b200_kv_append    <- update_incremental / update (:78-172): k_new / v_new [heads, new_tokens, head_dim] appended at the
 *                           layer's length; "Layer index out of bounds for KV cache" and "KV cache overflow: sequence too long"
 *                           are the reference's errors.  Each layer has its OWN length (the reference advances one shared counter
 *                           on every layer's update, which overflows a model of L layers after max/L tokens).
 *   ti_b200_kv_read      <- the (full_keys, full_values) pair update_incremental returns: [heads, length, head_dim]
 *   ti_b200_kv_attention <- attention over the cached tokens read in place (no copy-out): multi_head_attention with q_len 1
 *                           (tensor_engine.cpp:1149-1252); heads = 1 is attention_fast_incremental (:1254-1388) */
int ti_b200_kv_create(int32_t layers, int32_t heads, int32_t head_dim, int32_t max_seq, int32_t page_tokens, ti_kv_t* out);
int ti_b200_kv_destroy(ti_kv_t kv);
int ti_b200_kv_reset(ti_kv_t kv);
int ti_b200_kv_length(ti_kv_t kv, int32_t layer, int32_t* current_length, int32_t* max_length);
int ti_b200_kv_append(ti_kv_t kv, int32_t layer, const float* k_new_host, const float* v_new_host, int32_t new_tokens);
int ti_b200_kv_read(ti_kv_t kv, int32_t layer, float* k_out_host, float* v_out_host);
int ti_b200_kv_attention(ti_kv_t kv, int32_t layer, const float* q_host, float* out_host);

/* device-pointer variant of the GEMV used by the bench (inputs already resident in HBM) */
int ti_b200_gemv_q_dev(ti_qweight_t w, const float* x_dev, float* y_dev);

/* ---- decoder model -------------------------------------------------------------------------------
 * replaces InferenceEngine(ModelData, InferenceConfig) / initialize_model (src/model/inference_engine.cpp:480-564),
 * KVCache (:25-172), TransformerLayer::forward_incremental (:244-401), forward_pass_incremental (:1493-1552),
 * forward_pass (:1429-1491) and the greedy branch of generate / sample_next_token (:734-802, :1554-1673). */
typedef struct ti_model_config {
    int32_t vocab, hidden, layers, heads, inter;
    float rope_theta;        /* ModelMetadata::rope_theta; 10000 */
    float rms_eps;           /* TensorEngine::rms_norm default 1e-5 */
    int32_t qtype;           /* TI_Q_INT4 / TI_Q_INT8: 2-D projection weights are quantized (symmetric, per tensor) and packed;
                                TI_Q_NONE is not a decode mode of this engine */
    int32_t attn_mode;       /* 0: one head over the whole hidden dim (literal reference); 1: `heads` heads */
    int32_t rope_mode;       /* 0: none (literal); 1: per head; 2: whole hidden */
    int32_t max_seq;         /* KV-cache capacity in tokens (KVCache::max_length, reference hard-codes 2048) */
    int32_t kv_page_tokens;  /* tokens per KV page; 0 = default 64 */
    int32_t compat_literal;  /* 1: placeholder embeddings 0.1f*(i%100) and unscaled integer weights */
    int32_t reserved[7];     /* [0] = 1: per-op graph engine instead of the persistent kernel; [1] = tensor parallel degree (0/1: none) */
} ti_model_config;

int ti_b200_model_new(const ti_model_config* cfg, ti_model_t* out);
/* `name` uses the reference tensor names accepted by initialize_model (:483-563), e.g.
 * "layers.3.attention.q_proj.weight"; data is fp32 [rows, cols] ([in, out] for projections, [H] for norms:
 * rows = 1).  Absent tensors trigger the reference's null-weight fall-backs (:293-296, :377-380, :392-395). */
int ti_b200_model_set_tensor(ti_model_t m, const char* name, const float* data_host, size_t rows, size_t cols);
/* An ALREADY QUANTIZED projection / lm_head tensor -- the output of Quantizer::quantize_model (src/optimize/quantization.cpp:67-87) or
 * a tensor of a .tinq file (load_quantized_model, :208-333): int8 elements for TI_Q_INT8, int32 elements for TI_Q_INT4 (the
 * reference's storage, :683-693), row-major [rows, cols], with the (scale, zero_point) they were quantized with.  The integers are
 * packed into the streaming layout as they are (no re-quantization); qtype must equal the model's.  INT4 integers lie in [-7, 7]
 * (zero_point 0) or [0, 15] (asymmetric).  Embedding rows and norm weights stay float32 (ti_b200_model_set_tensor). */
int ti_b200_model_set_tensor_q(ti_model_t m, const char* name, const void* q_host, size_t rows, size_t cols, int qtype, float scale,
                               float zero_point);
/* same, but the fp32 source is generated on the device: uniform(-amp, amp) from a counter-based hash of
 * (seed, element index) -- for benchmark-size models whose fp32 weights would not fit host RAM. */
int ti_b200_model_set_tensor_synthetic(ti_model_t m, const char* name, size_t rows, size_t cols, uint64_t seed, float amp);
int ti_b200_model_finalize(ti_model_t m);
int ti_b200_model_free(ti_model_t m);
/* KVCache::reset (:60-69): frees the page list, length = 0 */
int ti_b200_model_reset(ti_model_t m);
int ti_b200_model_kv_length(ti_model_t m, int32_t* length);
/* which engine decodes this (finalized) model: 1 = the persistent decode kernel (one cooperative launch per generation), 0 =
 * the per-op graph engine (models with absent tensors -- the reference's null-weight fall-backs -- and the NCCL baseline) */
int ti_b200_model_engine(ti_model_t m, int32_t* persistent);
/* bytes one decode step must move at cache length t: packed weights + scales + KV read/write */
int ti_b200_model_step_bytes(ti_model_t m, int32_t t, double* weight_bytes, double* kv_bytes);
/* forward_pass_incremental for one token: appends to the KV cache; logits_host may be NULL */
int ti_b200_decode_step(ti_model_t m, int32_t token, float* logits_host, int32_t* argmax);
/* generate(): prefill `prompt`, then n_new greedy tokens (top_k = 1).  out_tokens gets the new ids, returns their
 * count in *n_out (early stop on token 2 when stop_on_eos, :760).  logits_host (optional) receives
 * [n_new, vocab]; decode_ms (optional) is the CUDA-event time of the decode loop only. */
int ti_b200_generate_greedy(ti_model_t m, const int32_t* prompt, int32_t n_prompt, int32_t n_new, int32_t stop_on_eos,
                            int32_t* out_tokens, int32_t* n_out, float* logits_host, float* decode_ms);

/* forward_pass (:1429-1491) over a prompt: resets the model's KV cache, fills it with the n tokens (tensor-core GEMM path for
 * prompts of >= 33 tokens, the decode engine otherwise) and returns the logits of the last position (last_logits_host may be NULL).
 * ti_b200_decode_step continues from there. */
int ti_b200_prefill(ti_model_t m, const int32_t* tokens, int32_t n, float* last_logits_host);

/* generate_batch() (src/model/inference_engine.cpp:804-828, a sequential loop of generate() in the reference): `batch`
 * prompts of equal length n_prompt ([batch][n_prompt]), n_new >= 1 greedy tokens each, advanced in lockstep so that the
 * weights are read once per step for all sequences (tensor-core GEMM path, one KV cache per sequence).  out_tokens:
 * [batch][n_new]; logits_last (optional): [batch][vocab] of the last step; decode_ms (optional): CUDA-event time of the
 * n_new - 1 steps after the prompt.  Single-GPU models with every projection present only. */
int ti_b200_generate_batch_greedy(ti_model_t m, const int32_t* prompts, int32_t batch, int32_t n_prompt, int32_t n_new,
                                  int32_t* out_tokens, float* logits_last, float* decode_ms);

/* ---- sampling -----------------------------------------------------------------
 * ti_b200_sample_logits    <- InferenceEngine::sample_next_token (src/model/inference_engine.cpp:1554-1673) on host logits
 *                             [rows, vocab]: temperature -> top-k -> softmax -> top-p -> inverse CDF, and the log-probability
 *                             of the pick.  The reference's RNG is a time-seeded std::mt19937; here the uniform of a pick is
 *                             a counter-based hash of (seed, step), so generations are reproducible.
 * ti_b200_generate_sampled <- generate() (:734-802) with that sampler ON THE DEVICE: the picked token feeds the next forward
 *                             pass without a logits download or a host sort; logprobs_out (optional) as include_logprobs.
 * ti_b200_compute_logprobs <- compute_logprobs (:873-954): out[pos] = log softmax(logits[pos])[tokens[pos]], -20 for an id
 *                             outside the vocabulary. */
int ti_b200_sample_logits(const float* logits_host, size_t rows, size_t vocab, float temperature, int32_t top_k, float top_p,
                          uint64_t seed, int32_t step, int32_t* tokens_out, float* logprobs_out);
int ti_b200_generate_sampled(ti_model_t m, const int32_t* prompt, int32_t n_prompt, int32_t n_new, float temperature,
                             int32_t top_k, float top_p, uint64_t seed, int32_t stop_on_eos, int32_t* out_tokens,
                             int32_t* n_out, float* logprobs_out, float* decode_ms);
int ti_b200_compute_logprobs(ti_model_t m, const int32_t* tokens, int32_t n, float* out);

/* the same for prompts of DIFFERENT lengths: prompts [batch][max_len] (row b: lens[b] tokens, the rest ignored).  Sequences are
 * left-aligned and advance in lockstep; one whose prompt has ended feeds its own picks while longer prompts are still being
 * read, so every sequence produces exactly what generate() would for it alone. */
int ti_b200_generate_batch_ragged(ti_model_t m, const int32_t* prompts, const int32_t* lens, int32_t batch, int32_t max_len,
                                  int32_t n_new, int32_t* out_tokens, float* decode_ms);

/* InferenceEngine::generate_beam_search (src/model/inference_engine.cpp:830-871) / beam_search_decode (:1912-2069) on the cached
 * engine: beams are rows of the lockstep step, a beam's KV cache is a page table into one shared pool (a fork copies the table and
 * bumps page reference counts; the prompt is cached once), the expansion (softmax(logits / temperature), top-k and top-p on the
 * probabilities, the beam_size most probable tokens) runs on the device.  Bookkeeping as the reference's: cumulative
 * log-probability, score = log_prob / (prompt + new tokens)^length_penalty, keep the beam_size best, a candidate finishes on
 * eos_token or at max_new tokens, stop when beam_size candidates have finished.  Results best score first:
 * out_tokens [beam_size][max_new] (new tokens only), out_lens / out_logprob / out_score / out_finished [beam_size] (the last three
 * optional), *n_results <= beam_size.  beam_size <= 64.  Ties (which std::sort leaves unspecified in the reference) go to the
 * earlier candidate / the lower token id. */
int ti_b200_beam_search(ti_model_t m, const int32_t* prompt, int32_t n_prompt, int32_t max_new, int32_t beam_size, float temperature,
                        int32_t top_k, float top_p, float length_penalty, int32_t eos_token, int32_t* out_tokens, int32_t* out_lens,
                        float* out_logprob, float* out_score, int32_t* out_finished, int32_t* n_results);
/* the expansion step alone, on host logits [rows][vocab]: probs / tokens [rows][beam_size], counts [rows] */
int ti_b200_beam_expand(const float* logits_host, size_t rows, size_t vocab, float temperature, int32_t top_k, float top_p, int32_t beam_size,
                        float* probs_out, int32_t* tokens_out, int32_t* counts_out);

/* CUDA-event time of the prompt phase (prefill) of the last ti_b200_generate_greedy call on this model, in ms */
int ti_b200_model_last_prefill_ms(ti_model_t m, float* ms);

/* ---- instrumentation for bench.py ------------------------------------------------------------------ */
/* number of kernels this library has launched since init (the bench's `gpu_launches`) */
int ti_b200_launch_count(uint64_t* n);
/* times `reps` back-to-back launches of the GEMV over `n_w` different packed weights (cycled, so the working
 * set exceeds L2 when their total size does) with CUDA events on the library stream; ms = total elapsed */
int ti_b200_bench_gemv(const ti_qweight_t* w, size_t n_w, size_t reps, float* ms);
/* same, over the model's own packed matrices of one kind, cycling through the layers:
 * slot 0 = fused q|k|v, 1 = o_proj, 2 = fused gate/up (or up), 3 = down_proj, 4 = lm_head.
 * alg_bytes_per_launch = K*N*bits/8 + 4*N (scales) + 4*(K + N) (x in, y out), the figure of SURVEY.md 8d. */
int ti_b200_model_bench_gemv(ti_model_t m, int slot, size_t reps, float* ms, double* alg_bytes_per_launch);
/* times `reps` launches of the tensor-core GEMM kernel alone (activations already converted and resident);
 * ms = average per launch, ops_per_launch = 2 * 3 * rows * K * N integer operations (three INT8 digit planes) */
int ti_b200_bench_gemm(ti_qweight_t w, size_t rows, size_t reps, float* ms, double* ops_per_launch);

/* debug: one decode step on the persistent-kernel engine with CTA 0 recording 6 SM-clock stamps per phase
 * (phase start, barrier passed, x staged, weights consumed, epilogue done, barrier arrived) */
int ti_b200_debug_timeline(ti_model_t m, int32_t token, int64_t* stamps, size_t cap, size_t* n_phases);
/* the same with EVERY CTA recording: stamps[cta][phase][32]; slots 25 / 26 = SM clock when the CTA finished the phase / when it
 * passed the grid barrier, 27 / 28 = the same on the global nanosecond timer (comparable across CTAs) */
int ti_b200_debug_timeline_all(ti_model_t m, int32_t token, int64_t* stamps, size_t cap, size_t* n_phases, size_t* n_ctas);

#ifdef __cplusplus
}
#endif
#endif
