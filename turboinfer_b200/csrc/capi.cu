// capi.cu -- implementation of include/ti_b200.h: handle tables, device memory, kernel launches, and the
// device-resident decoder (weights quantized + packed once, paged KV cache, one CUDA graph per decode step).
// Host-side only orchestration lives here; all arithmetic of the hot path runs in the sm_100a kernels of
// gemv.cuh / kernels.cuh.  There is no CPU fallback.
#include "../../include/ti_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "mega.cuh"
#include "gemm_tc.cuh"
#include "prefill.cuh"
#include "batch.cuh"
#include "sampling.cuh"

using namespace tib;

namespace {

thread_local std::string g_err;
int g_device = -1;
int g_num_sms = 0;
cudaStream_t g_stream = nullptr;
uint64_t g_launches = 0;
bool g_use_pdl = false;
static int kEosChunk = 32;   // decode steps per launch when the generation may end early (stop_on_eos); TURBOINFER_B200_EOS_CHUNK for experiments
bool g_attr_done = false;
bool g_batch_carveout_done = false;   // the batched step's uniform carve-out preference (reset by shutdown, like g_attr_done)

// ---- tensor parallelism: NCCL is loaded at run time (dlopen), so the library has no link-time dependency on it ----
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;
ncclComm_t g_comm = nullptr;
int g_tp_size = 1, g_tp_rank = 0;

int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr)            \
    do {                     \
        int rc_ = (expr);    \
        if (rc_ != 0) return rc_; \
    } while (0)

int need_init() {
    if (g_device < 0) return fail("ti_b200: not initialised -- call ti_b200_init(device) first (no CPU fallback exists)");
    return 0;
}

int grid_for(size_t n, int block = 256) {
    size_t g = (n + block - 1) / block;
    size_t cap = (size_t)g_num_sms * 8;
    return (int)std::max<size_t>(1, std::min(g, cap));
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) return 0;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
        n = count;
        return 0;
    }
};

// ---- packed weight ------------------------------------------------------------------------------------
struct QWeight {
    QLayout L{};
    int qtype = TI_Q_INT4;
    int stages = 0;
    size_t smem = 0;
    float scale = 0.f, zp = 0.f;  // of the first (or only) source tensor
    int offset4 = 8;              // what was added to INT4 values when stored
    bool has_zterm = false;
    DevBuf<uint8_t> packed;
    DevBuf<float> colscale, colzterm;
    // byte copy for the tensor-core GEMMs (built on first use): tile-major, pre-swizzled (gemm_tc.cuh wtile_offset)
    DevBuf<uint8_t> kmajor;
    int k_pad = 0, n_pad = 0;
    size_t bytes() const { return layout_bytes(L); }
};

int pick_stages(const QLayout& L, int* stages, size_t* smem) {
    for (int s = kMaxStages; s >= 2; --s) {
        size_t b = gemv_smem_bytes(L, s);
        if (b <= 227 * 1024) {
            *stages = s;
            *smem = b;
            return 0;
        }
    }
    return fail("GEMV K=%d does not fit shared memory (x digits need %d KiB)", L.K, layout_kpad(L) * 3 / 1024);
}

int set_kernel_attrs() {
    // every cudaFuncSetAttribute of the library in one place, behind one flag that ti_b200_shutdown resets: attributes
    // belong to the device context, so a fresh init (same or another device) must set them again
    if (g_attr_done) return 0;
    CK(cudaFuncSetAttribute(gemv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(gemv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(attn_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(mega_decode_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(mega_decode_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(mega_decode_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(gemm_i8_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmemBytes));
    CK(cudaFuncSetAttribute(gemm_i8_tc_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmemBytes));
    CK(cudaFuncSetAttribute(rmsnorm_digits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(causal_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPfSmemBytes));
    CK(cudaFuncSetAttribute(causal_attention_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_tc_smem_bytes<128>()));
    CK(cudaFuncSetAttribute(causal_attention_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_tc_smem_bytes<64>()));
    CK(cudaFuncSetAttribute(causal_attention_h3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_h3_smem_bytes<128>()));
    CK(cudaFuncSetAttribute(causal_attention_h3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_h3_smem_bytes<64>()));
    CK(cudaFuncSetAttribute(sample_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024));
    CK(cudaFuncSetAttribute(sample_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024));
    CK(cudaFuncSetAttribute(beam_expand_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024));
    CK(cudaFuncSetAttribute(beam_expand_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024));
    g_attr_done = true;
    return 0;
}

void fill_weight(const QWeight& w, GemvArgs& a) {
    a.wq = w.packed.p;
    a.colscale = w.colscale.p;
    a.colzterm = w.has_zterm ? w.colzterm.p : nullptr;
    a.L = w.L;
    a.stages = w.stages;
    a.woff = w.L.bits == 4 ? w.offset4 : 0;
}

int launch_gemv(const QWeight& w, GemvArgs a, cudaStream_t st) {
    fill_weight(w, a);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(w.L.P);
    cfg.blockDim = dim3(kGemvThreads);
    cfg.dynamicSmemBytes = w.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    if (w.L.bits == 4) CK(cudaLaunchKernelEx(&cfg, gemv_kernel<4>, a));
    else CK(cudaLaunchKernelEx(&cfg, gemv_kernel<8>, a));
    ++g_launches;
    return 0;
}

// source matrix on the device awaiting packing: fp32 (quantized while packing), or the int8 / int32 integers of an already
// quantized tensor with the (scale, zero_point) they were made with (ti_b200_model_set_tensor_q, the .tinq path)
struct RawTensor {
    DevBuf<float> data;
    DevBuf<uint8_t> qdata;
    int kind = 0;   // 0 fp32, 1 int8, 2 int32
    float scale = 0.f, zp = 0.f;
    size_t rows = 0, cols = 0;
    bool present() const { return data.p != nullptr || qdata.p != nullptr; }
    void release() { data.release(); qdata.release(); }
};

// min/max -> (scale, zero_point) on the device; sz = 2 floats
int device_quant_params(const float* x, size_t n, int qtype, int symmetric, float* sz_dev, uint32_t* mm_dev, cudaStream_t st) {
    const uint32_t init[2] = {0xFFFFFFFFu, 0u};
    CK(cudaMemcpyAsync(mm_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
    minmax_kernel<<<grid_for(n), 256, 0, st>>>(x, n, mm_dev);
    quant_params_kernel<<<1, 1, 0, st>>>(mm_dev, qtype, symmetric, sz_dev);
    g_launches += 2;
    CK(cudaGetLastError());
    return 0;
}

// A source of a packed weight: a view (rows [row0, row0 + K), columns [col0, col0 + n)) of a whole fp32 tensor
// [rows][cols] on the device.  The quantization parameters always come from the WHOLE tensor.
struct SrcView {
    const void* full;
    size_t rows, cols;
    size_t row0, col0;
    int n;
    int kind = 0;                  // as RawTensor::kind
    float scale = 0.f, zp = 0.f;   // kind != 0: the given parameters
};

int build_qweight_views(const SrcView* v, int nsrc, int mode, int K, int qtype, int symmetric, bool unit_scale, std::unique_ptr<QWeight>* out);

// Quantize + pack up to three device fp32 sources (all [K][n_i]) into one streaming weight.
int build_qweight(const float* const* src, const int* src_n, int nsrc, int mode, int K, int qtype, int symmetric, bool unit_scale,
                  std::unique_ptr<QWeight>* out) {
    SrcView v[3];
    for (int i = 0; i < nsrc; ++i) v[i] = SrcView{src[i], (size_t)K, (size_t)src_n[i], 0, 0, src_n[i], 0, 0.f, 0.f};
    return build_qweight_views(v, nsrc, mode, K, qtype, symmetric, unit_scale, out);
}

int build_qweight_views(const SrcView* v, int nsrc, int mode, int K, int qtype, int symmetric, bool unit_scale, std::unique_ptr<QWeight>* out) {
    const void* src[3];
    int src_n[3];
    for (int i = 0; i < nsrc; ++i) {
        const size_t esz = v[i].kind == 1 ? 1 : 4;
        src[i] = static_cast<const char*>(v[i].full) + (v[i].row0 * v[i].cols + v[i].col0) * esz;
        src_n[i] = v[i].n;
        if (v[i].kind != 0 && v[i].kind != (qtype == TI_Q_INT8 ? 1 : 2)) return fail("quantized tensor type does not match the weight's quantization type");
    }
    auto w = std::make_unique<QWeight>();
    int N = 0;
    for (int i = 0; i < nsrc; ++i) N += src_n[i];
    w->L = make_layout(K, N, qtype == TI_Q_INT4 ? 4 : 8, g_num_sms);
    w->qtype = qtype;
    TRY(pick_stages(w->L, &w->stages, &w->smem));
    TRY(w->packed.alloc(layout_bytes(w->L)));
    TRY(w->colscale.alloc((size_t)4 * w->L.U));
    TRY(w->colzterm.alloc((size_t)4 * w->L.U));
    DevBuf<float> sz;
    DevBuf<uint32_t> mm;
    TRY(sz.alloc(2 * 3));
    TRY(mm.alloc(2));
    PackArgs pa{};
    pa.nsrc = nsrc;
    pa.mode = mode;
    pa.qtype = qtype;
    pa.off4 = symmetric ? 8 : 0;
    pa.unit_scale = unit_scale ? 1 : 0;
    pa.L = w->L;
    pa.out = w->packed.p;
    pa.colscale = w->colscale.p;
    pa.colzterm = w->colzterm.p;
    for (int i = 0; i < nsrc; ++i) {
        if (v[i].kind == 0) {
            TRY(device_quant_params(static_cast<const float*>(v[i].full), v[i].rows * v[i].cols, qtype, symmetric, sz.p + 2 * i, mm.p, g_stream));
        } else {
            const float given[2] = {v[i].scale, v[i].zp};
            CK(cudaMemcpyAsync(sz.p + 2 * i, given, sizeof(given), cudaMemcpyHostToDevice, g_stream));
            // INT4 integers made with a zero-point live in [0, 15] (quantize_to_int4, quantization.cpp:683-693); under the signed
            // nibble code they are stored as q - 8
            pa.src[i].shift4 = (qtype == TI_Q_INT4 && pa.off4 == 8 && v[i].zp != 0.0f) ? 8 : 0;
        }
        pa.src[i].kind = v[i].kind;
        pa.src[i].w = src[i];
        pa.src[i].n = src_n[i];
        pa.src[i].ld = (int)v[i].cols;
        pa.src[i].sz = sz.p + 2 * i;
    }
    pack_kernel<<<w->L.P, kConsumerThreads, 0, g_stream>>>(pa);
    ++g_launches;
    CK(cudaGetLastError());
    float h[6] = {0};
    CK(cudaMemcpyAsync(h, sz.p, sizeof(float) * 2 * nsrc, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    w->scale = h[0];
    w->zp = h[1];
    w->has_zterm = false;
    for (int i = 0; i < nsrc; ++i) w->has_zterm |= (h[2 * i + 1] != 0.0f);   // (a shifted source has a zero-point by definition)
    w->offset4 = symmetric ? 8 : 0;
    *out = std::move(w);
    return 0;
}

// ---- model ---------------------------------------------------------------------------------------------
struct Layer {
    RawTensor raw_q, raw_k, raw_v, raw_o, raw_up, raw_gate, raw_down;  // pending fp32 sources
    std::unique_ptr<QWeight> qkv, o, gateup, down;
    DevBuf<float> attn_norm, ffn_norm;
    DevBuf<float> k_pool, v_pool;
    bool has_gate = false;
    // compat_literal: reference-layout fp32 (or integer-valued fp32) matrices
    DevBuf<float> lit_up, lit_down, lit_gate;
};

// Batched decode: B sequences in lockstep, each with its own pages of every layer's KV pools.
struct BatchState {
    int B = 0, pages_per_seq = 0, max_splits = 1;
    std::vector<DevBuf<float>> k, v;     // per layer: [B * pages_per_seq][page_tokens][H]
    DevBuf<int> tables;                  // [B][pages_per_seq] physical page of each logical page
    DevBuf<int> tokens;                  // [B] tokens of the current step
    DevBuf<int> prompts;                 // [max prompt length][B] (row p = the tokens of prompt position p, padded)
    DevBuf<int> lens;                    // [B] prompt length of every sequence
    DevBuf<int> out;                     // [B][n_new]
    DevBuf<int> pos_step;                // [0] tokens in every cache, [1] output column
    DevBuf<float> part_o, part_ml, logits;
    DevBuf<float> ar;                    // tensor parallel: [B][H] partial output of a row-parallel GEMM, all-reduced in place
    cudaGraphExec_t graph[2] = {nullptr, nullptr};   // [0] step without sampling, [1] with lm_head + argmax
    unsigned int cap_scratch = 0xFFFFFFFFu;          // what the captured graphs were built against: the scratch generation
    int cap_stride = 0;                              // ... and the row stride of `out`
    void drop_graphs() { for (auto& g : graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; } }
    ~BatchState() { drop_graphs(); }
};

// beam search (generate_beam_search, :830-871 / :1912-2069) on the lockstep engine: a beam is a row of the batch, its KV cache a
// page table into one shared pool -- forking a beam copies the table (reference counts per page), not the cache
struct BeamState {
    BatchState bs;                       // rows = beams; bs.k / bs.v hold `pool_pages` pages per layer
    int beam = 0, pool_pages = 0;
    DevBuf<float> cand_prob;             // [beam][beam] expansion of every row
    DevBuf<int> cand_tok, cand_cnt;
    DevBuf<float*> k_ptrs, v_ptrs;       // per-layer pool addresses for the page copy
    DevBuf<int> copy_pairs;              // [beam][2]
    cudaGraphExec_t graph[2] = {nullptr, nullptr};   // forward without / with lm_head + expansion
    unsigned int cap_scratch = 0xFFFFFFFFu;
    float cap_temperature = 0.f, cap_top_p = 0.f;
    int cap_top_k = -1;
    void drop_graphs() { for (auto& g : graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; } }
    ~BeamState() { drop_graphs(); }
};

struct Model {
    ti_model_config cfg{};
    std::vector<Layer> layers;
    DevBuf<float> tok_emb, out_norm;
    RawTensor raw_lm;
    std::unique_ptr<QWeight> lm_head;
    std::unique_ptr<QWeight> lm_head_shard;   // tensor parallel: this rank's V / tp columns (fused engine; the full copy serves the other engines)
    bool lm_sharded = false;
    size_t xchg_key_off = 0;
    DevBuf<float> lit_lm;
    DevBuf<float> lit_x, lit_pa, lit_n, lit_u, lit_g, lit_f;   // compat_literal activations, sized for lit_rows rows
    int lit_rows = 0;
    bool finalized = false;
    int page_tokens = 64, num_pages = 0;
    DevBuf<int> page_table;
    DevBuf<StepState> state;
    DevBuf<float> x, nrm, q, attn_out, act, logits, part_o, part_ml, inv_freq, hist;
    DevBuf<int> out_tokens, prompt;
    DevBuf<int> smp_tokens, smp_count;      // sampled generation: token history [n_new], the step counter of the RNG
    DevBuf<float> smp_logprobs;             // ... and the log-probability of every pick
    int attn_heads = 1, attn_dim = 0, max_splits = 1;
    size_t attn_smem = 0;
    cudaGraphExec_t graph_decode = nullptr, graph_prefill = nullptr;
    DevBuf<StepIO> io;
    int launches_decode = 0, launches_prefill = 0;  // kernels per captured step
    // persistent-kernel engine
    bool use_mega = false;
    // batched prefill (tensor-core GEMM path): scratch sized for pf_cap prompt rows
    int pf_cap = 0;
    float last_prefill_ms = 0.f;   // CUDA-event time of the prompt phase of the last generate call
    DevBuf<float> pf_x, pf_qkv, pf_attn, pf_gu, pf_act, pf_sx;
    DevBuf<int8_t> pf_planes;
    DevBuf<long long> pf_sxf;
    DevBuf<unsigned long long> pf_ws;   // split-K workspace of the 32-row GEMM: [n_pad][32] integer partial sums, zero between GEMMs
    DevBuf<unsigned int> pf_cnt;        // ... and its per-tile arrival counters
    unsigned int pf_gen = 0;            // bumped whenever the scratch buffers are (re)allocated: captured graphs hold their pointers
    bool pf_small_layout = false;       // pf_planes currently holds the 32-row GEMM's tile images (batch.cuh) rather than [3][m_pad][k_pad]
    DevBuf<int> pf_tokens;
    std::unique_ptr<struct BatchState> batch;   // batched decode (generate_batch): per-sequence KV pages, step graphs
    std::unique_ptr<struct BeamState> beams;    // beam search: rows of the lockstep engine over a shared, reference-counted page pool
    int tp = 1, tp_rank = 0;   // tensor-parallel degree / rank of this model
    DevBuf<float> ar_tmp;      // [H] partial output of a row-parallel GEMV, all-reduced in place
    // fused tensor-parallel engine: this rank's exchange block {barrier words, partial buffers 0 / 1 [tp][H]} and the
    // CUDA IPC mappings of every rank's block (own rank: the block itself)
    bool tp_fused = false;
    void* xchg = nullptr;
    void* peer_base[kMaxTp] = {};
    size_t xchg_flag_off = 0, xchg_part_off = 0, xchg_part_bytes = 0;
    bool tp_p2p = false;   // point-to-point exchange per column slice instead of the barrier across the GPUs + reduce phase
    bool tp_ll = false;    // ... with the partials as tagged 8-byte words polled by the reader (one NVLink crossing per exchange)
    DevBuf<MegaPhase> phases;
    DevBuf<ProdRec> prod;
    int nphases = 0;
    DevBuf<unsigned int> sync_buf;   // kBarWords barrier words (128 B apart), then two 64-bit argmax keys
    DevBuf<unsigned int> head_cnt;
    int mega_stages = 0, mega_max_kpad = 0, mega_max_units = 0, mega_attn_floats = 0;
    DevBuf<long long> dbg;
    bool dbg_on = false, dbg_all = false;
    DevBuf<XStats> emb_stats;
    size_t mega_smem = 0;
    int host_pos = 0;  // mirror of state.pos
    ~Model() {
        for (int r = 0; r < kMaxTp; ++r)
            if (peer_base[r] && peer_base[r] != xchg) cudaIpcCloseMemHandle(peer_base[r]);
        if (xchg) cudaFree(xchg);
        if (graph_decode) cudaGraphExecDestroy(graph_decode);
        if (graph_prefill) cudaGraphExecDestroy(graph_prefill);
    }
};

std::mutex g_mu;
std::vector<std::unique_ptr<QWeight>> g_qweights;
std::vector<std::unique_ptr<Model>> g_models;

// stand-alone paged KV cache (ti_b200_kv_*): the reference's KVCache (src/model/inference_engine.cpp:25-172) as an object of
// its own -- per layer two pools [page][page_tokens][heads * head_dim] fp32 and one page table; append is a scatter, the
// attention reads the pages in place (no copy-out), reset() forgets the lengths
struct KvCache {
    int layers = 0, heads = 0, head_dim = 0, max_seq = 0, page_tokens = 64, num_pages = 0;
    std::vector<DevBuf<float>> k, v;
    std::vector<int> len;            // tokens cached per layer
    DevBuf<int> page_table, pos;
    DevBuf<float> part_o, part_ml;
    int max_splits = 1;
};
std::vector<std::unique_ptr<KvCache>> g_kvs;
KvCache* get_kv(uint64_t h) {
    if (h == 0 || h > g_kvs.size()) return nullptr;
    return g_kvs[h - 1].get();
}

QWeight* get_qw(ti_qweight_t h) {
    if (h == 0 || h > g_qweights.size()) return nullptr;
    return g_qweights[h - 1].get();
}
Model* get_model(ti_model_t h) {
    if (h == 0 || h > g_models.size()) return nullptr;
    return g_models[h - 1].get();
}

int upload(DevBuf<float>& dst, const float* host, size_t n) {
    TRY(dst.alloc(n));
    CK(cudaMemcpyAsync(dst.p, host, n * sizeof(float), cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

// "layers.3.attention.q_proj.weight" -> (3, slot).  Accepts the naming variants of initialize_model (:483-563).
enum Slot { S_NONE, S_EMB, S_NORM, S_LM, S_Q, S_K, S_V, S_O, S_AN, S_FN, S_UP, S_DOWN, S_GATE };
Slot parse_name(const std::string& name, int* layer) {
    *layer = -1;
    if (name == "token_embeddings.weight" || name == "embed_tokens.weight" || name == "model.embed_tokens.weight") return S_EMB;
    if (name == "norm.weight" || name == "model.norm.weight") return S_NORM;
    if (name == "lm_head.weight" || name == "output.weight") return S_LM;
    std::string rest = name;
    if (rest.rfind("model.", 0) == 0) rest = rest.substr(6);
    if (rest.rfind("layers.", 0) != 0) return S_NONE;
    rest = rest.substr(7);
    size_t dot = rest.find('.');
    if (dot == std::string::npos) return S_NONE;
    *layer = atoi(rest.substr(0, dot).c_str());
    rest = rest.substr(dot + 1);
    auto is = [&](const char* a, const char* b) { return rest == a || rest == b; };
    if (is("self_attn.q_proj.weight", "attention.q_proj.weight")) return S_Q;
    if (is("self_attn.k_proj.weight", "attention.k_proj.weight")) return S_K;
    if (is("self_attn.v_proj.weight", "attention.v_proj.weight")) return S_V;
    if (is("self_attn.o_proj.weight", "attention.o_proj.weight")) return S_O;
    if (is("input_layernorm.weight", "attention_norm.weight")) return S_AN;
    if (is("post_attention_layernorm.weight", "ffn_norm.weight")) return S_FN;
    if (is("mlp.up_proj.weight", "feed_forward.w1.weight")) return S_UP;
    if (is("mlp.down_proj.weight", "feed_forward.w2.weight")) return S_DOWN;
    if (is("mlp.gate_proj.weight", "feed_forward.w3.weight")) return S_GATE;
    return S_NONE;
}

enum ShardKind { SH_NONE, SH_COLS, SH_ROWS };   // column-parallel (split N) / row-parallel (split K)
SrcView shard_view(const RawTensor& raw, ShardKind kind, int tp, int rank) {
    SrcView v{raw.kind ? (const void*)raw.qdata.p : (const void*)raw.data.p, raw.rows, raw.cols, 0, 0, (int)raw.cols, raw.kind, raw.scale, raw.zp};
    if (tp > 1 && kind == SH_COLS) { v.n = (int)(raw.cols / tp); v.col0 = (size_t)rank * v.n; }
    if (tp > 1 && kind == SH_ROWS) v.row0 = (size_t)rank * (raw.rows / tp);
    return v;
}
int pack_single(RawTensor& raw, int qtype, std::unique_ptr<QWeight>* out, ShardKind kind = SH_NONE, int tp = 1, int rank = 0) {
    const SrcView v = shard_view(raw, kind, tp, rank);
    const int K = (tp > 1 && kind == SH_ROWS) ? (int)(raw.rows / tp) : (int)raw.rows;
    TRY(build_qweight_views(&v, 1, 0, K, qtype, 1, false, out));
    raw.release();
    return 0;
}

// pack whatever groups of a layer have become complete (keeps peak fp32 residency to one group)
int pack_ready(Model& m, Layer& ly, bool final_pass) {
    const int qt = m.cfg.qtype;
    if (m.cfg.compat_literal) return 0;
    if (!ly.qkv && ly.raw_q.present() && ly.raw_k.present() && ly.raw_v.present()) {
        // tensor parallel: each rank keeps the columns of its own heads of q, k and v (column-parallel)
        const SrcView v[3] = {shard_view(ly.raw_q, SH_COLS, m.tp, m.tp_rank), shard_view(ly.raw_k, SH_COLS, m.tp, m.tp_rank),
                              shard_view(ly.raw_v, SH_COLS, m.tp, m.tp_rank)};
        TRY(build_qweight_views(v, 3, 0, (int)ly.raw_q.rows, qt, 1, false, &ly.qkv));
        ly.raw_q.release(); ly.raw_k.release(); ly.raw_v.release();
    }
    if (!ly.o && ly.raw_o.present()) TRY(pack_single(ly.raw_o, qt, &ly.o, SH_ROWS, m.tp, m.tp_rank));          // row-parallel
    if (!ly.down && ly.raw_down.present()) TRY(pack_single(ly.raw_down, qt, &ly.down, SH_ROWS, m.tp, m.tp_rank));
    if (!ly.gateup && ly.raw_up.present() && ly.raw_gate.present()) {
        const SrcView v[2] = {shard_view(ly.raw_gate, SH_COLS, m.tp, m.tp_rank), shard_view(ly.raw_up, SH_COLS, m.tp, m.tp_rank)};  // even columns = gate, odd = up
        TRY(build_qweight_views(v, 2, 1, (int)ly.raw_up.rows, qt, 1, false, &ly.gateup));
        ly.has_gate = true;
        ly.raw_up.release(); ly.raw_gate.release();
    }
    if (final_pass && !ly.gateup && ly.raw_up.present()) {  // no gate: relu(up) (:392-395)
        TRY(pack_single(ly.raw_up, qt, &ly.gateup, SH_COLS, m.tp, m.tp_rank));
        ly.has_gate = false;
    }
    return 0;
}

int store_tensor(Model& m, const std::string& name, DevBuf<float>&& dev, size_t rows, size_t cols, RawTensor* quantized = nullptr) {
    int li = -1;
    Slot s = parse_name(name, &li);
    if (s == S_NONE) return fail("unknown tensor name '%s'", name.c_str());
    if (quantized && (s == S_EMB || s == S_NORM || s == S_AN || s == S_FN))
        return fail("tensor '%s' is read as float32 (embedding rows and norm weights are not quantized on this path)", name.c_str());
    if (quantized && m.cfg.compat_literal) return fail("the literal benchmark path takes float32 tensors");
    const size_t H = m.cfg.hidden, V = m.cfg.vocab, I = m.cfg.inter;
    auto expect = [&](size_t r, size_t c) -> int {
        if (rows != r || cols != c) return fail("tensor '%s' has shape [%zu,%zu], expected [%zu,%zu]", name.c_str(), rows, cols, r, c);
        return 0;
    };
    auto take = [&](RawTensor& t) {
        t.release();
        if (quantized) {
            t.qdata = std::move(quantized->qdata);
            t.kind = quantized->kind; t.scale = quantized->scale; t.zp = quantized->zp;
        } else {
            t.data.p = dev.p; t.data.n = dev.n; dev.p = nullptr; dev.n = 0;
            t.kind = 0;
        }
        t.rows = rows; t.cols = cols;
    };
    auto take_vec = [&](DevBuf<float>& t) {
        t.release();
        t.p = dev.p; t.n = dev.n; dev.p = nullptr; dev.n = 0;
    };
    if (s == S_EMB) { TRY(expect(V, H)); take_vec(m.tok_emb); return 0; }
    if (s == S_NORM) { if (rows * cols != H) return fail("norm.weight must have %zu elements", H); take_vec(m.out_norm); return 0; }
    if (s == S_LM) {
        TRY(expect(H, V));
        take(m.raw_lm);
        if (!m.cfg.compat_literal) {
            if (m.tp > 1 && V % m.tp == 0) {   // column-parallel slice of the vocabulary, same whole-tensor quantization parameters
                const SrcView v = shard_view(m.raw_lm, SH_COLS, m.tp, m.tp_rank);
                TRY(build_qweight_views(&v, 1, 0, (int)H, m.cfg.qtype, 1, false, &m.lm_head_shard));
            }
            TRY(pack_single(m.raw_lm, m.cfg.qtype, &m.lm_head));
        }
        return 0;
    }
    if (li < 0 || li >= (int)m.layers.size()) return fail("layer index out of range in '%s'", name.c_str());
    Layer& ly = m.layers[li];
    switch (s) {
        case S_Q: TRY(expect(H, H)); take(ly.raw_q); break;
        case S_K: TRY(expect(H, H)); take(ly.raw_k); break;
        case S_V: TRY(expect(H, H)); take(ly.raw_v); break;
        case S_O: TRY(expect(H, H)); take(ly.raw_o); break;
        case S_UP: TRY(expect(H, I)); take(ly.raw_up); break;
        case S_GATE: TRY(expect(H, I)); take(ly.raw_gate); break;
        case S_DOWN: TRY(expect(I, H)); take(ly.raw_down); break;
        case S_AN: if (rows * cols != H) return fail("bad norm size"); take_vec(ly.attn_norm); break;
        case S_FN: if (rows * cols != H) return fail("bad norm size"); take_vec(ly.ffn_norm); break;
        default: break;
    }
    return pack_ready(m, ly, false);
}

// ---- decode step -------------------------------------------------------------------------------------
int enqueue_attention(Model& m, Layer& ly, cudaStream_t st) {
    AttnArgs a{};
    a.q = m.q.p;
    a.k_pool = ly.k_pool.p;
    a.v_pool = ly.v_pool.p;
    a.page_table = m.page_table.p;
    a.page_tokens = m.page_tokens;
    a.page_shift = log2_if_pow2(m.page_tokens);
    a.pos_ptr = &m.state.p->pos;
    a.t_bias = 1;
    a.H = m.cfg.hidden / m.tp;
    a.D = m.attn_dim;
    a.heads = m.attn_heads;
    a.max_splits = m.max_splits;
    a.min_chunk = 64;
    a.scale = 1.0f / sqrtf((float)m.attn_dim);  // :1288 (D = hidden in literal single-head mode, head_dim otherwise)
    a.part_o = m.part_o.p;
    a.part_ml = m.part_ml.p;
    a.out = m.attn_out.p;
    attn_partial_kernel<<<dim3(m.attn_heads, m.max_splits), kAttnThreads, m.attn_smem, st>>>(a);
    attn_combine_kernel<<<m.attn_heads, 256, 0, st>>>(a);
    g_launches += 2;
    CK(cudaGetLastError());
    return 0;
}

// row-parallel GEMV output under tensor parallelism: partial [H] -> NCCL all-reduce (sum) -> x += partial
int tp_allreduce_add(Model& m, cudaStream_t st) {
    const int H = m.cfg.hidden;
    const ncclResult_t r = g_nccl.AllReduce(m.ar_tmp.p, m.ar_tmp.p, (size_t)H, ncclFloat32, ncclSum, g_comm, st);
    if (r != ncclSuccess) return fail("ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
    elementwise_kernel<<<grid_for(H), 256, 0, st>>>(m.x.p, m.ar_tmp.p, m.x.p, H, EW_ADD);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

int enqueue_step(Model& m, bool with_head, cudaStream_t st) {
    const int H = m.cfg.hidden;
    const int Hl = H / m.tp;   // this rank's share of the attention width (its heads)
    StepState* S = m.state.p;
    embed_kernel<<<grid_for(H), 256, 0, st>>>(m.tok_emb.p, S, m.x.p, H, 0);
    ++g_launches;
    for (auto& ly : m.layers) {
        const bool attn = ly.qkv && ly.o;
        if (attn) {
            GemvArgs a{};
            a.x = m.x.p;
            a.norm_w = ly.attn_norm.p;
            a.rms_eps = m.cfg.rms_eps;
            a.epi = EPI_QKV;
            a.out = m.q.p;
            a.hidden = Hl;
            a.rope_dim = m.cfg.rope_mode == 1 ? H / m.cfg.heads : (m.cfg.rope_mode == 2 ? H : 0);
            a.inv_freq = m.inv_freq.p;
            a.pos_ptr = &S->pos;
            a.k_pool = ly.k_pool.p;
            a.v_pool = ly.v_pool.p;
            a.page_table = m.page_table.p;
            a.page_tokens = m.page_tokens;
            TRY(launch_gemv(*ly.qkv, a, st));
            TRY(enqueue_attention(m, ly, st));
            GemvArgs o{};
            o.x = m.attn_out.p;
            if (m.tp > 1) {
                o.epi = EPI_STORE;
                o.out = m.ar_tmp.p;
                TRY(launch_gemv(*ly.o, o, st));
                TRY(tp_allreduce_add(m, st));
            } else {
                o.epi = EPI_RESIDUAL;
                o.resid = m.x.p;
                o.out = m.x.p;
                TRY(launch_gemv(*ly.o, o, st));
            }
        } else {
            // compute_attention returns its (normalised) input when a projection is missing (:293-296): x <- x + n
            const float* n = m.x.p;
            if (ly.attn_norm.p) {
                rms_norm_kernel<<<1, 256, 0, st>>>(m.x.p, ly.attn_norm.p, m.nrm.p, H, m.cfg.rms_eps);
                ++g_launches;
                n = m.nrm.p;
            }
            elementwise_kernel<<<grid_for(H), 256, 0, st>>>(m.x.p, n, m.x.p, H, EW_ADD);
            ++g_launches;
        }
        if (ly.gateup && ly.down) {
            GemvArgs g{};
            g.x = m.x.p;
            g.norm_w = ly.ffn_norm.p;
            g.rms_eps = m.cfg.rms_eps;
            g.epi = ly.has_gate ? EPI_SWIGLU : EPI_RELU;
            g.out = m.act.p;
            TRY(launch_gemv(*ly.gateup, g, st));
            GemvArgs d{};
            d.x = m.act.p;
            if (m.tp > 1) {
                d.epi = EPI_STORE;
                d.out = m.ar_tmp.p;
                TRY(launch_gemv(*ly.down, d, st));
                TRY(tp_allreduce_add(m, st));
            } else {
                d.epi = EPI_RESIDUAL;
                d.resid = m.x.p;
                d.out = m.x.p;
                TRY(launch_gemv(*ly.down, d, st));
            }
        } else {
            const float* f = m.x.p;  // compute_ffn returns its input (:377-380)
            if (ly.ffn_norm.p) {
                rms_norm_kernel<<<1, 256, 0, st>>>(m.x.p, ly.ffn_norm.p, m.nrm.p, H, m.cfg.rms_eps);
                ++g_launches;
                f = m.nrm.p;
            }
            elementwise_kernel<<<grid_for(H), 256, 0, st>>>(m.x.p, f, m.x.p, H, EW_ADD);
            ++g_launches;
        }
    }
    if (with_head) {
        GemvArgs l{};
        l.x = m.x.p;
        l.norm_w = m.out_norm.p;
        l.rms_eps = m.cfg.rms_eps;
        l.epi = EPI_LOGITS;
        l.out = m.logits.p;
        l.argmax_key = &S->argmax_key;
        TRY(launch_gemv(*m.lm_head, l, st));
    }
    return 0;
}

int capture_graph(Model& m, bool with_head, cudaGraphExec_t* out) {
    StepIO* io = m.io.p;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
    const uint64_t launches_before = g_launches;
    int rc = enqueue_step(m, with_head, g_stream);
    if (rc == 0) {
        step_finish_kernel<<<1, 1024, 0, g_stream>>>(m.state.p, io, m.logits.p, m.cfg.vocab, with_head ? 1 : 0);
        ++g_launches;
    }
    cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
    (with_head ? m.launches_decode : m.launches_prefill) = (int)(g_launches - launches_before);
    g_launches = launches_before;  // captured, not launched
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return 0;
}

// one forward pass on the per-op engine: replay the captured graph, or (tensor parallel) enqueue the kernels directly
int run_step(Model& m, bool with_head) {
    if (m.graph_decode != nullptr) {
        CK(cudaGraphLaunch(with_head ? m.graph_decode : m.graph_prefill, g_stream));
        g_launches += with_head ? m.launches_decode : m.launches_prefill;
        return 0;
    }
    TRY(enqueue_step(m, with_head, g_stream));
    step_finish_kernel<<<1, 1024, 0, g_stream>>>(m.state.p, m.io.p, m.logits.p, m.cfg.vocab, with_head ? 1 : 0);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

// ---- fused tensor parallelism: peer memory over NVLink ---------------------------------------------------------
// Every rank allocates one exchange block and maps the other ranks' blocks with CUDA IPC (the handles travel through an
// NCCL all-gather, the only collective left on this path).  Inside the persistent kernel a row-parallel GEMV stores its
// partial output straight into every rank's partial buffer and the ranks meet at a barrier of system-scope atomics; no
// NCCL call, no extra launch per layer.
int tp_exchange_setup(Model& m) {
    const int H = m.cfg.hidden, P = m.tp;
    if (P > kMaxTp) return fail("fused tensor parallelism supports up to %d ranks", kMaxTp);
    const size_t bar_bytes = (size_t)kBarWords * kBarStride * sizeof(unsigned int) + 256;   // + the barrier sequence number
    m.xchg_flag_off = (bar_bytes + 255) & ~size_t(255);                                     // flags [2 buffers][P][256 CTAs]
    m.xchg_key_off = m.xchg_flag_off + (size_t)2 * kMaxTp * 256 * sizeof(unsigned int);      // arg-max keys [2][kMaxTp]
    m.xchg_part_off = m.xchg_key_off + 256;
    m.xchg_part_bytes = ((size_t)P * H * sizeof(unsigned long long) + 255) & ~size_t(255);   // sized for the tagged 8-byte words of the LL exchange
    const size_t total = m.xchg_part_off + 2 * m.xchg_part_bytes;
    CK(cudaMalloc(&m.xchg, total));
    CK(cudaMemset(m.xchg, 0, total));
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, m.xchg));
    DevBuf<uint8_t> hsend, hrecv;
    TRY(hsend.alloc(sizeof(mine)));
    TRY(hrecv.alloc(sizeof(mine) * P));
    CK(cudaMemcpyAsync(hsend.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, g_stream));
    const ncclResult_t r = g_nccl.AllGather(hsend.p, hrecv.p, sizeof(mine), ncclUint8, g_comm, g_stream);
    if (r != ncclSuccess) return fail("ncclAllGather (IPC handles) failed: %s", g_nccl.GetErrorString(r));
    std::vector<cudaIpcMemHandle_t> all(P);
    CK(cudaMemcpyAsync(all.data(), hrecv.p, sizeof(mine) * P, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    for (int q = 0; q < P; ++q) {
        if (q == m.tp_rank) { m.peer_base[q] = m.xchg; continue; }
        const cudaError_t e = cudaIpcOpenMemHandle(&m.peer_base[q], all[q], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { m.peer_base[q] = nullptr; return fail("cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e)); }
    }
    // nobody may signal a block that its owner has not zeroed yet: one more collective as a barrier
    const ncclResult_t r2 = g_nccl.AllGather(hsend.p, hrecv.p, sizeof(mine), ncclUint8, g_comm, g_stream);
    if (r2 != ncclSuccess) return fail("ncclAllGather (barrier) failed: %s", g_nccl.GetErrorString(r2));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

// ---- persistent-kernel engine ---------------------------------------------------------------------------
int build_mega(Model& m) {
    const int H = m.cfg.hidden;
    std::vector<MegaPhase> ph;
    int max_kpad = 0, max_units = 0;
    auto gemv_phase = [&](const QWeight& w, GemvArgs g, int x_src, int resid_src, int is_head) {
        MegaPhase p{};
        p.type = PH_GEMV;
        p.x_src = x_src;
        p.resid_src = resid_src;
        p.is_head = is_head;
        fill_weight(w, g);
        p.g = g;
        max_kpad = std::max(max_kpad, layout_kpad(w.L));
        max_units = std::max(max_units, slab_max_units(w.L));
        ph.push_back(p);
    };
    const int mega_splits = std::max(1, std::min(g_num_sms / m.attn_heads, m.max_splits));
    bool lean_attn = m.attn_dim <= 128 && log2_if_pow2(m.page_tokens) >= 0;   // mega.cuh attn_item_lean + deferred merge
    if (const char* e = getenv("TURBOINFER_B200_LEAN_ATTN")) lean_attn = lean_attn && atoi(e) != 0;   // A/B experiments
    const bool tp = m.tp > 1;
    const int Hl = H / m.tp;   // this rank's share of the attention width (its heads)
    float* part[2] = {nullptr, nullptr};
    if (tp) {
        part[0] = reinterpret_cast<float*>(static_cast<uint8_t*>(m.xchg) + m.xchg_part_off);
        part[1] = reinterpret_cast<float*>(static_cast<uint8_t*>(m.xchg) + m.xchg_part_off + m.xchg_part_bytes);
    }
    // second half of the all-reduce after a row-parallel GEMV: x <- residual + the partials of all ranks
    auto reduce_phase = [&](int sel, int resid_src, const float* next_norm_w) {
        MegaPhase p{};
        p.type = PH_REDUCE;
        p.resid_src = resid_src;
        p.g.x = part[sel];
        p.g.resid = m.x.p;
        p.g.out = m.x.p;
        p.g.next_norm_w = next_norm_w;
        p.g.L.N = H;
        ph.push_back(p);
    };
    for (size_t l = 0; l < m.layers.size(); ++l) {
        Layer& ly = m.layers[l];
        const int src0 = l == 0 ? SRC_EMB : SRC_PTR;
        GemvArgs a{};
        a.x = m.x.p;
        a.norm_w = ly.attn_norm.p;
        a.rms_eps = m.cfg.rms_eps;
        a.epi = EPI_QKV;
        a.out = m.q.p;
        a.hidden = Hl;
        a.rope_dim = m.cfg.rope_mode == 1 ? H / m.cfg.heads : (m.cfg.rope_mode == 2 ? H : 0);
        a.inv_freq = m.inv_freq.p;
        a.pos_ptr = &m.state.p->pos;
        a.k_pool = ly.k_pool.p;
        a.v_pool = ly.v_pool.p;
        a.page_table = m.page_table.p;
        a.page_tokens = m.page_tokens;
        gemv_phase(*ly.qkv, a, src0, SRC_PTR, 0);
        MegaPhase at{};
        at.type = PH_ATTN;
        at.at.q = m.q.p;
        at.at.k_pool = ly.k_pool.p;
        at.at.v_pool = ly.v_pool.p;
        at.at.page_table = m.page_table.p;
        at.at.page_tokens = m.page_tokens;
        at.at.page_shift = lean_attn ? log2_if_pow2(m.page_tokens) : -1;   // -1: the in-phase merge of mega_attention
        at.at.pos_ptr = &m.state.p->pos;
        at.at.t_bias = 1;
        at.at.H = Hl;
        at.at.D = m.attn_dim;
        at.at.heads = m.attn_heads;
        at.at.max_splits = mega_splits;
        // lean item: 16 warps x 5 tokens in flight = one round trip per 80 tokens.  While a head is one item only `heads` of the 148 CTAs
        // work, so a second round trip already costs more than the split's atomic + merge (measured, 7B decode-256: 843 -> 851 tok/s
        // against a 320-token item; TinyLlama 1573 -> 1631)
        at.at.min_chunk = lean_attn ? 80 : 64;
        if (const char* e = getenv("TURBOINFER_B200_ATTN_CHUNK")) at.at.min_chunk = std::max(16, atoi(e));   // A/B experiments   // lean item: 16 warps x 5 tokens in flight = one round trip per 80 tokens; up to 4 round trips before a split pays
        at.at.scale = 1.0f / sqrtf((float)m.attn_dim);
        at.at.part_o = m.part_o.p;
        at.at.part_ml = m.part_ml.p;
        at.at.out = m.attn_out.p;
        ph.push_back(at);
        const int o_src = SRC_PTR;
        GemvArgs o{};
        o.x = m.attn_out.p;
        if (tp && m.tp_p2p) {   // row-parallel: partial output to every rank, the P CTAs of a column slice meet and reduce it
            o.epi = EPI_STORE;
            o.resid = m.x.p;
            o.out = m.x.p;
            o.next_norm_w = ly.ffn_norm.p;
            gemv_phase(*ly.o, o, o_src, src0, 0);
            ph.back().at = at.at;
            ph.back().mgpu = 2;
            ph.back().part_sel = 0;
        } else if (tp) {   // row-parallel: partial output to every rank, barrier across the GPUs, then the reduce phase
            o.epi = EPI_STORE;
            o.out = part[0];
            gemv_phase(*ly.o, o, o_src, SRC_PTR, 0);
            ph.back().at = at.at;
            ph.back().mgpu = 1;
            ph.back().part_sel = 0;
            reduce_phase(0, src0, ly.ffn_norm.p);
        } else {
            o.epi = EPI_RESIDUAL;
            o.resid = m.x.p;
            o.out = m.x.p;
            o.next_norm_w = ly.ffn_norm.p;
            gemv_phase(*ly.o, o, o_src, src0, 0);
            ph.back().at = at.at;
        }
        GemvArgs g{};
        g.x = m.x.p;
        g.norm_w = ly.ffn_norm.p;
        g.rms_eps = m.cfg.rms_eps;
        g.epi = ly.has_gate ? EPI_SWIGLU : EPI_RELU;
        g.out = m.act.p;
        gemv_phase(*ly.gateup, g, SRC_PTR, SRC_PTR, 0);
        GemvArgs d{};
        d.x = m.act.p;
        const float* norm_after = l + 1 < m.layers.size() ? m.layers[l + 1].attn_norm.p : m.out_norm.p;
        if (tp && m.tp_p2p) {
            d.epi = EPI_STORE;
            d.resid = m.x.p;
            d.out = m.x.p;
            d.next_norm_w = norm_after;
            gemv_phase(*ly.down, d, SRC_PTR, SRC_PTR, 0);
            ph.back().mgpu = 2;
            ph.back().part_sel = 1;
        } else if (tp) {
            d.epi = EPI_STORE;
            d.out = part[1];
            gemv_phase(*ly.down, d, SRC_PTR, SRC_PTR, 0);
            ph.back().mgpu = 1;
            ph.back().part_sel = 1;
            reduce_phase(1, SRC_PTR, norm_after);
        } else {
            d.epi = EPI_RESIDUAL;
            d.resid = m.x.p;
            d.out = m.x.p;
            d.next_norm_w = norm_after;
            gemv_phase(*ly.down, d, SRC_PTR, SRC_PTR, 0);
        }
    }
    GemvArgs lm{};
    lm.x = m.x.p;
    lm.norm_w = m.out_norm.p;
    lm.rms_eps = m.cfg.rms_eps;
    lm.epi = EPI_LOGITS;
    lm.out = m.logits.p;
    m.lm_sharded = tp && m.lm_head_shard != nullptr;
    if (const char* e = getenv("TURBOINFER_B200_TP_LMHEAD")) if (std::string(e) == "replicated") m.lm_sharded = false;   // A/B
    if (m.lm_sharded) {
        // column-parallel lm_head: every rank computes the logits of its V / tp columns (stored at their
        // global position of m.logits, the rest stays zero) and a local arg-max key with GLOBAL indices; the PH_KEYX phase
        // exchanges the P keys over NVLink
        const int Vl = m.cfg.vocab / m.tp;
        lm.out = m.logits.p + (size_t)m.tp_rank * Vl;
        lm.col_off = m.tp_rank * Vl;
        gemv_phase(*m.lm_head_shard, lm, m.layers.empty() ? SRC_EMB : SRC_PTR, SRC_PTR, 1);
        MegaPhase kx{};
        kx.type = PH_KEYX;
        kx.is_head = 1;
        kx.mgpu = 1;
        ph.push_back(kx);
    } else {
        gemv_phase(*m.lm_head, lm, m.layers.empty() ? SRC_EMB : SRC_PTR, SRC_PTR, 1);
    }

    m.mega_max_kpad = max_kpad;
    m.mega_max_units = max_units;
    m.mega_attn_floats = attn_scratch_floats(m.attn_dim, kConsumerThreads);
    if (m.tp_ll) m.mega_attn_floats = std::max(m.mega_attn_floats, 4 * max_units * m.tp);   // the exchange hands (column, rank) values over through this scratch
    m.mega_stages = 0;
    int max_stages = kMaxStages;
    if (const char* e = getenv("TURBOINFER_B200_STAGES")) max_stages = std::max(2, std::min(kMaxStages, atoi(e)));   // experiments
    for (int s = max_stages; s >= 2; --s)
        if (mega_smem_bytes(s, max_kpad, max_units, m.mega_attn_floats, m.num_pages) <= 227 * 1024) { m.mega_stages = s; break; }
    if (m.mega_stages == 0) return fail("persistent kernel does not fit shared memory");
    m.mega_smem = mega_smem_bytes(m.mega_stages, max_kpad, max_units, m.mega_attn_floats, m.num_pages);
    for (auto& p : ph) p.g.stages = m.mega_stages;
    m.nphases = (int)ph.size();
    TRY(m.phases.alloc(ph.size()));
    CK(cudaMemcpyAsync(m.phases.p, ph.data(), ph.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice, g_stream));
    std::vector<ProdRec> prod(ph.size());
    for (size_t i = 0; i < ph.size(); ++i) {
        prod[i].wq = (unsigned long long)(uintptr_t)ph[i].g.wq;
        prod[i].L = ph[i].g.L;
        prod[i].flags = (ph[i].type == PH_GEMV ? 1 : 0) | (ph[i].is_head ? 2 : 0);
    }
    TRY(m.prod.alloc(prod.size()));
    CK(cudaMemcpyAsync(m.prod.p, prod.data(), prod.size() * sizeof(ProdRec), cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    TRY(m.emb_stats.alloc(m.cfg.vocab));
    emb_stats_kernel<<<m.cfg.vocab, 256, 0, g_stream>>>(m.tok_emb.p, m.layers.empty() ? m.out_norm.p : m.layers[0].attn_norm.p,
                                                        m.emb_stats.p, H);
    ++g_launches;
    CK(cudaGetLastError());
    TRY(m.sync_buf.alloc(kBarWords * kBarStride + 8));
    TRY(m.head_cnt.alloc(std::max(1, m.attn_heads)));
    CK(cudaMemsetAsync(m.sync_buf.p, 0, (kBarWords * kBarStride + 8) * sizeof(unsigned int), g_stream));
    CK(cudaMemsetAsync(m.head_cnt.p, 0, std::max(1, m.attn_heads) * sizeof(unsigned int), g_stream));
    CK(cudaStreamSynchronize(g_stream));
    int bits = m.cfg.qtype == TI_Q_INT4 ? 4 : 8;
    int per_sm = 0;
    if (bits == 4) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mega_decode_kernel<4>, kMegaThreads, m.mega_smem));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mega_decode_kernel<8>, kMegaThreads, m.mega_smem));
    if (per_sm < 1) return fail("persistent kernel cannot be resident (occupancy 0)");
    return 0;
}

// n_steps forward passes in one cooperative launch; steps < n_prompt take their token from m.prompt, later ones from
// the previous step's argmax; steps >= first_sample run the lm_head and publish a token
int run_mega(Model& m, int n_prompt, int n_steps, int first_sample) {
    if (n_steps <= 0) return 0;
    CK(cudaMemsetAsync(m.sync_buf.p, 0, (kBarWords * kBarStride + 8) * sizeof(unsigned int), g_stream));
    MegaArgs a{};
    a.phases = m.phases.p;
    a.prod = m.prod.p;
    a.nphases = m.nphases;
    a.emb = m.tok_emb.p;
    a.H = m.cfg.hidden;
    a.V = m.cfg.vocab;
    a.st = m.state.p;
    a.io = m.io.p;
    a.prompt = m.prompt.p;
    a.n_prompt = n_prompt;
    a.n_steps = n_steps;
    a.first_sample = first_sample;
    a.grid_bar = m.sync_buf.p;
    a.head_cnt = m.head_cnt.p;
    a.keys = reinterpret_cast<unsigned long long*>(m.sync_buf.p + kBarWords * kBarStride);
    a.logits = m.logits.p;
    a.stages = m.mega_stages;
    a.max_kpad = m.mega_max_kpad;
    a.max_units = m.mega_max_units;
    a.attn_floats = m.mega_attn_floats;
    a.dbg = m.dbg_on ? m.dbg.p : nullptr;
    a.emb_stats = m.emb_stats.p;
    a.kv_page_table = m.page_table.p;
    a.kv_pages = m.num_pages;
    a.n_head = m.lm_sharded ? 2 : 1;
    a.dbg_flags = getenv("TURBOINFER_B200_DBG_FLAGS") ? atoi(getenv("TURBOINFER_B200_DBG_FLAGS")) : 0;
    if (m.dbg_on && m.dbg_all) a.dbg_flags |= 4;   // every CTA stamps
    a.tp = m.tp_fused ? m.tp : 1;
    a.tp_rank = m.tp_rank;
    if (m.tp_fused) {
        for (int r = 0; r < m.tp; ++r) {
            uint8_t* b = static_cast<uint8_t*>(m.peer_base[r]);
            a.peer_bar[r] = reinterpret_cast<unsigned int*>(b);
            a.peer_part[0][r] = reinterpret_cast<float*>(b + m.xchg_part_off);
            a.peer_part[1][r] = reinterpret_cast<float*>(b + m.xchg_part_off + m.xchg_part_bytes);
            a.peer_flag[0][r] = reinterpret_cast<unsigned int*>(b + m.xchg_flag_off);
            a.peer_flag[1][r] = reinterpret_cast<unsigned int*>(b + m.xchg_flag_off) + (size_t)kMaxTp * 256;
            a.peer_keys[r] = reinterpret_cast<unsigned long long*>(b + m.xchg_key_off);
        }
        a.tp_p2p = m.tp_p2p ? 1 : 0;
        a.tp_ll = m.tp_ll ? 1 : 0;
        a.mg_seq = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(m.xchg) + (size_t)kBarWords * kBarStride * sizeof(unsigned int));
    }
    void* args[] = {&a};
    const void* fn = m.cfg.qtype == TI_Q_INT4 ? (const void*)mega_decode_kernel<4> : (const void*)mega_decode_kernel<8>;
    if (m.dbg_on) {   // the timeline instance exists for INT4 only (compile time)
        if (m.cfg.qtype != TI_Q_INT4) return fail("the debug timeline is compiled for INT4 models only");
        fn = (const void*)mega_decode_kernel<4, true>;
    }
    CK(cudaLaunchCooperativeKernel(fn, dim3(g_num_sms), dim3(kMegaThreads), args, m.mega_smem, g_stream));
    ++g_launches;
    return 0;
}

// ---- tensor-core GEMM (prefill / batched decode) ---------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_tmap_u8_2d(CUtensorMap* map, void* base, uint64_t cols, uint64_t rows, int box_rows = kGemmBM) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
        if (!p || qr != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols};          // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

int ensure_kmajor(QWeight& w) {
    if (w.kmajor.p) return 0;
    w.k_pad = (layout_kpad(w.L) + kGemmBK - 1) / kGemmBK * kGemmBK;
    w.n_pad = (4 * w.L.U + kGemmBN - 1) / kGemmBN * kGemmBN;
    TRY(w.kmajor.alloc((size_t)w.n_pad * w.k_pad));
    CK(cudaMemsetAsync(w.kmajor.p, 0, (size_t)w.n_pad * w.k_pad, g_stream));
    unpack_kmajor_kernel<<<w.L.P, kConsumerThreads, 0, g_stream>>>(w.packed.p, w.L, w.k_pad, w.kmajor.p);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

// x_dev [M][K] fp32 -> y_dev [M][N] fp32; scratch buffers are allocated per call (this entry point is not the decode path)
int gemm_q_dev(QWeight& w, const float* x_dev, float* y_dev, int M, float* kernel_ms, int reps) {
    TRY(ensure_kmajor(w));
    const int K = w.L.K, N = w.L.N;
    const int m_pad = (M + kGemmBM - 1) / kGemmBM * kGemmBM;
    DevBuf<int8_t> planes;
    DevBuf<float> sx;
    DevBuf<long long> sxf;
    TRY(planes.alloc((size_t)3 * m_pad * w.k_pad));
    TRY(sx.alloc(M));
    TRY(sxf.alloc(M));
    CK(cudaMemsetAsync(planes.p, 0, (size_t)3 * m_pad * w.k_pad, g_stream));
    gemm_digits_kernel<<<M, 256, 0, g_stream>>>(x_dev, M, K, m_pad, w.k_pad, planes.p, sx.p, sxf.p);
    ++g_launches;
    CK(cudaGetLastError());
    CUtensorMap map_a;
    TRY(make_tmap_u8_2d(&map_a, planes.p, (uint64_t)w.k_pad, (uint64_t)3 * m_pad));
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K;
    g.m_pad = m_pad; g.k_pad = w.k_pad;
    g.a_signed_b = w.L.bits == 8 ? 1 : 0;
    g.woff = w.L.bits == 4 ? w.offset4 : 0;
    g.sx = sx.p; g.sxf = sxf.p;
    g.colscale = w.colscale.p;
    g.colzterm = w.has_zterm ? w.colzterm.p : nullptr;
    g.y = y_dev;
    g.wt = w.kmajor.p;
    const dim3 grid((N + kGemmBN - 1) / kGemmBN, m_pad / kGemmBM);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (kernel_ms) { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventRecord(e0, g_stream)); }
    for (int r = 0; r < std::max(reps, 1); ++r) {
        gemm_i8_tc_kernel<<<grid, kGemmThreads, kGemmSmemBytes, g_stream>>>(map_a, g);
        ++g_launches;
    }
    CK(cudaGetLastError());
    if (kernel_ms) CK(cudaEventRecord(e1, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    if (kernel_ms) {
        CK(cudaEventElapsedTime(kernel_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return 0;
}

// Launch of a kernel of the lockstep step with programmatic stream serialization: the kernel may start while its predecessor is
// still running and blocks in griddepcontrol.wait (every kernel launched through here has one at its top, the 32-row GEMM after its
// shared-memory / TMEM prologue) until the predecessor has completed.  Captured into the step graphs as programmatic edges.
static bool g_batch_pdl = true;
template <typename... KArgs, typename... Args>
int launch_step(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = g_stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_batch_pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
    ++g_launches;
    return 0;
}

// ---- batched prefill on the tensor cores --------------------------------------------------------------------------
// One GEMM of the prefill path: activations are already digit planes; no host synchronisation.
// split factor of the 32-row GEMM: (N / 128) * S CTAs should fill the SMs in whole waves, with >= 6 k-steps per CTA
int small_gemm_splits(int tiles, int ksteps) {
    // One CTA per SM, and a CTA's time is its k-loop (a serial chain of TMA round trips) plus, when split, the atomics + fix-up:
    // splitting pays for GEMMs with few column tiles (o and down projections: 32 tiles on 148 SMs).
    static const int min_steps = getenv("TURBOINFER_B200_SPLIT_KSTEPS") ? std::max(2, atoi(getenv("TURBOINFER_B200_SPLIT_KSTEPS"))) : 8;   // A/B knob; 8 measured best (7B batch 32: 5.06 -> 4.96 ms per step)
    if (ksteps <= 2 * min_steps || tiles * 2 > g_num_sms) return 1;
    int s = std::min(g_num_sms / tiles, ksteps / min_steps);
    return std::max(1, std::min(s, 8));
}
int launch_small_gemm(Model& m, QWeight& w, int m_pad, const GemmArgs& g) {
    const int tiles = (w.L.N + kGemmBN - 1) / kGemmBN;
    if (m.pf_ws.n < (size_t)w.n_pad * kSmallRows || m.pf_cnt.n < (size_t)tiles) return fail("internal: split-K workspace too small");
    SplitKArgs sk{reinterpret_cast<const uint8_t*>(m.pf_planes.p), m.pf_ws.p, m.pf_cnt.p, w.n_pad, nullptr};
    const int S = small_gemm_splits(tiles, w.k_pad / kGemmBK);
    TRY(launch_step(gemm_i8_tc_small_kernel, dim3(tiles, S), dim3(kSmallThreads), kSmallSmemBytes, g, sk));
    --g_launches;   // (the caller counts this launch)
    return 0;
}
int pf_gemm(Model& m, QWeight& w, int M, int m_pad, float* y, const float* resid) {
    TRY(ensure_kmajor(w));
    if (w.k_pad != (layout_kpad(w.L) + kGemmBK - 1) / kGemmBK * kGemmBK) return fail("internal: k_pad mismatch");
    GemmArgs g{};
    g.M = M; g.N = w.L.N; g.K = w.L.K;
    g.m_pad = m_pad; g.k_pad = w.k_pad;
    g.a_signed_b = w.L.bits == 8 ? 1 : 0;
    g.woff = w.L.bits == 4 ? w.offset4 : 0;
    g.sx = m.pf_sx.p; g.sxf = m.pf_sxf.p;
    g.colscale = w.colscale.p;
    g.colzterm = w.has_zterm ? w.colzterm.p : nullptr;
    g.y = y;
    g.resid = resid;
    g.wt = w.kmajor.p;
    if (M <= kSmallRows && m.pf_small_layout) {   // batched decode: 32-row tiles, every SM streams weights
        TRY(launch_small_gemm(m, w, m_pad, g));
        ++g_launches;
        CK(cudaGetLastError());
        return 0;
    }
    CUtensorMap map_a;
    TRY(make_tmap_u8_2d(&map_a, m.pf_planes.p, (uint64_t)w.k_pad, (uint64_t)3 * m_pad));
    const dim3 grid((w.L.N + kGemmBN - 1) / kGemmBN, m_pad / kGemmBM);
    gemm_i8_tc_kernel<<<grid, kGemmThreads, kGemmSmemBytes, g_stream>>>(map_a, g);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}
// gu != nullptr: x is computed on the fly as up * silu(gate) from gu[M][2K] (needs the fast path; callers check pf_digits_fast)
bool pf_digits_fast(int K, int k_pad) { return K % 4 == 0 && k_pad <= 4 * kDigitsThreads * kDigitsVecs; }
int pf_digits(Model& m, const float* x, const float* norm_w, int M, int K, int m_pad, int k_pad, const float* gu = nullptr) {
    m.pf_small_layout = false;
    if (pf_digits_fast(K, k_pad)) {
        // register-resident conversion, 4-byte stores; it writes every column < k_pad of the rows < M, and rows >= M only
        // produce output rows nobody stores: nothing to clear
        rmsnorm_digits_small_kernel<<<M, kDigitsThreads, 0, g_stream>>>(x, gu, norm_w, m.cfg.rms_eps, K, m_pad, k_pad, m.pf_planes.p, m.pf_sx.p, m.pf_sxf.p, 0);
        ++g_launches;
        CK(cudaGetLastError());
        return 0;
    }
    // planes of rows >= M and columns >= K stay zero: only rows < M, columns < k_pad are written -- k_pad differs per
    // weight, so clear explicitly
    CK(cudaMemsetAsync(m.pf_planes.p, 0, (size_t)3 * m_pad * k_pad, g_stream));
    rmsnorm_digits_kernel<<<M, 256, (size_t)K * sizeof(float), g_stream>>>(x, norm_w, m.cfg.rms_eps, M, K, m_pad, k_pad, m.pf_planes.p, m.pf_sx.p, m.pf_sxf.p);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

bool prefill_gemm_eligible(const Model& m, int M) {
    if (M < 32 || m.tp != 1 || m.cfg.attn_mode != 1 || m.cfg.rope_mode == 2 || m.attn_dim > 128 || (m.attn_dim & 3)) return false;
    if (const char* e = getenv("TURBOINFER_B200_PREFILL")) if (std::string(e) == "decode") return false;
    for (auto& ly : m.layers) if (!(ly.qkv && ly.o && ly.gateup && ly.down)) return false;
    return true;
}

// forward_pass over prompt[0 .. M): fills the KV cache for positions pos0 .. pos0 + M - 1 (pos0 = 0 after reset()).
// The hidden states stay in pf_x; the caller runs the last prompt token through the decode engine for the logits.
int ensure_pf_scratch(Model& m, int M);
int prefill_gemm(Model& m, const int* prompt_dev, int M) {
    const int H = m.cfg.hidden, I = std::max(m.cfg.inter, 1);
    const int m_pad = (M + kGemmBM - 1) / kGemmBM * kGemmBM;
    TRY(ensure_pf_scratch(m, M));
    const int pos0 = 0;
    const int rope_dim = m.cfg.rope_mode == 1 ? H / m.cfg.heads : 0;
    embed_rows_kernel<<<M, 256, 0, g_stream>>>(m.tok_emb.p, prompt_dev, m.pf_x.p, H);
    ++g_launches;
    for (auto& ly : m.layers) {
        TRY(ensure_kmajor(*ly.qkv));
        TRY(pf_digits(m, m.pf_x.p, ly.attn_norm.p, M, H, m_pad, ly.qkv->k_pad));
        TRY(pf_gemm(m, *ly.qkv, M, m_pad, m.pf_qkv.p, nullptr));
        rope_kv_kernel<<<M, 256, 0, g_stream>>>(m.pf_qkv.p, M, H, rope_dim, m.inv_freq.p, pos0, ly.k_pool.p, ly.v_pool.p, m.page_table.p, m.page_tokens);
        static const std::string attn_sel = getenv("TURBOINFER_B200_PREFILL_ATTN") ? getenv("TURBOINFER_B200_PREFILL_ATTN") : "";   // A/B: fp32 | tf32
        const bool attn_fp32 = attn_sel == "fp32", attn_tf32 = attn_sel == "tf32";
        const float att_scale = 1.0f / sqrtf((float)m.attn_dim);
        const dim3 tc_grid(m.attn_heads, (M + kTcQ - 1) / kTcQ);
        if (!attn_fp32 && !attn_tf32 && m.attn_dim == 128) {          // tensor cores, split FP16 (prefill.cuh)
            causal_attention_h3_kernel<128><<<tc_grid, kTcThreads, attn_h3_smem_bytes<128>(), g_stream>>>(
                m.pf_qkv.p, M, H, att_scale, pos0, ly.k_pool.p, ly.v_pool.p, m.page_table.p, m.page_tokens, m.pf_attn.p);
        } else if (!attn_fp32 && !attn_tf32 && m.attn_dim == 64) {
            causal_attention_h3_kernel<64><<<tc_grid, kTcThreads, attn_h3_smem_bytes<64>(), g_stream>>>(
                m.pf_qkv.p, M, H, att_scale, pos0, ly.k_pool.p, ly.v_pool.p, m.page_table.p, m.page_tokens, m.pf_attn.p);
        } else if (!attn_fp32 && m.attn_dim == 128) {   // split TF32
            causal_attention_tc_kernel<128><<<dim3(m.attn_heads, (M + kTcQ - 1) / kTcQ), kTcThreads, attn_tc_smem_bytes<128>(), g_stream>>>(
                m.pf_qkv.p, M, H, att_scale, pos0, ly.k_pool.p, ly.v_pool.p, m.page_table.p, m.page_tokens, m.pf_attn.p);
        } else if (!attn_fp32 && m.attn_dim == 64) {
            causal_attention_tc_kernel<64><<<dim3(m.attn_heads, (M + kTcQ - 1) / kTcQ), kTcThreads, attn_tc_smem_bytes<64>(), g_stream>>>(
                m.pf_qkv.p, M, H, att_scale, pos0, ly.k_pool.p, ly.v_pool.p, m.page_table.p, m.page_tokens, m.pf_attn.p);
        } else {
            causal_attention_kernel<<<dim3(m.attn_heads, (M + kPfQ - 1) / kPfQ), kPfThreads, kPfSmemBytes, g_stream>>>(
                m.pf_qkv.p, M, H, m.attn_dim, att_scale, pos0, ly.k_pool.p, ly.v_pool.p, m.page_table.p, m.page_tokens, m.pf_attn.p);
        }
        g_launches += 2;
        TRY(ensure_kmajor(*ly.o));
        TRY(pf_digits(m, m.pf_attn.p, nullptr, M, H, m_pad, ly.o->k_pad));
        TRY(pf_gemm(m, *ly.o, M, m_pad, m.pf_x.p, m.pf_x.p));
        TRY(ensure_kmajor(*ly.gateup));
        TRY(pf_digits(m, m.pf_x.p, ly.ffn_norm.p, M, H, m_pad, ly.gateup->k_pad));
        TRY(pf_gemm(m, *ly.gateup, M, m_pad, m.pf_gu.p, nullptr));
        TRY(ensure_kmajor(*ly.down));
        if (ly.has_gate && pf_digits_fast(I, ly.down->k_pad)) {
            TRY(pf_digits(m, nullptr, nullptr, M, I, m_pad, ly.down->k_pad, m.pf_gu.p));   // SwiGLU fused into the conversion
        } else {
            if (ly.has_gate) swiglu_rows_kernel<<<grid_for((size_t)M * I), 256, 0, g_stream>>>(m.pf_gu.p, m.pf_act.p, (size_t)M, (size_t)I);
            else relu_rows_kernel<<<grid_for((size_t)M * I), 256, 0, g_stream>>>(m.pf_gu.p, m.pf_act.p, (size_t)M * I);
            ++g_launches;
            TRY(pf_digits(m, m.pf_act.p, nullptr, M, I, m_pad, ly.down->k_pad));
        }
        TRY(pf_gemm(m, *ly.down, M, m_pad, m.pf_x.p, m.pf_x.p));
    }
    CK(cudaGetLastError());
    return 0;
}

// ---- batched decode on the tensor cores (generate_batch, SURVEY.md 8 a5 / 8e "DP: batched-sequence path") ------------
int ensure_pf_scratch(Model& m, int M) {
    const int H = m.cfg.hidden, I = std::max(m.cfg.inter, 1);
    const int m_pad = (M + kGemmBM - 1) / kGemmBM * kGemmBM;
    int kmax = 0;
    auto upd = [&](QWeight* w) { if (w) kmax = std::max(kmax, (layout_kpad(w->L) + kGemmBK - 1) / kGemmBK * kGemmBK); };
    for (auto& ly : m.layers)
        for (QWeight* w : {ly.qkv.get(), ly.o.get(), ly.gateup.get(), ly.down.get()}) upd(w);
    upd(m.lm_head.get());
    if (M > m.pf_cap || m.pf_planes.n < (size_t)3 * m_pad * kmax) {
        TRY(m.pf_x.alloc((size_t)M * H));
        TRY(m.pf_qkv.alloc((size_t)M * 3 * H));
        TRY(m.pf_attn.alloc((size_t)M * H));
        TRY(m.pf_gu.alloc((size_t)M * 2 * I));
        TRY(m.pf_act.alloc((size_t)M * I));
        TRY(m.pf_sx.alloc(M));
        TRY(m.pf_sxf.alloc(M));
        TRY(m.pf_planes.alloc((size_t)3 * m_pad * kmax));
        m.pf_cap = M;
        ++m.pf_gen;
    }
    size_t nmax = 0;
    auto updn = [&](QWeight* w) { if (w) nmax = std::max(nmax, (size_t)(4 * w->L.U + kGemmBN - 1) / kGemmBN * kGemmBN); };
    for (auto& ly : m.layers)
        for (QWeight* w : {ly.qkv.get(), ly.o.get(), ly.gateup.get(), ly.down.get()}) updn(w);
    updn(m.lm_head.get());
    if (m.pf_ws.n < nmax * kSmallRows) {
        TRY(m.pf_ws.alloc(nmax * kSmallRows));
        TRY(m.pf_cnt.alloc(nmax / kGemmBN));
        CK(cudaMemsetAsync(m.pf_ws.p, 0, nmax * kSmallRows * sizeof(unsigned long long), g_stream));
        CK(cudaMemsetAsync(m.pf_cnt.p, 0, nmax / kGemmBN * sizeof(unsigned int), g_stream));
        ++m.pf_gen;
    }
    return 0;
}

bool batch_eligible(const Model& m) {
    if (m.cfg.compat_literal || m.cfg.rope_mode == 2 || !m.lm_head || (m.tp > 1 && (!g_comm || m.cfg.attn_mode != 1))) return false;
    for (auto& ly : m.layers) if (!(ly.qkv && ly.o && ly.gateup && ly.down)) return false;
    return true;
}

// digit planes for the batched decode step (M <= 32 rows read by the GEMM): gu != nullptr fuses SwiGLU in front
int batch_digits(Model& m, const float* x, const float* gu, const float* norm_w, int B, int K, int m_pad, int k_pad) {
    const bool small = K % 4 == 0 && k_pad <= 4 * kDigitsThreads * kDigitsVecs;
    if (small && B <= kSmallRows) {
        TRY(launch_step(rmsnorm_digits_small_kernel, dim3(B), dim3(kDigitsThreads), 0, x, gu, norm_w, m.cfg.rms_eps, K, m_pad, k_pad, m.pf_planes.p, m.pf_sx.p,
                        m.pf_sxf.p, 1));
        m.pf_small_layout = true;
        CK(cudaGetLastError());
        return 0;
    }
    if (gu) {
        swiglu_rows_kernel<<<grid_for((size_t)B * K), 256, 0, g_stream>>>(gu, m.pf_act.p, (size_t)B, (size_t)K);
        ++g_launches;
        x = m.pf_act.p;
    }
    return pf_digits(m, x, norm_w, B, K, m_pad, k_pad);
}

// one lockstep step of all B sequences: tokens[b] -> KV append at *pos -> (sample: logits, argmax -> tokens[b], out)
// row-parallel GEMM output of the batched step under tensor parallelism: partial [B][H] -> all-reduce (sum) -> x += partial
int batch_allreduce_add(Model& m, BatchState& bs, int B) {
    const size_t n = (size_t)B * m.cfg.hidden;
    const ncclResult_t r = g_nccl.AllReduce(bs.ar.p, bs.ar.p, n, ncclFloat32, ncclSum, g_comm, g_stream);
    if (r != ncclSuccess) return fail("ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
    elementwise_kernel<<<grid_for(n), 256, 0, g_stream>>>(m.pf_x.p, bs.ar.p, m.pf_x.p, n, EW_ADD);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

// one lockstep forward pass of B sequences: bs.tokens at position bs.pos_step[0] -> (lm_head) bs.logits
int batch_forward(Model& m, BatchState& bs, bool lm_head) {
    const int B = bs.B, H = m.cfg.hidden;
    const int Hl = H / m.tp;   // tensor parallel: this rank's heads (q / k / v / attention width) -- SURVEY.md 8e
    const int m_pad = (B + kGemmBM - 1) / kGemmBM * kGemmBM;
    const int rope_dim = m.cfg.rope_mode == 1 ? H / m.cfg.heads : 0;
    const bool tp = m.tp > 1;
    TRY(launch_step(embed_rows_kernel, dim3(B), dim3(256), 0, (const float*)m.tok_emb.p, (const int*)bs.tokens.p, m.pf_x.p, H));
    for (size_t l = 0; l < m.layers.size(); ++l) {
        Layer& ly = m.layers[l];
        const int Il = ly.down->L.K;   // this rank's share of the intermediate width
        TRY(batch_digits(m, m.pf_x.p, nullptr, ly.attn_norm.p, B, H, m_pad, ly.qkv->k_pad));
        TRY(pf_gemm(m, *ly.qkv, B, m_pad, m.pf_qkv.p, nullptr));
        TRY(launch_step(rope_kv_batch_kernel, dim3(B, std::max(1, Hl / 512)), dim3(256), 0, m.pf_qkv.p, Hl, rope_dim, (const float*)m.inv_freq.p, (const int*)bs.pos_step.p,
                        bs.k[l].p, bs.v[l].p, (const int*)bs.tables.p, bs.pages_per_seq, m.page_tokens));
        AttnArgs a{};
        a.q = m.pf_qkv.p;
        a.k_pool = bs.k[l].p;
        a.v_pool = bs.v[l].p;
        a.page_table = bs.tables.p;
        a.page_tokens = m.page_tokens;
    a.page_shift = log2_if_pow2(m.page_tokens);
        a.pos_ptr = bs.pos_step.p;
        a.t_bias = 1;
        a.H = Hl;
        a.D = m.attn_dim;
        a.heads = m.attn_heads;
        a.max_splits = bs.max_splits;
        a.min_chunk = 64;
        a.scale = 1.0f / sqrtf((float)m.attn_dim);
        a.part_o = bs.part_o.p;
        a.part_ml = bs.part_ml.p;
        a.out = m.pf_attn.p;
        a.zq = 3 * Hl;
        a.zout = Hl;
        a.ztable = bs.pages_per_seq;
        a.zpart_o = (size_t)m.attn_heads * bs.max_splits * m.attn_dim;
        a.zpart_ml = (size_t)m.attn_heads * bs.max_splits * 2;
        TRY(launch_step(attn_partial_kernel, dim3(m.attn_heads, bs.max_splits, B), dim3(kAttnThreads), m.attn_smem, a));
        TRY(launch_step(attn_combine_kernel, dim3(m.attn_heads, 1, B), dim3(256), 0, a));
        TRY(batch_digits(m, m.pf_attn.p, nullptr, nullptr, B, Hl, m_pad, ly.o->k_pad));
        if (tp) {   // row-parallel: partial sums of all ranks, then the residual
            TRY(pf_gemm(m, *ly.o, B, m_pad, bs.ar.p, nullptr));
            TRY(batch_allreduce_add(m, bs, B));
        } else {
            TRY(pf_gemm(m, *ly.o, B, m_pad, m.pf_x.p, m.pf_x.p));
        }
        TRY(batch_digits(m, m.pf_x.p, nullptr, ly.ffn_norm.p, B, H, m_pad, ly.gateup->k_pad));
        TRY(pf_gemm(m, *ly.gateup, B, m_pad, m.pf_gu.p, nullptr));
        if (ly.has_gate) {
            TRY(batch_digits(m, nullptr, m.pf_gu.p, nullptr, B, Il, m_pad, ly.down->k_pad));   // SwiGLU fused into the conversion
        } else {
            relu_rows_kernel<<<grid_for((size_t)B * Il), 256, 0, g_stream>>>(m.pf_gu.p, m.pf_act.p, (size_t)B * Il);
            ++g_launches;
            TRY(batch_digits(m, m.pf_act.p, nullptr, nullptr, B, Il, m_pad, ly.down->k_pad));
        }
        if (tp) {
            TRY(pf_gemm(m, *ly.down, B, m_pad, bs.ar.p, nullptr));
            TRY(batch_allreduce_add(m, bs, B));
        } else {
            TRY(pf_gemm(m, *ly.down, B, m_pad, m.pf_x.p, m.pf_x.p));
        }
    }
    if (lm_head) {
        TRY(batch_digits(m, m.pf_x.p, nullptr, m.out_norm.p, B, H, m_pad, m.lm_head->k_pad));
        TRY(pf_gemm(m, *m.lm_head, B, m_pad, bs.logits.p, nullptr));
    }
    CK(cudaGetLastError());
    return 0;
}

int batch_step(Model& m, BatchState& bs, bool sample, int out_stride) {
    const int B = bs.B, V = m.cfg.vocab;
    TRY(launch_step(batch_feed_kernel, dim3((B + 127) / 128), dim3(128), 0, (const int*)bs.prompts.p, (const int*)bs.lens.p, (const int*)bs.pos_step.p, B, bs.tokens.p));
    TRY(batch_forward(m, bs, sample));
    if (sample)
        TRY(launch_step(argmax_rows_kernel, dim3(B), dim3(1024), 0, (const float*)bs.logits.p, V, bs.tokens.p, bs.out.p, out_stride, (const int*)bs.pos_step.p,
                        (const int*)bs.lens.p));
    TRY(launch_step(batch_advance_kernel, dim3(1), dim3(1), 0, bs.pos_step.p, sample ? 1 : 0));
    CK(cudaGetLastError());
    return 0;
}

// Tensor parallel with the lm_head sharded over the vocabulary: a rank holds the logits of its own columns (zero elsewhere), so
// the full rows -- needed only when the caller asks for logits, samples, or scores log-probabilities -- are one all-reduce (sum
// with exact zeros) away.
int gather_sharded_logits(Model& m, float* buf, size_t count) {
    if (!m.lm_sharded) return 0;
    const ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclFloat32, ncclSum, g_comm, g_stream);
    if (r != ncclSuccess) return fail("ncclAllReduce (logits) failed: %s", g_nccl.GetErrorString(r));
    return 0;
}

// ---- the prompt of a generation (forward_pass, :749): fills the KV cache; the last prompt token's logits are left in
// m.logits (steps < n_prompt - 1 skip the lm_head).  Long prompts go through the tensor-core GEMM path (all tokens but the
// last in one batched forward pass), short ones through the decode engine token by token.  m.prompt holds the prompt.
int prompt_pass(Model& m, const int32_t* prompt_host, int n_prompt) {
    if (m.use_mega) {
        if (prefill_gemm_eligible(m, n_prompt - 1)) {
            const int M = n_prompt - 1;
            TRY(prefill_gemm(m, m.prompt.p, M));
            CK(cudaMemcpyAsync(&m.state.p->pos, &M, sizeof(int), cudaMemcpyHostToDevice, g_stream));
            CK(cudaMemcpyAsync(m.prompt.p, prompt_host + M, sizeof(int), cudaMemcpyHostToDevice, g_stream));
            TRY(run_mega(m, 1, 1, 0));
        } else {
            TRY(run_mega(m, n_prompt, n_prompt, n_prompt - 1));
        }
        return 0;
    }
    for (int i = 0; i < n_prompt; ++i) {
        set_token_kernel<<<1, 1, 0, g_stream>>>(m.state.p, m.prompt.p, i);
        ++g_launches;
        TRY(run_step(m, i == n_prompt - 1));
    }
    return 0;
}

// one more forward pass from the token in state.token (decode), logits in m.logits
int decode_pass(Model& m) {
    if (m.use_mega) return run_mega(m, 0, 1, 0);
    return run_step(m, true);
}

int launch_sample(const float* logits, int rows, int V, int ld, float temperature, int top_k, float top_p, uint64_t seed, const int* step_ptr, int step,
                  int* token_out, float* logprob_out, int* hist_tokens, float* hist_logprobs, int hist_stride, int* feed_token) {
    SampleArgs a{};
    a.logits = logits; a.V = V; a.ld = ld;
    a.temperature = temperature; a.top_k = top_k; a.top_p = top_p;
    a.seed = seed; a.step_ptr = step_ptr; a.step = step;
    a.token_out = token_out; a.logprob_out = logprob_out;
    a.hist_tokens = hist_tokens; a.hist_logprobs = hist_logprobs; a.hist_stride = hist_stride;
    a.feed_token = feed_token;
    const size_t row_smem = sample_row_smem_bytes(V);
    a.staged = row_smem ? 1 : 0;
    if (sample_keys_in_registers(V)) sample_kernel<32><<<rows, kSampleThreads, row_smem, g_stream>>>(a);
    else sample_kernel<0><<<rows, kSampleThreads, row_smem, g_stream>>>(a);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

// ---- compat_literal: BASELINE.json configs[0], the path benchmarks/benchmark_inference runs today --------------------
// The unmodified InferenceEngine::generate on create_test_model (benchmarks/benchmark_inference.cpp:145-225): placeholder
// embeddings 0.1f*(i%100) over the flattened [1,T,H] index (token ids never reach the arithmetic), the
// attention fall-back (a projection is missing: compute_attention returns its input, x <- x + n, :293-296), the FFN
// x <- x + down(relu(up(n))) (or SwiGLU with a gate, :376-401), logits = x . lm_head; after quantize_model the integer
// weights are cast to fp32 WITHOUT their scale.  Everything is fp32 in the reference build's order of
// roundings (matmul_f32_exact_kernel), so the logits -- ties between identical lm_head columns included -- are the
// reference's bit for bit, and the greedy pick can apply the reference's own rule to them (literal_pick).
int literal_quantize(DevBuf<float>& w, int qtype) {
    if (!w.p || (qtype != TI_Q_INT8 && qtype != TI_Q_INT4)) return 0;
    DevBuf<float> sz;
    DevBuf<uint32_t> mm;
    TRY(sz.alloc(2));
    TRY(mm.alloc(2));
    TRY(device_quant_params(w.p, w.n, qtype, 1, sz.p, mm.p, g_stream));
    literal_quant_kernel<<<grid_for(w.n), 256, 0, g_stream>>>(w.p, w.n, qtype, sz.p);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}
int literal_finalize(Model& m) {
    const int H = m.cfg.hidden, V = m.cfg.vocab;
    if (!m.raw_lm.present()) return fail("lm_head.weight missing");
    auto move = [](RawTensor& r, DevBuf<float>& d) { d = std::move(r.data); };
    move(m.raw_lm, m.lit_lm);
    TRY(literal_quantize(m.lit_lm, m.cfg.qtype));
    for (auto& ly : m.layers) {
        if (ly.raw_q.present() && ly.raw_k.present() && ly.raw_v.present() && ly.raw_o.present())
            return fail("compat_literal reproduces benchmark_inference's model (no o_proj: the attention fall-back); a layer with a complete attention block needs the normal engine");
        ly.raw_q.release(); ly.raw_k.release(); ly.raw_v.release(); ly.raw_o.release();   // unused by the reference too
        if (ly.raw_up.present() && ly.raw_down.present()) {
            move(ly.raw_up, ly.lit_up);
            move(ly.raw_down, ly.lit_down);
            if (ly.raw_gate.present()) move(ly.raw_gate, ly.lit_gate);
            TRY(literal_quantize(ly.lit_up, m.cfg.qtype));
            TRY(literal_quantize(ly.lit_down, m.cfg.qtype));
            TRY(literal_quantize(ly.lit_gate, m.cfg.qtype));
        }
        ly.raw_up.release(); ly.raw_down.release(); ly.raw_gate.release();
    }
    TRY(m.logits.alloc(V));
    TRY(m.state.alloc(1));
    CK(cudaMemsetAsync(m.state.p, 0, sizeof(StepState), g_stream));
    (void)H;
    m.use_mega = false;
    m.finalized = true;
    m.host_pos = 0;
    return 0;
}
// forward_pass over T rows (:1429-1491); only the last row's logits are sampled (:1571-1576)
int literal_forward(Model& m, int T) {
    const size_t H = m.cfg.hidden, I = std::max(m.cfg.inter, 1), V = m.cfg.vocab;
    if (T > m.lit_rows) {
        for (DevBuf<float>* b : {&m.lit_x, &m.lit_pa, &m.lit_n, &m.lit_f}) TRY(b->alloc((size_t)T * H));
        for (DevBuf<float>* b : {&m.lit_u, &m.lit_g}) TRY(b->alloc((size_t)T * I));
        m.lit_rows = T;
    }
    auto mm = [&](const float* a, const float* b, float* c, size_t M, size_t K, size_t N) {
        matmul_f32_exact_kernel<<<dim3((unsigned)((N + 127) / 128), (unsigned)M), 128, 0, g_stream>>>(a, b, c, (int)M, (int)K, (int)N, 0);
        ++g_launches;
    };
    auto ew = [&](const float* a, const float* b, float* y, size_t n, int op) {
        elementwise_kernel<<<grid_for(n), 256, 0, g_stream>>>(a, b, y, n, op);
        ++g_launches;
    };
    auto norm = [&](const float* x, const DevBuf<float>& w) -> const float* {
        if (!w.p) return x;
        rms_norm_kernel<<<(unsigned)T, 256, 0, g_stream>>>(x, w.p, m.lit_n.p, (int)H, m.cfg.rms_eps);
        ++g_launches;
        return m.lit_n.p;
    };
    literal_embed_kernel<<<grid_for((size_t)T * H), 256, 0, g_stream>>>(m.lit_x.p, (size_t)T * H);
    ++g_launches;
    for (auto& ly : m.layers) {
        ew(m.lit_x.p, norm(m.lit_x.p, ly.attn_norm), m.lit_pa.p, (size_t)T * H, EW_ADD);           // x + attention fall-back (:264, :293-296)
        const float* f = norm(m.lit_pa.p, ly.ffn_norm);
        const float* ffn = f;                                                                        // compute_ffn returns its input (:377-380)
        if (ly.lit_up.p && ly.lit_down.p) {
            mm(f, ly.lit_up.p, m.lit_u.p, T, H, I);
            if (ly.lit_gate.p) {
                mm(f, ly.lit_gate.p, m.lit_g.p, T, H, I);
                ew(m.lit_g.p, m.lit_u.p, m.lit_u.p, (size_t)T * I, EW_SILU_MUL);                     // multiply(up, silu(gate))
            } else {
                ew(m.lit_u.p, nullptr, m.lit_u.p, (size_t)T * I, EW_RELU);                           // :392-395
            }
            mm(m.lit_u.p, ly.lit_down.p, m.lit_f.p, T, I, H);
            ffn = m.lit_f.p;
        }
        ew(m.lit_pa.p, ffn, m.lit_x.p, (size_t)T * H, EW_ADD);
    }
    // final norm over all rows (the fall-back norm buffer is free again), then the last row against lm_head
    const float* hn = norm(m.lit_x.p, m.out_norm);
    mm(hn + (size_t)(T - 1) * H, m.lit_lm.p, m.logits.p, 1, H, V);
    CK(cudaGetLastError());
    return 0;
}
// greedy = the top_k 1 branch of sample_next_token (:1585-1598): (logit, index) pairs, std::sort descending by logit,
// element 0 survives.  The benchmark model's lm_head has identical columns (v and v + 500), so WHICH of the tied maxima
// comes first is decided by the sort itself: apply the same std::sort to the same values.
int literal_pick(const std::vector<float>& logits) {
    std::vector<std::pair<float, int>> p;
    p.reserve(logits.size());
    for (size_t i = 0; i < logits.size(); ++i) p.emplace_back(logits[i], static_cast<int>(i));
    std::sort(p.begin(), p.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    return p[0].second;
}
int literal_generate(Model& m, int n_prompt, int n_new, int stop_on_eos, int32_t* out_tokens, int32_t* n_out, float* logits_host, float* decode_ms) {
    const int V = m.cfg.vocab;
    std::vector<float> logits(V);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    TRY(literal_forward(m, n_prompt));   // prefill (:749)
    CK(cudaEventRecord(e0, g_stream));
    int produced = 0, total = n_prompt;
    for (int i = 0; i < n_new; ++i) {
        CK(cudaMemcpyAsync(logits.data(), m.logits.p, (size_t)V * 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (logits_host) memcpy(logits_host + (size_t)i * V, logits.data(), (size_t)V * 4);
        const int best = literal_pick(logits);
        out_tokens[produced++] = best;
        ++total;
        if (stop_on_eos && best == 2) break;        // :760
        if (total >= m.cfg.max_seq) break;          // max_sequence_length (:767)
        if (i + 1 < n_new) TRY(literal_forward(m, 1));   // decode (:774): one row, position-independent
    }
    CK(cudaEventRecord(e1, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (decode_ms) *decode_ms = ms;
    if (n_out) *n_out = produced;
    m.host_pos = std::min(total, m.cfg.max_seq);
    return 0;
}

}  // namespace

// =====================================================================================================
// extern "C" surface
// =====================================================================================================
extern "C" {

int ti_b200_abi_version(void) { return TI_B200_ABI_VERSION; }
const char* ti_b200_last_error(void) { return g_err.c_str(); }

int ti_b200_device_count(int* n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *n = 0; return fail("cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *n = c;
    return 0;
}

int ti_b200_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail("ti_b200_init: no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (device < 0 || device >= count) return fail("ti_b200_init: device %d out of range [0,%d)", device, count);
    if (g_device == device) return 0;
    // one device per process (one process per GPU): the stream, the models and the kernel attributes belong
    // to the device of the first init; switching needs ti_b200_shutdown first
    if (g_device >= 0) return fail("ti_b200_init: already initialised on device %d; call ti_b200_shutdown before selecting device %d", g_device, device);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("ti_b200_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    g_num_sms = prop.multiProcessorCount;
    if (!g_stream) CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_device = device;
    const char* pdl = getenv("TURBOINFER_B200_PDL");
    g_use_pdl = pdl ? atoi(pdl) != 0 : true;   // programmatic dependent launch of the stand-alone GEMV: on unless TURBOINFER_B200_PDL=0
    if (const char* e = getenv("TURBOINFER_B200_EOS_CHUNK")) kEosChunk = std::max(1, atoi(e));
    if (const char* e = getenv("TURBOINFER_B200_BATCH_PDL")) g_batch_pdl = atoi(e) != 0;   // A/B: programmatic launches inside the lockstep step
    return set_kernel_attrs();
}

int ti_b200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device < 0) return 0;
    cudaStreamSynchronize(g_stream);
    g_models.clear();
    g_qweights.clear();
    g_kvs.clear();
    if (g_comm) {   // the tensor-parallel group dies with the library state: a fresh init may form a new one
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
        g_tp_size = 1;
        g_tp_rank = 0;
    }
    cudaStreamDestroy(g_stream);
    g_stream = nullptr;
    g_device = -1;
    g_attr_done = false;
    g_batch_carveout_done = false;
    return 0;
}

int ti_b200_device_info(char* buf, size_t cap) {
    TRY(need_init());
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g_device));
    size_t fr = 0, tot = 0;
    CK(cudaMemGetInfo(&fr, &tot));
    snprintf(buf, cap, "%s sm_%d%d, %d SMs, %.1f GiB HBM (%.1f free), smem/CTA %zu KiB", prop.name, prop.major, prop.minor,
             prop.multiProcessorCount, tot / 1073741824.0, fr / 1073741824.0, prop.sharedMemPerBlockOptin / 1024);
    return 0;
}

static int nccl_load() {
    if (g_nccl.lib) return 0;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail("tensor parallelism needs NCCL: %s", dlerror());
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(lib, "ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(lib, "ncclAllGather"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
        return fail("the NCCL library lacks a required symbol");
    g_nccl.lib = lib;
    return 0;
}

int ti_b200_tp_unique_id(uint8_t* out, size_t cap) {
    TRY(need_init());
    TRY(nccl_load());
    if (cap < sizeof(ncclUniqueId)) return fail("unique id buffer too small: need %zu bytes", sizeof(ncclUniqueId));
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail("ncclGetUniqueId failed: %s", g_nccl.GetErrorString(r));
    memcpy(out, &id, sizeof(id));
    return 0;
}

int ti_b200_tp_init(int nranks, int rank, const uint8_t* unique_id, size_t id_bytes) {
    TRY(need_init());
    TRY(nccl_load());
    if (nranks < 2 || rank < 0 || rank >= nranks) return fail("invalid tensor parallel geometry (%d ranks, rank %d)", nranks, rank);
    if (id_bytes < sizeof(ncclUniqueId)) return fail("unique id too short");
    if (g_comm) return fail("tensor parallel group already initialised");
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    const ncclResult_t r = g_nccl.CommInitRank(&g_comm, nranks, id, rank);
    if (r != ncclSuccess) { g_comm = nullptr; return fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    g_tp_size = nranks;
    g_tp_rank = rank;
    return 0;
}

int ti_b200_sync(void) {
    TRY(need_init());
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}
int ti_b200_malloc(void** dev, size_t bytes) {
    TRY(need_init());
    CK(cudaMalloc(dev, bytes));
    return 0;
}
int ti_b200_free(void* dev) {
    TRY(need_init());
    CK(cudaFree(dev));
    return 0;
}
int ti_b200_memcpy_h2d(void* dev, const void* host, size_t bytes) {
    TRY(need_init());
    CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}
int ti_b200_memcpy_d2h(void* host, const void* dev, size_t bytes) {
    TRY(need_init());
    CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}
int ti_b200_launch_count(uint64_t* n) { *n = g_launches; return 0; }

// ---- quantizer ---------------------------------------------------------------------------------------
int ti_b200_quant_info(const float* x_host, size_t n, int qtype, int symmetric, float* scale, float* zero_point) {
    TRY(need_init());
    if (n == 0) return fail("Cannot quantize an empty tensor");
    if (qtype != TI_Q_INT8 && qtype != TI_Q_INT4) return fail("Unsupported quantization type");
    DevBuf<float> x, sz;
    DevBuf<uint32_t> mm;
    TRY(upload(x, x_host, n));
    TRY(sz.alloc(2));
    TRY(mm.alloc(2));
    TRY(device_quant_params(x.p, n, qtype, symmetric, sz.p, mm.p, g_stream));
    float h[2];
    CK(cudaMemcpyAsync(h, sz.p, sizeof(h), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    *scale = h[0];
    *zero_point = h[1];
    return 0;
}

int ti_b200_quantize(const float* x_host, size_t n, int qtype, float scale, float zero_point, void* q_out_host) {
    TRY(need_init());
    if (qtype != TI_Q_INT8 && qtype != TI_Q_INT4) return fail("Unsupported quantization type");
    if (n == 0) return 0;
    DevBuf<float> x;
    DevBuf<int8_t> q8;
    DevBuf<int32_t> q32;
    TRY(upload(x, x_host, n));
    if (qtype == TI_Q_INT8) TRY(q8.alloc(n)); else TRY(q32.alloc(n));
    quantize_flat_kernel<<<grid_for(n), 256, 0, g_stream>>>(x.p, n, qtype, scale, zero_point, q8.p, q32.p);
    ++g_launches;
    CK(cudaGetLastError());
    if (qtype == TI_Q_INT8) CK(cudaMemcpyAsync(q_out_host, q8.p, n, cudaMemcpyDeviceToHost, g_stream));
    else CK(cudaMemcpyAsync(q_out_host, q32.p, n * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int ti_b200_dequantize(const void* q_host, size_t n, int qtype, float scale, float zero_point, float* x_out_host) {
    TRY(need_init());
    if (qtype != TI_Q_INT8 && qtype != TI_Q_INT4) return fail("Unsupported quantization type for dequantization");
    if (n == 0) return 0;
    DevBuf<float> x;
    DevBuf<int8_t> q8;
    DevBuf<int32_t> q32;
    TRY(x.alloc(n));
    if (qtype == TI_Q_INT8) {
        TRY(q8.alloc(n));
        CK(cudaMemcpyAsync(q8.p, q_host, n, cudaMemcpyHostToDevice, g_stream));
    } else {
        TRY(q32.alloc(n));
        CK(cudaMemcpyAsync(q32.p, q_host, n * 4, cudaMemcpyHostToDevice, g_stream));
    }
    dequantize_flat_kernel<<<grid_for(n), 256, 0, g_stream>>>(q8.p, q32.p, n, qtype, scale, zero_point, x.p);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(x_out_host, x.p, n * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int ti_b200_quantize_pack(const float* w_host, size_t K, size_t N, int qtype, int symmetric, ti_qweight_t* out) {
    TRY(need_init());
    if (K == 0 || N == 0) return fail("Cannot quantize an empty tensor");
    if (qtype != TI_Q_INT8 && qtype != TI_Q_INT4) return fail("Unsupported quantization type");
    DevBuf<float> w;
    TRY(upload(w, w_host, K * N));
    const float* src[1] = {w.p};
    int n[1] = {(int)N};
    std::unique_ptr<QWeight> q;
    TRY(build_qweight(src, n, 1, 0, (int)K, qtype, symmetric, false, &q));
    std::lock_guard<std::mutex> lk(g_mu);
    g_qweights.push_back(std::move(q));
    *out = g_qweights.size();
    return 0;
}

int ti_b200_qweight_info(ti_qweight_t h, size_t* K, size_t* N, int* qtype, float* scale, float* zero_point, size_t* packed_bytes) {
    QWeight* w = get_qw(h);
    if (!w) return fail("invalid qweight handle");
    if (K) *K = w->L.K;
    if (N) *N = w->L.N;
    if (qtype) *qtype = w->qtype;
    if (scale) *scale = w->scale;
    if (zero_point) *zero_point = w->zp;
    if (packed_bytes) *packed_bytes = w->bytes();
    return 0;
}

int ti_b200_qweight_unpack(ti_qweight_t h, int32_t* q_out_host) {
    TRY(need_init());
    QWeight* w = get_qw(h);
    if (!w) return fail("invalid qweight handle");
    DevBuf<int32_t> q;
    const size_t n = (size_t)w->L.K * w->L.N;
    TRY(q.alloc(n));
    unpack_kernel<<<w->L.P, kConsumerThreads, 0, g_stream>>>(w->packed.p, w->L, w->offset4, q.p);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(q_out_host, q.p, n * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int ti_b200_qweight_free(ti_qweight_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (h == 0 || h > g_qweights.size() || !g_qweights[h - 1]) return fail("invalid qweight handle");
    g_qweights[h - 1].reset();
    return 0;
}

// ---- ops ---------------------------------------------------------------------------------------------
int ti_b200_gemv_q_dev(ti_qweight_t h, const float* x_dev, float* y_dev) {
    TRY(need_init());
    QWeight* w = get_qw(h);
    if (!w) return fail("invalid qweight handle");
    GemvArgs a{};
    a.x = x_dev;
    a.epi = EPI_STORE;
    a.out = y_dev;
    return launch_gemv(*w, a, g_stream);
}

int ti_b200_gemv_q(ti_qweight_t h, const float* x_host, float* y_host, size_t rows) {
    TRY(need_init());
    QWeight* w = get_qw(h);
    if (!w) return fail("invalid qweight handle");
    if (rows == 0) return fail("Cannot perform matrix multiplication on empty tensors");
    const size_t K = w->L.K, N = w->L.N;
    DevBuf<float> x, y;
    TRY(upload(x, x_host, rows * K));
    TRY(y.alloc(rows * N));
    for (size_t r = 0; r < rows; ++r) TRY(ti_b200_gemv_q_dev(h, x.p + r * K, y.p + r * N));
    CK(cudaMemcpyAsync(y_host, y.p, rows * N * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int ti_b200_gemm_q(ti_qweight_t h, const float* x_host, float* y_host, size_t rows) {
    TRY(need_init());
    QWeight* w = get_qw(h);
    if (!w) return fail("invalid qweight handle");
    if (rows == 0) return fail("Cannot perform matrix multiplication on empty tensors");
    DevBuf<float> x, y;
    TRY(upload(x, x_host, rows * w->L.K));
    TRY(y.alloc(rows * w->L.N));
    TRY(gemm_q_dev(*w, x.p, y.p, (int)rows, nullptr, 1));
    CK(cudaMemcpyAsync(y_host, y.p, rows * w->L.N * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int ti_b200_bench_gemm(ti_qweight_t h, size_t rows, size_t reps, float* ms, double* ops_per_launch) {
    TRY(need_init());
    QWeight* w = get_qw(h);
    if (!w) return fail("invalid qweight handle");
    if (rows == 0 || reps == 0) return fail("nothing to time");
    DevBuf<float> x, y;
    TRY(x.alloc(rows * w->L.K));
    TRY(y.alloc(rows * w->L.N));
    synth_fill_kernel<<<grid_for(rows * w->L.K), 256, 0, g_stream>>>(x.p, rows * w->L.K, 4242, 1.0f);
    ++g_launches;
    TRY(gemm_q_dev(*w, x.p, y.p, (int)rows, nullptr, 2));   // warm-up (also builds the K-major copy)
    float total = 0.f;
    TRY(gemm_q_dev(*w, x.p, y.p, (int)rows, &total, (int)reps));
    *ms = total / (float)reps;
    if (ops_per_launch) *ops_per_launch = 2.0 * 3.0 * (double)rows * w->L.K * w->L.N;   // three INT8 digit planes
    return 0;
}

int ti_b200_matmul_f32(const float* a_host, const float* b_host, float* c_host, size_t M, size_t K, size_t N) {
    TRY(need_init());
    if (M == 0 || K == 0 || N == 0) return fail("Cannot perform matrix multiplication on empty tensors");
    DevBuf<float> a, b, c;
    TRY(upload(a, a_host, M * K));
    TRY(upload(b, b_host, K * N));
    TRY(c.alloc(M * N));
    matmul_f32_exact_kernel<<<dim3((unsigned)((N + 127) / 128), (unsigned)M), 128, 0, g_stream>>>(a.p, b.p, c.p, (int)M, (int)K, (int)N, 0);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c_host, c.p, M * N * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int ti_b200_rms_norm(const float* x_host, const float* w_host, float* y_host, size_t rows, size_t H, float eps) {
    TRY(need_init());
    if (rows == 0 || H == 0) return fail("Cannot apply RMS normalization to empty tensors");
    DevBuf<float> x, w, y;
    TRY(upload(x, x_host, rows * H));
    TRY(upload(w, w_host, H));
    TRY(y.alloc(rows * H));
    rms_norm_kernel<<<(unsigned)rows, 256, 0, g_stream>>>(x.p, w.p, y.p, (int)H, eps);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y_host, y.p, rows * H * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

static std::vector<float> host_inv_freq(size_t D, float theta) {
    std::vector<float> f(D / 2);
    for (size_t i = 0; i < D / 2; ++i) f[i] = 1.0f / std::pow(theta, static_cast<float>(2 * i) / static_cast<float>(D));  // :1562-1565
    return f;
}

int ti_b200_rope(const float* x_host, const float* pos_host, float* y_host, size_t B, size_t nh, size_t T, size_t D, int ndim,
                 int pos_2d, float theta) {
    TRY(need_init());
    if (ndim != 3 && ndim != 4) return fail("RoPE supports 3D or 4D input tensors only");
    if (D % 2 != 0) return fail("Hidden dimension must be even for RoPE");
    const size_t heads = ndim == 4 ? nh : 1;
    const size_t rows = B * heads * T;
    if (rows == 0 || D == 0) return fail("Cannot apply RoPE to empty tensors");
    DevBuf<float> x, pos, y, fr;
    TRY(upload(x, x_host, rows * D));
    TRY(upload(pos, pos_host, pos_2d ? B * T : T));
    std::vector<float> f = host_inv_freq(D, theta);
    TRY(upload(fr, f.data(), f.size()));
    TRY(y.alloc(rows * D));
    rope_kernel<<<grid_for(rows * D / 2), 256, 0, g_stream>>>(x.p, pos.p, fr.p, y.p, (int)heads, (int)T, (int)D, pos_2d, rows);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y_host, y.p, rows * D * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

static int elementwise_host(const float* a_host, const float* b_host, float* y_host, size_t n, int op) {
    TRY(need_init());
    if (n == 0) return fail("Cannot apply an element-wise operation to empty tensors");
    DevBuf<float> a, b, y;
    TRY(upload(a, a_host, n));
    if (b_host) TRY(upload(b, b_host, n));
    TRY(y.alloc(n));
    elementwise_kernel<<<grid_for(n), 256, 0, g_stream>>>(a.p, b.p, y.p, n, op);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y_host, y.p, n * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}
int ti_b200_silu(const float* x, float* y, size_t n) { return elementwise_host(x, nullptr, y, n, EW_SILU); }
int ti_b200_relu(const float* x, float* y, size_t n) { return elementwise_host(x, nullptr, y, n, EW_RELU); }
int ti_b200_add(const float* a, const float* b, float* y, size_t n) { return elementwise_host(a, b, y, n, EW_ADD); }
int ti_b200_mul(const float* a, const float* b, float* y, size_t n) { return elementwise_host(a, b, y, n, EW_MUL); }
int ti_b200_silu_mul(const float* g, const float* u, float* y, size_t n) { return elementwise_host(g, u, y, n, EW_SILU_MUL); }

int ti_b200_softmax(const float* x_host, float* y_host, size_t rows, size_t n, float temperature) {
    TRY(need_init());
    if (rows == 0 || n == 0) return fail("Cannot apply softmax to empty tensor");
    DevBuf<float> x, y;
    TRY(upload(x, x_host, rows * n));
    TRY(y.alloc(rows * n));
    softmax_kernel<<<(unsigned)rows, 256, 0, g_stream>>>(x.p, y.p, (int)n, temperature);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y_host, y.p, rows * n * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

static size_t attn_smem_bytes(int D) { return sizeof(float) * (size_t)attn_scratch_floats(D, kAttnThreads); }
static int attn_check_dim(size_t H, size_t heads) {
    if (heads == 0 || H % heads != 0) return fail("Hidden size must be divisible by number of heads");
    const size_t D = H / heads;
    if (D % 4 != 0) return fail("attention head dimension %zu must be a multiple of 4", D);
    if (D > (size_t)kAttnThreads * 32) return fail("attention head dimension %zu too large", D);
    return 0;
}

int ti_b200_attention_decode(const float* q_host, const float* k_host, const float* v_host, float* out_host, size_t B, size_t t,
                             size_t H, size_t num_heads) {
    TRY(need_init());
    if (B == 0 || t == 0 || H == 0) return fail("Cannot compute attention with empty tensors");
    TRY(attn_check_dim(H, num_heads));
    const int D = (int)(H / num_heads);
    DevBuf<float> q, k, v, out, po, pml;
    DevBuf<int> table, pos;
    TRY(upload(q, q_host, B * H));
    TRY(upload(k, k_host, B * t * H));
    TRY(upload(v, v_host, B * t * H));
    TRY(out.alloc(B * H));
    const int page_tokens = 64;
    const int pages = (int)((t + page_tokens - 1) / page_tokens);
    // the host tensors are contiguous [t, H]: identity page table over a pool that is the tensor itself
    std::vector<int> tab(pages);
    for (int i = 0; i < pages; ++i) tab[i] = i;
    TRY(table.alloc(pages));
    CK(cudaMemcpyAsync(table.p, tab.data(), pages * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    const int tpos = (int)t;
    TRY(pos.alloc(1));
    CK(cudaMemcpyAsync(pos.p, &tpos, sizeof(int), cudaMemcpyHostToDevice, g_stream));
    const int max_splits = std::max(1, std::min<int>((2 * g_num_sms + (int)num_heads - 1) / (int)num_heads, 512));
    TRY(po.alloc((size_t)num_heads * max_splits * D));
    TRY(pml.alloc((size_t)num_heads * max_splits * 2));
    for (size_t b = 0; b < B; ++b) {
        AttnArgs a{};
        a.q = q.p + b * H;
        a.k_pool = k.p + b * t * H;
        a.v_pool = v.p + b * t * H;
        a.page_table = table.p;
        a.page_tokens = page_tokens;
        a.page_shift = log2_if_pow2(page_tokens);
        a.pos_ptr = pos.p;
        a.t_bias = 0;
        a.H = (int)H;
        a.D = D;
        a.heads = (int)num_heads;
        a.max_splits = max_splits;
        a.min_chunk = 64;
        a.scale = 1.0f / sqrtf((float)D);
        a.part_o = po.p;
        a.part_ml = pml.p;
        a.out = out.p + b * H;
        attn_partial_kernel<<<dim3((unsigned)num_heads, max_splits), kAttnThreads, attn_smem_bytes(D), g_stream>>>(a);
        attn_combine_kernel<<<(unsigned)num_heads, 256, 0, g_stream>>>(a);
        g_launches += 2;
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(out_host, out.p, B * H * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

static int check_capacity(Model& m, int extra);

// ---- KV cache manager (reference KVCache, src/model/inference_engine.cpp:25-172) --------------------------
int ti_b200_kv_create(int32_t layers, int32_t heads, int32_t head_dim, int32_t max_seq, int32_t page_tokens, ti_kv_t* out) {
    TRY(need_init());
    if (!out || layers <= 0 || heads <= 0 || head_dim <= 0 || max_seq <= 0) return fail("invalid KV cache geometry");
    if (head_dim % 4 != 0) return fail("attention head dimension %d must be a multiple of 4", head_dim);
    auto kv = std::make_unique<KvCache>();
    kv->layers = layers; kv->heads = heads; kv->head_dim = head_dim; kv->max_seq = max_seq;
    kv->page_tokens = page_tokens > 0 ? page_tokens : 64;
    kv->num_pages = (max_seq + kv->page_tokens - 1) / kv->page_tokens;
    const size_t H = (size_t)heads * head_dim;
    kv->k.resize(layers);
    kv->v.resize(layers);
    for (int l = 0; l < layers; ++l) {
        TRY(kv->k[l].alloc((size_t)kv->num_pages * kv->page_tokens * H));
        TRY(kv->v[l].alloc((size_t)kv->num_pages * kv->page_tokens * H));
    }
    kv->len.assign(layers, 0);
    std::vector<int> tab(kv->num_pages);
    for (int i = 0; i < kv->num_pages; ++i) tab[i] = kv->num_pages - 1 - i;   // physical order deliberately not the identity
    TRY(kv->page_table.alloc(kv->num_pages));
    CK(cudaMemcpyAsync(kv->page_table.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    TRY(kv->pos.alloc(1));
    kv->max_splits = std::max(1, std::min((2 * g_num_sms + heads - 1) / heads, 512));
    TRY(kv->part_o.alloc((size_t)heads * kv->max_splits * head_dim));
    TRY(kv->part_ml.alloc((size_t)heads * kv->max_splits * 2));
    CK(cudaStreamSynchronize(g_stream));
    std::lock_guard<std::mutex> lk(g_mu);
    g_kvs.push_back(std::move(kv));
    *out = g_kvs.size();
    return 0;
}
int ti_b200_kv_destroy(ti_kv_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (h == 0 || h > g_kvs.size() || !g_kvs[h - 1]) return fail("invalid KV cache handle");
    cudaStreamSynchronize(g_stream);
    g_kvs[h - 1].reset();
    return 0;
}
int ti_b200_kv_reset(ti_kv_t h) {   // KVCache::reset (:57-69): length 0; the pages need no clearing, nothing reads beyond the length
    KvCache* kv = get_kv(h);
    if (!kv) return fail("invalid KV cache handle");
    std::fill(kv->len.begin(), kv->len.end(), 0);
    return 0;
}
int ti_b200_kv_length(ti_kv_t h, int32_t layer, int32_t* current_length, int32_t* max_length) {
    KvCache* kv = get_kv(h);
    if (!kv) return fail("invalid KV cache handle");
    if (layer < 0 || layer >= kv->layers) return fail("Layer index out of bounds for KV cache");
    if (current_length) *current_length = kv->len[layer];
    if (max_length) *max_length = kv->max_seq;
    return 0;
}
// update_incremental (:78-160): k_new / v_new [heads, new_tokens, head_dim] appended at the layer's current length
int ti_b200_kv_append(ti_kv_t h, int32_t layer, const float* k_new_host, const float* v_new_host, int32_t new_tokens) {
    TRY(need_init());
    KvCache* kv = get_kv(h);
    if (!kv) return fail("invalid KV cache handle");
    if (layer < 0 || layer >= kv->layers) return fail("Layer index out of bounds for KV cache");       // :82-84
    if (new_tokens <= 0) return fail("new_tokens must be >= 1");
    if (kv->len[layer] + new_tokens > kv->max_seq) return fail("KV cache overflow: sequence too long");   // :100-102
    const size_t n = (size_t)kv->heads * new_tokens * kv->head_dim;
    DevBuf<float> kn, vn;
    TRY(upload(kn, k_new_host, n));
    TRY(upload(vn, v_new_host, n));
    kv_append_kernel<<<grid_for(n), 256, 0, g_stream>>>(kn.p, vn.p, kv->heads, new_tokens, kv->head_dim, kv->len[layer], kv->k[layer].p, kv->v[layer].p,
                                                       kv->page_table.p, kv->page_tokens);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g_stream));
    kv->len[layer] += new_tokens;
    return 0;
}
// the (full_keys, full_values) update_incremental returns (:132-157): [heads, length, head_dim]
int ti_b200_kv_read(ti_kv_t h, int32_t layer, float* k_out_host, float* v_out_host) {
    TRY(need_init());
    KvCache* kv = get_kv(h);
    if (!kv) return fail("invalid KV cache handle");
    if (layer < 0 || layer >= kv->layers) return fail("Layer index out of bounds for KV cache");
    const int len = kv->len[layer];
    if (len == 0) return 0;
    const size_t n = (size_t)kv->heads * len * kv->head_dim;
    DevBuf<float> ko, vo;
    TRY(ko.alloc(n));
    TRY(vo.alloc(n));
    kv_read_kernel<<<grid_for(n), 256, 0, g_stream>>>(kv->k[layer].p, kv->v[layer].p, kv->page_table.p, kv->page_tokens, kv->heads, len, kv->head_dim, ko.p, vo.p);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(k_out_host, ko.p, n * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaMemcpyAsync(v_out_host, vo.p, n * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}
// attention of one query token over the layer's cached tokens, read in place (flash-decoding, split over the context):
// multi_head_attention with q_len 1 (tensor_engine.cpp:1149-1252); heads = 1: attention_fast_incremental (:1254-1388)
int ti_b200_kv_attention(ti_kv_t h, int32_t layer, const float* q_host, float* out_host) {
    TRY(need_init());
    KvCache* kv = get_kv(h);
    if (!kv) return fail("invalid KV cache handle");
    if (layer < 0 || layer >= kv->layers) return fail("Layer index out of bounds for KV cache");
    const int len = kv->len[layer];
    if (len == 0) return fail("Cannot compute attention with empty tensors");
    const int H = kv->heads * kv->head_dim;
    DevBuf<float> q, out;
    TRY(upload(q, q_host, H));
    TRY(out.alloc(H));
    CK(cudaMemcpyAsync(kv->pos.p, &len, sizeof(int), cudaMemcpyHostToDevice, g_stream));
    AttnArgs a{};
    a.q = q.p;
    a.k_pool = kv->k[layer].p;
    a.v_pool = kv->v[layer].p;
    a.page_table = kv->page_table.p;
    a.page_tokens = kv->page_tokens;
    a.page_shift = log2_if_pow2(kv->page_tokens);
    a.pos_ptr = kv->pos.p;
    a.t_bias = 0;
    a.H = H;
    a.D = kv->head_dim;
    a.heads = kv->heads;
    a.max_splits = kv->max_splits;
    a.min_chunk = 64;
    a.scale = 1.0f / sqrtf((float)kv->head_dim);
    a.part_o = kv->part_o.p;
    a.part_ml = kv->part_ml.p;
    a.out = out.p;
    attn_partial_kernel<<<dim3(kv->heads, kv->max_splits), kAttnThreads, sizeof(float) * (size_t)attn_scratch_floats(kv->head_dim, kAttnThreads), g_stream>>>(a);
    attn_combine_kernel<<<kv->heads, 256, 0, g_stream>>>(a);
    g_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_host, out.p, (size_t)H * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

// ---- fused GEMV entry (north_star item 4: RMSNorm prologue, residual / ReLU / SwiGLU epilogues) --------------------------
int ti_b200_quantize_pack_fused(const float* const* w_host, const size_t* n_cols, int32_t n_src, int32_t interleave, size_t K, int qtype, ti_qweight_t* out) {
    TRY(need_init());
    if (n_src < 1 || n_src > 3 || K == 0) return fail("1 to 3 source matrices with K >= 1 rows");
    if (interleave && n_src != 2) return fail("interleaving takes exactly two sources (gate, up)");
    if (qtype != TI_Q_INT8 && qtype != TI_Q_INT4) return fail("Unsupported quantization type");
    DevBuf<float> w[3];
    const float* src[3] = {nullptr, nullptr, nullptr};
    int n[3] = {0, 0, 0};
    for (int i = 0; i < n_src; ++i) {
        if (n_cols[i] == 0) return fail("Cannot quantize an empty tensor");
        if (interleave && n_cols[i] != n_cols[0]) return fail("interleaved sources need equal widths");
        TRY(upload(w[i], w_host[i], K * n_cols[i]));
        src[i] = w[i].p;
        n[i] = (int)n_cols[i];
    }
    std::unique_ptr<QWeight> q;
    TRY(build_qweight(src, n, n_src, interleave ? 1 : 0, (int)K, qtype, 1, false, &q));
    std::lock_guard<std::mutex> lk(g_mu);
    g_qweights.push_back(std::move(q));
    *out = g_qweights.size();
    return 0;
}
// y = epilogue(dequant(W)^T . prologue(x)): norm_w_host != NULL fuses rms_norm(x, norm_w, eps) in front (:1452-1508);
// epilogue TI_EPI_STORE / TI_EPI_RESIDUAL (y += resid, :1626) / TI_EPI_SWIGLU (interleaved gate/up weight: y[i] = up_i * silu(gate_i),
// N / 2 outputs, :900-923 + :1680) / TI_EPI_RELU (:828)
int ti_b200_gemv_q_ex(ti_qweight_t h, const float* x_host, float* y_host, int32_t epilogue, const float* resid_host, const float* norm_w_host, float eps) {
    TRY(need_init());
    QWeight* w = get_qw(h);
    if (!w) return fail("invalid qweight handle");
    if (epilogue != EPI_STORE && epilogue != EPI_RESIDUAL && epilogue != EPI_SWIGLU && epilogue != EPI_RELU) return fail("unsupported epilogue %d", epilogue);
    if (epilogue == EPI_RESIDUAL && !resid_host) return fail("the residual epilogue needs a residual vector");
    if (epilogue == EPI_SWIGLU && (w->L.N & 1)) return fail("the SwiGLU epilogue needs an interleaved (gate, up) weight with an even number of columns");
    const size_t K = w->L.K, N = w->L.N, n_out = epilogue == EPI_SWIGLU ? N / 2 : N;
    DevBuf<float> x, y, r, nw;
    TRY(upload(x, x_host, K));
    TRY(y.alloc(n_out));
    if (resid_host) TRY(upload(r, resid_host, N));
    if (norm_w_host) TRY(upload(nw, norm_w_host, K));
    GemvArgs a{};
    a.x = x.p;
    a.norm_w = norm_w_host ? nw.p : nullptr;
    a.rms_eps = eps;
    a.epi = epilogue;
    a.out = y.p;
    a.resid = r.p;
    TRY(launch_gemv(*w, a, g_stream));
    CK(cudaMemcpyAsync(y_host, y.p, n_out * 4, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

// forward_pass (:1429-1491) over a prompt: resets the model's KV cache, fills it with the n tokens (tensor-core GEMM path for
// long prompts) and returns the logits of the last position (optional)
int ti_b200_prefill(ti_model_t h, const int32_t* tokens, int32_t n, float* last_logits_host) {
    TRY(need_init());
    Model* mp = get_model(h);
    if (!mp || !mp->finalized) return fail("invalid or unfinalized model handle");
    Model& m = *mp;
    if (m.cfg.compat_literal) return fail("ti_b200_prefill is not available on the literal benchmark path");
    if (n <= 0) return fail("Input tokens cannot be empty");
    for (int i = 0; i < n; ++i)
        if (tokens[i] < 0 || tokens[i] >= m.cfg.vocab) return fail("token id %d out of range", tokens[i]);
    TRY(ti_b200_model_reset(h));
    TRY(check_capacity(m, n));
    if ((int)m.prompt.n < n) TRY(m.prompt.alloc(n));
    CK(cudaMemcpyAsync(m.prompt.p, tokens, n * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    StepIO io{};
    CK(cudaMemcpyAsync(m.io.p, &io, sizeof(io), cudaMemcpyHostToDevice, g_stream));
    TRY(prompt_pass(m, tokens, n));
    m.host_pos = n;
    if (last_logits_host) {
        const size_t V = m.cfg.vocab;
        if (m.lm_sharded) {
            if (m.hist.n < V) TRY(m.hist.alloc(V));
            CK(cudaMemcpyAsync(m.hist.p, m.logits.p, V * 4, cudaMemcpyDeviceToDevice, g_stream));
            TRY(gather_sharded_logits(m, m.hist.p, V));
            CK(cudaMemcpyAsync(last_logits_host, m.hist.p, V * 4, cudaMemcpyDeviceToHost, g_stream));
        } else {
            CK(cudaMemcpyAsync(last_logits_host, m.logits.p, V * 4, cudaMemcpyDeviceToHost, g_stream));
        }
    }
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

// ---- model -------------------------------------------------------------------------------------------
int ti_b200_model_new(const ti_model_config* cfg, ti_model_t* out) {
    TRY(need_init());
    if (!cfg || !out) return fail("null argument");
    if (cfg->hidden <= 0 || cfg->vocab <= 0 || cfg->layers < 0 || cfg->heads <= 0) return fail("invalid model config");
    if (cfg->qtype != TI_Q_INT4 && cfg->qtype != TI_Q_INT8 && !cfg->compat_literal)
        return fail("qtype must be TI_Q_INT4 or TI_Q_INT8 (fp32 weights are only supported with compat_literal)");
    auto m = std::make_unique<Model>();
    m->cfg = *cfg;
    if (m->cfg.rms_eps == 0.f) m->cfg.rms_eps = 1e-5f;
    if (m->cfg.rope_theta == 0.f) m->cfg.rope_theta = 10000.0f;
    if (m->cfg.max_seq <= 0) m->cfg.max_seq = 2048;  // KVCache hard-codes 2048 (inference_engine.cpp:569)
    m->page_tokens = cfg->kv_page_tokens > 0 ? cfg->kv_page_tokens : 64;
    if (cfg->reserved[1] > 1) {   // tensor-parallel model: every rank creates it with the same config after ti_b200_tp_init
        if (g_comm == nullptr || cfg->reserved[1] != g_tp_size) return fail("tensor parallel degree %d needs ti_b200_tp_init with as many ranks first", cfg->reserved[1]);
        if (cfg->attn_mode != 1 || cfg->rope_mode == 2) return fail("tensor parallelism shards attention heads: attn_mode 1 and per-head RoPE only");
        if (cfg->heads % g_tp_size || cfg->inter % g_tp_size || cfg->hidden % (4 * g_tp_size)) return fail("heads, inter and hidden must be divisible by the tensor parallel degree");
        m->tp = g_tp_size;
        m->tp_rank = g_tp_rank;
    }
    m->layers.resize(cfg->layers);
    std::lock_guard<std::mutex> lk(g_mu);
    g_models.push_back(std::move(m));
    *out = g_models.size();
    return 0;
}

int ti_b200_model_set_tensor(ti_model_t h, const char* name, const float* data_host, size_t rows, size_t cols) {
    TRY(need_init());
    Model* m = get_model(h);
    if (!m) return fail("invalid model handle");
    if (m->finalized) return fail("model already finalized");
    if (!data_host || rows * cols == 0) return fail("tensor '%s' is empty", name);
    DevBuf<float> dev;
    TRY(upload(dev, data_host, rows * cols));
    return store_tensor(*m, name, std::move(dev), rows, cols);
}

// An already quantized tensor (Quantizer::quantize_model output, or a tensor of a .tinq file, quantization.cpp:208-333): the
// integers are packed into the streaming layout as they are, with the (scale, zero_point) they were made with.
int ti_b200_model_set_tensor_q(ti_model_t h, const char* name, const void* q_host, size_t rows, size_t cols, int qtype, float scale, float zero_point) {
    TRY(need_init());
    Model* m = get_model(h);
    if (!m) return fail("invalid model handle");
    if (m->finalized) return fail("model already finalized");
    if (!q_host || rows * cols == 0) return fail("tensor '%s' is empty", name);
    if (qtype != TI_Q_INT4 && qtype != TI_Q_INT8) return fail("qtype must be TI_Q_INT4 (int32 elements) or TI_Q_INT8 (int8 elements)");
    if (qtype != m->cfg.qtype) return fail("tensor '%s' is %s but the model was created for %s weights", name, qtype == TI_Q_INT4 ? "INT4" : "INT8",
                                           m->cfg.qtype == TI_Q_INT4 ? "INT4" : "INT8");
    if (!(scale > 0.f) || !std::isfinite(scale) || !std::isfinite(zero_point)) return fail("tensor '%s': scale must be positive and finite", name);
    if (qtype == TI_Q_INT4) {   // the nibble code holds [-8, 7] (zero_point 0) or [0, 15]: anything else would wrap silently when packed
        const int32_t* q4 = static_cast<const int32_t*>(q_host);
        const int32_t lo = zero_point == 0.0f ? -8 : 0, hi = zero_point == 0.0f ? 7 : 15;
        int32_t mn = q4[0], mx = q4[0];
        for (size_t i = 1; i < rows * cols; ++i) { mn = std::min(mn, q4[i]); mx = std::max(mx, q4[i]); }
        if (mn < lo || mx > hi) return fail("tensor '%s': INT4 integers must lie in [%d, %d] (found %d .. %d)", name, lo, hi, mn, mx);
    }
    RawTensor q;
    q.kind = qtype == TI_Q_INT8 ? 1 : 2;
    q.scale = scale;
    q.zp = zero_point;
    const size_t bytes = rows * cols * (qtype == TI_Q_INT8 ? 1 : 4);
    TRY(q.qdata.alloc(bytes));
    CK(cudaMemcpyAsync(q.qdata.p, q_host, bytes, cudaMemcpyHostToDevice, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return store_tensor(*m, name, DevBuf<float>(), rows, cols, &q);
}

int ti_b200_model_set_tensor_synthetic(ti_model_t h, const char* name, size_t rows, size_t cols, uint64_t seed, float amp) {
    TRY(need_init());
    Model* m = get_model(h);
    if (!m) return fail("invalid model handle");
    if (m->finalized) return fail("model already finalized");
    DevBuf<float> dev;
    TRY(dev.alloc(rows * cols));
    synth_fill_kernel<<<grid_for(rows * cols), 256, 0, g_stream>>>(dev.p, rows * cols, seed, amp);
    ++g_launches;
    CK(cudaGetLastError());
    return store_tensor(*m, name, std::move(dev), rows, cols);
}

int ti_b200_model_finalize(ti_model_t h) {
    TRY(need_init());
    Model* mp = get_model(h);
    if (!mp) return fail("invalid model handle");
    Model& m = *mp;
    if (m.finalized) return 0;
    if (m.cfg.compat_literal) {
        if (m.tp > 1) return fail("compat_literal is a single-GPU path");
        return literal_finalize(m);
    }
    const int H = m.cfg.hidden, V = m.cfg.vocab, I = m.cfg.inter;
    if (!m.tok_emb.p) return fail("token_embeddings.weight missing");
    if (!m.lm_head) return fail("lm_head.weight missing");  // the reference draws random logits here (:1544-1549); refuse instead
    const int Hl = H / m.tp;   // attention width owned by this rank
    m.attn_heads = m.cfg.attn_mode == 1 ? m.cfg.heads / m.tp : 1;
    TRY(attn_check_dim(Hl, m.attn_heads));
    m.attn_dim = Hl / m.attn_heads;
    m.attn_smem = attn_smem_bytes(m.attn_dim);
    m.max_splits = std::max(1, std::min((2 * g_num_sms + m.attn_heads - 1) / m.attn_heads, 512));
    m.max_splits = std::max(m.max_splits, g_num_sms / m.attn_heads);
    m.num_pages = (m.cfg.max_seq + m.page_tokens - 1) / m.page_tokens;
    for (auto& ly : m.layers) {
        TRY(pack_ready(m, ly, true));
        if (ly.qkv && ly.o) {
            TRY(ly.k_pool.alloc((size_t)m.num_pages * m.page_tokens * Hl));
            TRY(ly.v_pool.alloc((size_t)m.num_pages * m.page_tokens * Hl));
        }
        // sources that cannot be used (e.g. q/k/v without o_proj) are dropped, like the reference ignores them
        ly.raw_q.release(); ly.raw_k.release(); ly.raw_v.release(); ly.raw_o.release();
        ly.raw_up.release(); ly.raw_gate.release(); ly.raw_down.release();
    }
    // page table: pages are handed out in order by reset(); physical order is deliberately not the identity
    std::vector<int> tab(m.num_pages);
    for (int i = 0; i < m.num_pages; ++i) tab[i] = m.num_pages - 1 - i;
    TRY(m.page_table.alloc(m.num_pages));
    CK(cudaMemcpyAsync(m.page_table.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    TRY(m.state.alloc(1));
    CK(cudaMemsetAsync(m.state.p, 0, sizeof(StepState), g_stream));
    TRY(m.io.alloc(1));
    CK(cudaMemsetAsync(m.io.p, 0, sizeof(StepIO), g_stream));
    TRY(m.x.alloc(H));
    TRY(m.nrm.alloc(H));
    TRY(m.q.alloc(H));
    TRY(m.attn_out.alloc(H));
    TRY(m.act.alloc(std::max(I, 1)));
    TRY(m.logits.alloc(V));
    CK(cudaMemsetAsync(m.logits.p, 0, (size_t)V * sizeof(float), g_stream));   // a sharded lm_head writes its own columns only: the rest stays 0
    TRY(m.part_o.alloc((size_t)m.attn_heads * m.max_splits * m.attn_dim));
    TRY(m.part_ml.alloc((size_t)m.attn_heads * m.max_splits * 2));
    const int rope_dim = m.cfg.rope_mode == 1 ? H / m.cfg.heads : H;
    std::vector<float> f = host_inv_freq(rope_dim, m.cfg.rope_theta);
    TRY(upload(m.inv_freq, f.data(), f.size()));
    CK(cudaStreamSynchronize(g_stream));
    bool complete = true;
    for (auto& ly : m.layers) complete &= (ly.qkv && ly.o && ly.gateup && ly.down);
    const char* eng = getenv("TURBOINFER_B200_ENGINE");
    const bool want_graph = m.cfg.reserved[0] == 1 || (eng && std::string(eng) == "graph");
    bool vec_ok = complete && H % 4 == 0;
    if (complete)
        for (auto& ly : m.layers)
            for (QWeight* w : {ly.qkv.get(), ly.o.get(), ly.gateup.get(), ly.down.get()}) vec_ok &= w->L.K % 4 == 0;
    // tensor parallel: the persistent kernel with the all-reduce fused in (peer memory over NVLink) unless the NCCL
    // baseline of the per-op engine is asked for (TURBOINFER_B200_TP_ENGINE=nccl)
    const char* tpe = getenv("TURBOINFER_B200_TP_ENGINE");
    m.tp_fused = m.tp > 1 && complete && vec_ok && !want_graph && !(tpe && std::string(tpe) == "nccl") && m.cfg.attn_mode == 1 &&
                 H <= kConsumerThreads * g_num_sms;
    m.use_mega = complete && !want_graph && vec_ok && (m.tp == 1 || m.tp_fused);   // the persistent kernel's prologue uses 128-bit loads only
    if (m.tp > 1) TRY(m.ar_tmp.alloc(H));
    if (m.tp_fused) {
        const char* red = getenv("TURBOINFER_B200_TP_REDUCE");   // "p2p" (per column slice) or "barrier" (all GPUs + reduce phase)
        // how the row-parallel partials are reduced: "ll" (default) point-to-point with tagged words, "p2p" point-to-point with
        // flags, "barrier" the barrier across the GPUs + reduce phase
        const std::string mode = red ? std::string(red) : std::string("ll");
        m.tp_p2p = (mode == "p2p" || mode == "ll") && g_num_sms <= 256;
        m.tp_ll = m.tp_p2p && mode == "ll";
        TRY(tp_exchange_setup(m));
    }
    TRY(m.prompt.alloc(16));
    if (m.use_mega) TRY(build_mega(m));
    // tensor-parallel steps hold NCCL all-reduces: NCCL supports stream capture, so they are replayed from a graph as
    // well (TURBOINFER_B200_TP_GRAPH=0 enqueues them directly instead)
    const char* tpg = getenv("TURBOINFER_B200_TP_GRAPH");
    if (m.tp == 1 || !(tpg && atoi(tpg) == 0)) {
        TRY(capture_graph(m, true, &m.graph_decode));
        TRY(capture_graph(m, false, &m.graph_prefill));
    }
    m.finalized = true;
    m.host_pos = 0;
    return 0;
}

int ti_b200_model_free(ti_model_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (h == 0 || h > g_models.size() || !g_models[h - 1]) return fail("invalid model handle");
    cudaStreamSynchronize(g_stream);
    g_models[h - 1].reset();
    return 0;
}

int ti_b200_model_reset(ti_model_t h) {
    TRY(need_init());
    Model* m = get_model(h);
    if (!m || !m->finalized) return fail("invalid or unfinalized model handle");
    CK(cudaMemsetAsync(m->state.p, 0, sizeof(StepState), g_stream));
    m->host_pos = 0;
    return 0;
}

int ti_b200_model_kv_length(ti_model_t h, int32_t* length) {
    Model* m = get_model(h);
    if (!m) return fail("invalid model handle");
    *length = m->host_pos;
    return 0;
}

int ti_b200_model_engine(ti_model_t h, int32_t* persistent) {
    Model* m = get_model(h);
    if (!m || !m->finalized) return fail("invalid or unfinalized model handle");
    *persistent = m->use_mega ? 1 : 0;
    return 0;
}

int ti_b200_model_step_bytes(ti_model_t h, int32_t t, double* weight_bytes, double* kv_bytes) {
    Model* m = get_model(h);
    if (!m || !m->finalized) return fail("invalid or unfinalized model handle");
    double wb = 0, kb = 0;
    const double H = m->cfg.hidden / m->tp;   // KV rows are as wide as this rank's heads
    auto add = [&](const std::unique_ptr<QWeight>& w) {
        if (!w) return;
        wb += (double)w->L.K * w->L.N * w->L.bits / 8.0 + 4.0 * w->L.N + 4.0 * (w->L.K + w->L.N);
    };
    for (auto& ly : m->layers) {
        add(ly.qkv); add(ly.o); add(ly.gateup); add(ly.down);
        if (ly.qkv && ly.o) kb += 2.0 * t * H * 4.0 + 2.0 * H * 4.0;  // K,V read over t tokens + one row written
    }
    add(m->lm_sharded ? m->lm_head_shard : m->lm_head);
    wb += H * 4.0;  // embedding row
    if (weight_bytes) *weight_bytes = wb;
    if (kv_bytes) *kv_bytes = kb;
    return 0;
}

static int check_capacity(Model& m, int extra) {
    if (m.host_pos + extra > m.cfg.max_seq) return fail("KV cache overflow: sequence too long");  // inference_engine.cpp:100-102
    return 0;
}

int ti_b200_decode_step(ti_model_t h, int32_t token, float* logits_host, int32_t* argmax) {
    TRY(need_init());
    Model* m = get_model(h);
    if (!m || !m->finalized) return fail("invalid or unfinalized model handle");
    if (token < 0 || token >= m->cfg.vocab) return fail("token id %d out of range", token);
    TRY(check_capacity(*m, 1));
    if (m->cfg.compat_literal) {   // one row, independent of the token and of the position
        TRY(literal_forward(*m, 1));
        std::vector<float> lg(m->cfg.vocab);
        CK(cudaMemcpyAsync(lg.data(), m->logits.p, lg.size() * 4, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        if (logits_host) memcpy(logits_host, lg.data(), lg.size() * 4);
        if (argmax) *argmax = literal_pick(lg);
        m->host_pos += 1;
        return 0;
    }
    StepIO io{};
    CK(cudaMemcpyAsync(m->io.p, &io, sizeof(io), cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(&m->state.p->token, &token, sizeof(int), cudaMemcpyHostToDevice, g_stream));
    const int zero = 0;
    CK(cudaMemcpyAsync(&m->state.p->step, &zero, sizeof(int), cudaMemcpyHostToDevice, g_stream));
    if (m->use_mega) {
        TRY(run_mega(*m, 0, 1, 0));
    } else {
        TRY(run_step(*m, true));
    }
    m->host_pos += 1;
    if (logits_host) {
        if (m->lm_sharded) {   // into a scratch copy: m->logits must keep its zeros outside this rank's columns
            if (m->hist.n < (size_t)m->cfg.vocab) TRY(m->hist.alloc((size_t)m->cfg.vocab));
            CK(cudaMemcpyAsync(m->hist.p, m->logits.p, (size_t)m->cfg.vocab * 4, cudaMemcpyDeviceToDevice, g_stream));
            TRY(gather_sharded_logits(*m, m->hist.p, (size_t)m->cfg.vocab));
            CK(cudaMemcpyAsync(logits_host, m->hist.p, (size_t)m->cfg.vocab * 4, cudaMemcpyDeviceToHost, g_stream));
        } else {
            CK(cudaMemcpyAsync(logits_host, m->logits.p, (size_t)m->cfg.vocab * 4, cudaMemcpyDeviceToHost, g_stream));
        }
    }
    int tok = 0;
    CK(cudaMemcpyAsync(&tok, &m->state.p->token, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    if (argmax) *argmax = tok;
    return 0;
}

int ti_b200_generate_greedy(ti_model_t h, const int32_t* prompt, int32_t n_prompt, int32_t n_new, int32_t stop_on_eos,
                            int32_t* out_tokens, int32_t* n_out, float* logits_host, float* decode_ms) {
    TRY(need_init());
    Model* mp = get_model(h);
    if (!mp || !mp->finalized) return fail("invalid or unfinalized model handle");
    Model& m = *mp;
    if (n_prompt <= 0) return fail("Input tokens cannot be empty");  // validate_input_tokens (:1409)
    if (n_new < 0) return fail("n_new must be >= 0");
    for (int i = 0; i < n_prompt; ++i)
        if (prompt[i] < 0 || prompt[i] >= m.cfg.vocab) return fail("token id %d out of range", prompt[i]);
    TRY(ti_b200_model_reset(h));  // generate() resets the KV cache first (:746)
    if (m.cfg.compat_literal) return literal_generate(m, n_prompt, n_new, stop_on_eos, out_tokens, n_out, logits_host, decode_ms);
    TRY(check_capacity(m, n_prompt + std::max(0, n_new - 1)));
    const int V = m.cfg.vocab;
    if ((int)m.prompt.n < n_prompt) TRY(m.prompt.alloc(n_prompt));
    if ((int)m.out_tokens.n < std::max(n_new, 1)) TRY(m.out_tokens.alloc(std::max(n_new, 1)));
    CK(cudaMemcpyAsync(m.prompt.p, prompt, n_prompt * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    StepIO io{};
    io.out_tokens = m.out_tokens.p;
    io.out_cap = n_new;
    if (logits_host && n_new > 0) {
        if (m.hist.n < (size_t)n_new * V) TRY(m.hist.alloc((size_t)n_new * V));
        io.hist = m.hist.p;
        io.hist_cap = n_new;
    }
    CK(cudaMemcpyAsync(m.io.p, &io, sizeof(io), cudaMemcpyHostToDevice, g_stream));
    cudaEvent_t e0, e1, ep;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&ep));
    CK(cudaEventRecord(ep, g_stream));
    int steps_done = 0;   // decode steps after the prompt
    if (m.use_mega) {
        // launch 1: the prompt (its last step picks token 0); launch 2: the decode loop, timed.
        // Long prompts (>= 32 tokens besides the last one, the reference's own GEMM threshold, tensor_engine.cpp:561) go
        // through the tensor-core GEMM path: all tokens but the last fill the KV cache in one batched forward pass, the
        // last one runs through the decode engine and yields the first logits.
        if (prefill_gemm_eligible(m, n_prompt - 1)) {
            const int M = n_prompt - 1;
            TRY(prefill_gemm(m, m.prompt.p, M));
            CK(cudaMemcpyAsync(&m.state.p->pos, &M, sizeof(int), cudaMemcpyHostToDevice, g_stream));
            CK(cudaMemcpyAsync(m.prompt.p, prompt + M, sizeof(int), cudaMemcpyHostToDevice, g_stream));
            TRY(run_mega(m, 1, 1, n_new > 0 ? 0 : 1));
        } else {
            TRY(run_mega(m, n_prompt, n_prompt, n_new > 0 ? n_prompt - 1 : n_prompt));
        }
        CK(cudaEventRecord(e0, g_stream));
        if (stop_on_eos && n_new - 1 > kEosChunk) {
            // generate() stops at the first EOS (:760).  The persistent launch cannot be cut short from inside (its producer warp
            // runs ahead of the consumers), so with stop_on_eos it is issued in chunks and the host looks at the tokens in between:
            // a generation that ends after 20 of 256 tokens costs one chunk, not 256 steps.
            std::vector<int> seen(n_new);
            while (steps_done < n_new - 1) {
                const int c = std::min(kEosChunk, n_new - 1 - steps_done);
                TRY(run_mega(m, 0, c, 0));
                steps_done += c;
                if (steps_done >= n_new - 1) break;
                CK(cudaMemcpyAsync(seen.data(), m.out_tokens.p, (size_t)(1 + steps_done) * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
                CK(cudaStreamSynchronize(g_stream));
                if (std::find(seen.begin(), seen.begin() + 1 + steps_done, 2) != seen.begin() + 1 + steps_done) break;
            }
        } else {
            TRY(run_mega(m, 0, n_new - 1, 0));
            steps_done = std::max(0, n_new - 1);
        }
        CK(cudaEventRecord(e1, g_stream));
    } else {
    // prefill: the prompt goes through the same incremental step, one token at a time; only the last one needs logits
    for (int i = 0; i < n_prompt; ++i) {
        set_token_kernel<<<1, 1, 0, g_stream>>>(m.state.p, m.prompt.p, i);
        ++g_launches;
        const bool last = i == n_prompt - 1;
        TRY(run_step(m, last && n_new > 0));
    }
    CK(cudaEventRecord(e0, g_stream));
    // the last prompt step already produced token 0; every further step feeds the token the previous one picked
    for (int i = 1; i < n_new; ++i) TRY(run_step(m, true));
    steps_done = std::max(0, n_new - 1);
    CK(cudaEventRecord(e1, g_stream));
    }
    m.host_pos = n_prompt + steps_done;
    const int have = n_new > 0 ? 1 + steps_done : 0;   // tokens on the device (fewer than n_new after an early EOS)
    std::vector<int> toks(std::max(n_new, 1), 0);
    if (have > 0) CK(cudaMemcpyAsync(toks.data(), m.out_tokens.p, have * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    if (logits_host && have > 0) {
        TRY(gather_sharded_logits(m, m.hist.p, (size_t)have * V));
        CK(cudaMemcpyAsync(logits_host, m.hist.p, (size_t)have * V * 4, cudaMemcpyDeviceToHost, g_stream));
    }
    CK(cudaStreamSynchronize(g_stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaEventElapsedTime(&m.last_prefill_ms, ep, e0));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(ep);
    if (decode_ms) *decode_ms = ms;
    int produced = have;
    if (stop_on_eos)
        for (int i = 0; i < have; ++i)
            if (toks[i] == 2) { produced = i + 1; break; }  // hard-coded EOS id 2 (:760)
    for (int i = 0; i < produced; ++i) out_tokens[i] = toks[i];
    if (n_out) *n_out = produced;
    return 0;
}

// InferenceEngine::generate_batch (src/model/inference_engine.cpp:804-828): `batch` prompts of equal length, n_new greedy
// tokens each.  The reference loops over generate(); here the sequences advance in lockstep so that the weights are read
// once per step for all of them (tensor-core GEMM path).  out_tokens: [batch][n_new]; logits_last (optional): [batch][vocab]
// of the last step; decode_ms (optional): CUDA-event time of the n_new - 1 decode steps after the prompt.
// every kernel of the lockstep step asks for the same L1 / shared-memory split as the GEMM (which needs nearly all of it): a
// kernel that wants a different carve-out than its predecessor makes the SMs reconfigure before it can start
static int batch_carveout() {
    if (g_batch_carveout_done) return 0;
    const void* fns[] = {(const void*)gemm_i8_tc_small_kernel, (const void*)gemm_i8_tc_kernel, (const void*)rmsnorm_digits_small_kernel,
                         (const void*)rmsnorm_digits_kernel, (const void*)rope_kv_batch_kernel, (const void*)attn_partial_kernel,
                         (const void*)attn_combine_kernel, (const void*)argmax_rows_kernel, (const void*)batch_advance_kernel,
                         (const void*)embed_rows_kernel, (const void*)swiglu_rows_kernel, (const void*)relu_rows_kernel, (const void*)batch_feed_kernel,
                         (const void*)beam_expand_kernel<32>, (const void*)beam_expand_kernel<0>, (const void*)kv_pages_copy_kernel};
    for (const void* f : fns) CK(cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    g_batch_carveout_done = true;
    return 0;
}

static int generate_batch_impl(ti_model_t h, const int32_t* prompts, const int32_t* lens, int32_t batch, int32_t max_len, int32_t n_new,
                               int32_t* out_tokens, float* logits_last, float* decode_ms) {
    TRY(need_init());
    Model* mp = get_model(h);
    if (!mp || !mp->finalized) return fail("invalid or unfinalized model handle");
    Model& m = *mp;
    if (batch <= 0) return fail("batch must be >= 1");
    if (max_len <= 0) return fail("Input tokens cannot be empty");  // validate_input_tokens (:1409)
    if (n_new <= 0) return fail("n_new must be >= 1");
    if (!batch_eligible(m)) return fail("generate_batch needs a complete, non-literal model (q/k/v/o, up/down, lm_head; rope per head or off)");
    const int B = batch, V = m.cfg.vocab, H = m.cfg.hidden;
    int min_len = max_len;
    for (int b = 0; b < B; ++b) {
        if (lens[b] <= 0 || lens[b] > max_len) return fail("Input tokens cannot be empty");
        min_len = std::min(min_len, (int)lens[b]);
        for (int p = 0; p < lens[b]; ++p)
            if (prompts[(size_t)b * max_len + p] < 0 || prompts[(size_t)b * max_len + p] >= V) return fail("token id %d out of range", prompts[(size_t)b * max_len + p]);
    }
    const int total = max_len + n_new - 1;   // the longest prompt decides how many lockstep steps run
    if (total > m.cfg.max_seq) return fail("KV cache overflow: sequence too long");  // :100-102
    TRY(batch_carveout());
    // everything that allocates or launches set-up kernels happens before the step graphs are captured
    for (auto& ly : m.layers)
        for (QWeight* w : {ly.qkv.get(), ly.o.get(), ly.gateup.get(), ly.down.get()}) TRY(ensure_kmajor(*w));
    TRY(ensure_kmajor(*m.lm_head));
    TRY(ensure_pf_scratch(m, B));
    const int pages = (total + m.page_tokens - 1) / m.page_tokens;
    if (!m.batch || m.batch->B != B || m.batch->pages_per_seq < pages) {
        m.batch.reset(new BatchState());
        BatchState& bs = *m.batch;
        bs.B = B;
        bs.pages_per_seq = pages;
        bs.max_splits = std::max(1, std::min(m.max_splits, (2 * g_num_sms) / std::max(1, B * m.attn_heads)));
        bs.k.resize(m.layers.size());
        bs.v.resize(m.layers.size());
        const size_t pool = (size_t)B * pages * m.page_tokens * (H / m.tp);   // a rank keeps the K / V rows of its own heads
        if (m.tp > 1) TRY(bs.ar.alloc((size_t)B * H));
        for (size_t l = 0; l < m.layers.size(); ++l) { TRY(bs.k[l].alloc(pool)); TRY(bs.v[l].alloc(pool)); }
        std::vector<int> tab((size_t)B * pages);
        for (int b = 0; b < B; ++b)
            for (int p = 0; p < pages; ++p) tab[(size_t)b * pages + p] = b * pages + (pages - 1 - p);   // deliberately not the identity
        TRY(bs.tables.alloc(tab.size()));
        CK(cudaMemcpyAsync(bs.tables.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        TRY(bs.tokens.alloc(B));
        TRY(bs.lens.alloc(B));
        TRY(bs.pos_step.alloc(2));
        TRY(bs.part_o.alloc((size_t)B * m.attn_heads * bs.max_splits * m.attn_dim));
        TRY(bs.part_ml.alloc((size_t)B * m.attn_heads * bs.max_splits * 2));
        TRY(bs.logits.alloc((size_t)B * V));
    }
    BatchState& bs = *m.batch;
    bool regraph = false;
    if (bs.out.n < (size_t)B * n_new) {
        TRY(bs.out.alloc((size_t)B * n_new));
        regraph = true;   // the graphs hold the old pointer
    }
    std::vector<int> cols((size_t)max_len * B, 0);
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < lens[b]; ++p) cols[(size_t)p * B + b] = prompts[(size_t)b * max_len + p];
    if (bs.prompts.n < cols.size()) {
        TRY(bs.prompts.alloc(cols.size()));
        regraph = true;
    }
    if (regraph || bs.cap_scratch != m.pf_gen || bs.cap_stride != n_new) {   // kernel parameters of the captured graphs
        bs.drop_graphs();
        bs.cap_scratch = m.pf_gen;
        bs.cap_stride = n_new;
    }
    CK(cudaMemcpyAsync(bs.prompts.p, cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(bs.lens.p, lens, B * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemsetAsync(bs.pos_step.p, 0, 2 * sizeof(int), g_stream));   // reset(): every cache is empty again
    CK(cudaMemsetAsync(bs.out.p, 0, (size_t)B * n_new * sizeof(int), g_stream));
    auto run = [&](bool sample) -> int {
        cudaGraphExec_t& ge = bs.graph[sample ? 1 : 0];
        if (!ge) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
            const int rc = batch_step(m, bs, sample, n_new);
            const cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
            if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail("cudaStreamEndCapture: %s", cudaGetErrorString(e));
            CK(cudaGraphInstantiate(&ge, graph, 0));
            cudaGraphDestroy(graph);
        }
        CK(cudaGraphLaunch(ge, g_stream));
        return 0;
    };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    // steps 0 .. min_len - 2 are prompt-only for every sequence (no lm_head); from step min_len - 1 on some sequence samples
    for (int s = 0; s < max_len; ++s) TRY(run(s >= min_len - 1));
    CK(cudaEventRecord(e0, g_stream));
    for (int i = 1; i < n_new; ++i) TRY(run(true));
    CK(cudaEventRecord(e1, g_stream));
    CK(cudaMemcpyAsync(out_tokens, bs.out.p, (size_t)B * n_new * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    if (logits_last) CK(cudaMemcpyAsync(logits_last, bs.logits.p, (size_t)B * V * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (decode_ms) *decode_ms = ms;
    return 0;
}

int ti_b200_generate_batch_greedy(ti_model_t h, const int32_t* prompts, int32_t batch, int32_t n_prompt, int32_t n_new, int32_t* out_tokens,
                                  float* logits_last, float* decode_ms) {
    if (batch <= 0) return fail("batch must be >= 1");
    std::vector<int32_t> lens(batch, n_prompt);
    return generate_batch_impl(h, prompts, lens.data(), batch, n_prompt, n_new, out_tokens, logits_last, decode_ms);
}

// generate_batch with prompts of DIFFERENT lengths (the reference loops over generate(), :804-828, so any lengths go): prompts is
// [batch][max_len] (row b holds lens[b] tokens, the rest is ignored).  The sequences are left-aligned and advance in lockstep;
// a sequence whose prompt has ended feeds its own picks while longer prompts are still being read.  out_tokens [batch][n_new].
// (logits_last of the equal-length entry has no ragged counterpart: the sequences' last steps differ.)
int ti_b200_generate_batch_ragged(ti_model_t h, const int32_t* prompts, const int32_t* lens, int32_t batch, int32_t max_len, int32_t n_new,
                                  int32_t* out_tokens, float* decode_ms) {
    if (batch <= 0 || !lens) return fail("batch must be >= 1");
    return generate_batch_impl(h, prompts, lens, batch, max_len, n_new, out_tokens, nullptr, decode_ms);
}

// generate_beam_search (:830-871) / beam_search_decode (:1912-2069) on the cached, lockstep engine.  The reference runs a full
// forward_pass over every candidate's whole sequence at every step (:1961, no cache); here a candidate is a row of the batched
// step and owns a page TABLE into one shared K/V pool: the prompt is cached once, a fork copies the table and bumps the pages'
// reference counts, and only a partly filled last page is copied when two beams are about to append to it.  The expansion
// (softmax, top-k / top-p on probabilities, the beam best tokens) runs on the device next to the logits; the host keeps the
// reference's bookkeeping (cumulative log-probability, length-normalised score, keep the beam best, early stop) on
// beam * beam pairs per step.  Where the reference's order is unspecified (ties in std::sort / the heap) the earlier candidate wins.
// out_tokens [beam_size][max_new] (new tokens only, like GenerationResult of :849-857), results best first.
int ti_b200_beam_search(ti_model_t h, const int32_t* prompt, int32_t n_prompt, int32_t max_new, int32_t beam_size, float temperature, int32_t top_k,
                        float top_p, float length_penalty, int32_t eos_token, int32_t* out_tokens, int32_t* out_lens, float* out_logprob,
                        float* out_score, int32_t* out_finished, int32_t* n_results) {
    TRY(need_init());
    Model* mp = get_model(h);
    if (!mp || !mp->finalized) return fail("invalid or unfinalized model handle");
    Model& m = *mp;
    if (beam_size <= 0) return fail("Beam size must be greater than 0");   // :836-838
    if (beam_size > kBeamMax) return fail("beam_size above %d is not supported", kBeamMax);
    if (n_prompt <= 0) return fail("Input tokens cannot be empty");
    if (max_new < 0) return fail("max_new must be >= 0");
    if (!(temperature > 0.0f)) return fail("Temperature must be positive");
    if (!out_tokens || !out_lens || !n_results) return fail("null argument");
    if (m.tp > 1) return fail("beam search runs on a single-GPU model");
    if (!batch_eligible(m)) return fail("beam search needs a complete, non-literal model (q/k/v/o, up/down, lm_head; rope per head or off)");
    const int B = beam_size, V = m.cfg.vocab, H = m.cfg.hidden, pt = m.page_tokens;
    for (int i = 0; i < n_prompt; ++i)
        if (prompt[i] < 0 || prompt[i] >= V) return fail("token id %d out of range", prompt[i]);
    struct Cand { std::vector<int> toks; float log_prob = 0.f, score = 0.f; bool finished = false; int row = -1; };
    auto emit = [&](std::vector<Cand>& done) {
        std::stable_sort(done.begin(), done.end(), [](const Cand& a, const Cand& b) { return a.score > b.score; });   // :2060-2063
        const int n = (int)std::min<size_t>(done.size(), (size_t)B);
        for (int i = 0; i < n; ++i) {
            out_lens[i] = (int)done[i].toks.size();
            for (size_t t = 0; t < done[i].toks.size(); ++t) out_tokens[(size_t)i * max_new + t] = done[i].toks[t];
            if (out_logprob) out_logprob[i] = done[i].log_prob;
            if (out_score) out_score[i] = done[i].score;
            if (out_finished) out_finished[i] = done[i].finished ? 1 : 0;
        }
        *n_results = n;
    };
    if (max_new == 0) {   // the loop of :1938 does not run: the initial candidate comes back as the only result
        std::vector<Cand> done(1);
        done[0].finished = true;
        emit(done);
        return 0;
    }
    const int total = n_prompt + max_new - 1;   // the last token of a sequence is never fed back
    if (total > m.cfg.max_seq) return fail("KV cache overflow: sequence too long");
    TRY(batch_carveout());
    for (auto& ly : m.layers)
        for (QWeight* w : {ly.qkv.get(), ly.o.get(), ly.gateup.get(), ly.down.get()}) TRY(ensure_kmajor(*w));
    TRY(ensure_kmajor(*m.lm_head));
    TRY(ensure_pf_scratch(m, B));
    const int pages = (total + pt - 1) / pt;
    const size_t page_elems = (size_t)pt * H;
    if (!m.beams || m.beams->beam != B || m.beams->bs.pages_per_seq < pages) {
        m.beams.reset(new BeamState());
        BeamState& st = *m.beams;
        BatchState& bs = st.bs;
        st.beam = B;
        st.pool_pages = B * pages + B;   // worst case every beam owns all its pages; + one scratch page per idle row
        bs.B = B;
        bs.pages_per_seq = pages;
        bs.max_splits = std::max(1, std::min(m.max_splits, (2 * g_num_sms) / std::max(1, B * m.attn_heads)));
        bs.k.resize(m.layers.size());
        bs.v.resize(m.layers.size());
        std::vector<float*> kp(m.layers.size()), vp(m.layers.size());
        for (size_t l = 0; l < m.layers.size(); ++l) {
            TRY(bs.k[l].alloc((size_t)st.pool_pages * page_elems));
            TRY(bs.v[l].alloc((size_t)st.pool_pages * page_elems));
            // idle rows read their scratch page through the attention: keep it finite
            CK(cudaMemsetAsync(bs.k[l].p + (size_t)B * pages * page_elems, 0, (size_t)B * page_elems * sizeof(float), g_stream));
            CK(cudaMemsetAsync(bs.v[l].p + (size_t)B * pages * page_elems, 0, (size_t)B * page_elems * sizeof(float), g_stream));
            kp[l] = bs.k[l].p;
            vp[l] = bs.v[l].p;
        }
        TRY(st.k_ptrs.alloc(std::max<size_t>(kp.size(), 1)));
        TRY(st.v_ptrs.alloc(std::max<size_t>(vp.size(), 1)));
        CK(cudaMemcpyAsync(st.k_ptrs.p, kp.data(), kp.size() * sizeof(float*), cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemcpyAsync(st.v_ptrs.p, vp.data(), vp.size() * sizeof(float*), cudaMemcpyHostToDevice, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        TRY(bs.tables.alloc((size_t)B * pages));
        TRY(bs.tokens.alloc(B));
        TRY(bs.pos_step.alloc(2));
        TRY(bs.part_o.alloc((size_t)B * m.attn_heads * bs.max_splits * m.attn_dim));
        TRY(bs.part_ml.alloc((size_t)B * m.attn_heads * bs.max_splits * 2));
        TRY(bs.logits.alloc((size_t)B * V));
        TRY(st.cand_prob.alloc((size_t)B * B));
        TRY(st.cand_tok.alloc((size_t)B * B));
        TRY(st.cand_cnt.alloc(B));
        TRY(st.copy_pairs.alloc((size_t)2 * B));
    }
    BeamState& st = *m.beams;
    BatchState& bs = st.bs;
    const int ppr = bs.pages_per_seq;   // row stride of the page tables
    if (st.cap_scratch != m.pf_gen || st.cap_temperature != temperature || st.cap_top_k != top_k || st.cap_top_p != top_p) {
        st.drop_graphs();
        st.cap_scratch = m.pf_gen;
        st.cap_temperature = temperature;
        st.cap_top_k = top_k;
        st.cap_top_p = top_p;
    }
    CK(cudaMemsetAsync(bs.pos_step.p, 0, 2 * sizeof(int), g_stream));
    auto run = [&](bool expand) -> int {
        cudaGraphExec_t& ge = st.graph[expand ? 1 : 0];
        if (!ge) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
            int rc = batch_forward(m, bs, expand);
            if (rc == 0 && expand) {
                BeamExpandArgs ea{};
                ea.logits = bs.logits.p;
                ea.V = V;
                ea.ld = V;
                ea.temperature = temperature;
                ea.top_k = top_k;
                ea.top_p = top_p;
                ea.beam = B;
                ea.cand_prob = st.cand_prob.p;
                ea.cand_tok = st.cand_tok.p;
                ea.cand_cnt = st.cand_cnt.p;
                ea.staged = sample_row_smem_bytes(V) ? 1 : 0;
                if (sample_keys_in_registers(V)) beam_expand_kernel<32><<<B, kSampleThreads, sample_row_smem_bytes(V), g_stream>>>(ea);
                else beam_expand_kernel<0><<<B, kSampleThreads, sample_row_smem_bytes(V), g_stream>>>(ea);
                ++g_launches;
            }
            batch_advance_kernel<<<1, 1, 0, g_stream>>>(bs.pos_step.p, 0);
            ++g_launches;
            const cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
            if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail("cudaStreamEndCapture: %s", cudaGetErrorString(e));
            CK(cudaGraphInstantiate(&ge, graph, 0));
            cudaGraphDestroy(graph);
        }
        CK(cudaGraphLaunch(ge, g_stream));
        return 0;
    };
    // ---- page bookkeeping (host): reference counts, a free list, one table per row ----
    const int data_pages = B * pages;
    std::vector<int> ref(data_pages, 0), free_pages;
    for (int p = data_pages - 1; p >= 0; --p) free_pages.push_back(p);
    std::vector<int> tab((size_t)B * ppr);
    for (int r = 0; r < B; ++r)
        for (int j = 0; j < ppr; ++j) tab[(size_t)r * ppr + j] = data_pages + r;   // idle: everything on the row's scratch page
    std::vector<int> pairs;
    auto alloc_page = [&]() { const int p = free_pages.back(); free_pages.pop_back(); ref[p] = 1; return p; };
    auto unref = [&](int p) { if (p < data_pages && --ref[p] == 0) free_pages.push_back(p); };
    // every active row must own the page position t falls into before the step appends to it
    auto prepare_append = [&](int nrows, int t) {
        const int j = t / pt;
        for (int r = 0; r < nrows; ++r) {
            int& e = tab[(size_t)r * ppr + j];
            if (t % pt == 0) {
                e = alloc_page();
            } else if (ref[e] > 1) {   // shared with another beam: copy the filled part, then append privately
                const int fresh = alloc_page();
                pairs.push_back(e);
                pairs.push_back(fresh);
                --ref[e];
                e = fresh;
            }
        }
    };
    auto push_step = [&](const std::vector<int>& feed, int t, bool expand) -> int {
        CK(cudaMemcpyAsync(bs.tables.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, g_stream));
        if (!pairs.empty()) {
            CK(cudaMemcpyAsync(st.copy_pairs.p, pairs.data(), pairs.size() * sizeof(int), cudaMemcpyHostToDevice, g_stream));
            kv_pages_copy_kernel<<<dim3((unsigned)(pairs.size() / 2), (unsigned)std::max<size_t>(m.layers.size(), 1), 2), 256, 0, g_stream>>>(
                st.k_ptrs.p, st.v_ptrs.p, st.copy_pairs.p, page_elems, (size_t)(t % pt) * H);
            ++g_launches;
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(bs.tokens.p, feed.data(), B * sizeof(int), cudaMemcpyHostToDevice, g_stream));
        TRY(run(expand));
        CK(cudaStreamSynchronize(g_stream));   // tab / pairs / feed are reused by the next step
        pairs.clear();
        return 0;
    };
    // ---- the prompt: cached ONCE, by row 0 (the other rows idle on their scratch pages) ----
    std::vector<int> feed(B, 0);
    for (int t = 0; t < n_prompt; ++t) {
        prepare_append(1, t);
        std::fill(feed.begin(), feed.end(), prompt[t]);
        TRY(push_step(feed, t, t == n_prompt - 1));
    }
    std::vector<Cand> active(1), done;
    active[0].row = 0;
    std::vector<float> cprob((size_t)B * B);
    std::vector<int> ctok((size_t)B * B), ccnt(B);
    for (int step = 0; step < max_new && !active.empty(); ++step) {
        CK(cudaMemcpyAsync(cprob.data(), st.cand_prob.p, cprob.size() * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
        CK(cudaMemcpyAsync(ctok.data(), st.cand_tok.p, ctok.size() * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
        CK(cudaMemcpyAsync(ccnt.data(), st.cand_cnt.p, B * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        std::vector<Cand> next;
        for (const Cand& c : active) {   // most probable candidate first (the heap order of :1941-1944)
            const int cnt = std::min(ccnt[c.row], B);
            for (int i = 0; i < cnt; ++i) {
                const float p = cprob[(size_t)c.row * B + i];
                if (!(p > 0.0f)) continue;
                Cand nc = c;
                nc.toks.push_back(ctok[(size_t)c.row * B + i]);
                nc.log_prob = c.log_prob + std::log(p);
                nc.finished = nc.toks.back() == eos_token || (int)nc.toks.size() >= max_new;   // :2015-2016
                const float penalty = std::pow((float)(n_prompt + (int)nc.toks.size()), length_penalty);   // :2024-2026
                nc.score = nc.log_prob / penalty;
                next.push_back(std::move(nc));
            }
        }
        std::stable_sort(next.begin(), next.end(), [](const Cand& a, const Cand& b) { return a.score > b.score; });   // :2030-2033
        std::vector<Cand> kept;
        for (size_t i = 0; i < next.size() && i < (size_t)B; ++i) {
            if (next[i].finished) done.push_back(std::move(next[i]));
            else kept.push_back(std::move(next[i]));
        }
        std::stable_sort(kept.begin(), kept.end(), [](const Cand& a, const Cand& b) { return a.log_prob > b.log_prob; });
        if (done.size() >= (size_t)B || kept.empty()) { active = std::move(kept); break; }   // :2046-2048
        // ---- fork: row r of the next step continues kept[r]; its table is its parent's ----
        const int t = n_prompt + step;   // tokens in every active cache; the step below appends position t
        const int used = (t + pt - 1) / pt;
        std::vector<int> ntab(tab.size());
        for (int r = 0; r < B; ++r)
            for (int j = 0; j < ppr; ++j) ntab[(size_t)r * ppr + j] = data_pages + r;
        for (size_t r = 0; r < kept.size(); ++r)
            for (int j = 0; j < used; ++j) {
                const int p = tab[(size_t)kept[r].row * ppr + j];
                ntab[r * ppr + j] = p;
                ++ref[p];
            }
        for (const Cand& c : active)
            for (int j = 0; j < used; ++j) unref(tab[(size_t)c.row * ppr + j]);
        tab.swap(ntab);
        std::fill(feed.begin(), feed.end(), 0);
        for (size_t r = 0; r < kept.size(); ++r) { kept[r].row = (int)r; feed[r] = kept[r].toks.back(); }
        active = std::move(kept);
        prepare_append((int)active.size(), t);
        TRY(push_step(feed, t, true));
    }
    for (Cand& c : active) { c.finished = true; done.push_back(std::move(c)); }   // :2051-2057
    emit(done);
    return 0;
}

int ti_b200_beam_expand(const float* logits_host, size_t rows, size_t vocab, float temperature, int32_t top_k, float top_p, int32_t beam_size,
                        float* probs_out, int32_t* tokens_out, int32_t* counts_out) {
    TRY(need_init());
    if (rows == 0 || vocab == 0) return fail("empty logits");
    if (beam_size <= 0 || beam_size > kBeamMax) return fail("beam_size must be in [1, %d]", kBeamMax);
    if (!(temperature > 0.0f)) return fail("Temperature must be positive");
    DevBuf<float> lg, pr;
    DevBuf<int> tk, cn;
    TRY(upload(lg, logits_host, rows * vocab));
    TRY(pr.alloc(rows * beam_size));
    TRY(tk.alloc(rows * beam_size));
    TRY(cn.alloc(rows));
    BeamExpandArgs ea{};
    ea.logits = lg.p;
    ea.V = (int)vocab;
    ea.ld = (int)vocab;
    ea.temperature = temperature;
    ea.top_k = top_k;
    ea.top_p = top_p;
    ea.beam = beam_size;
    ea.cand_prob = pr.p;
    ea.cand_tok = tk.p;
    ea.cand_cnt = cn.p;
    ea.staged = sample_row_smem_bytes((int)vocab) ? 1 : 0;
    if (sample_keys_in_registers((int)vocab)) beam_expand_kernel<32><<<(unsigned)rows, kSampleThreads, sample_row_smem_bytes((int)vocab), g_stream>>>(ea);
    else beam_expand_kernel<0><<<(unsigned)rows, kSampleThreads, sample_row_smem_bytes((int)vocab), g_stream>>>(ea);
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(probs_out, pr.p, rows * beam_size * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaMemcpyAsync(tokens_out, tk.p, rows * beam_size * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaMemcpyAsync(counts_out, cn.p, rows * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int ti_b200_sample_logits(const float* logits_host, size_t rows, size_t vocab, float temperature, int32_t top_k, float top_p, uint64_t seed,
                          int32_t step, int32_t* tokens_out, float* logprobs_out) {
    TRY(need_init());
    if (rows == 0 || vocab == 0) return fail("Cannot sample from empty logits");   // :1555-1557
    if (temperature <= 0.0f) return fail("Temperature must be positive");          // apply_temperature (:1676-1678)
    DevBuf<float> lg, lp;
    DevBuf<int> tok;
    TRY(upload(lg, logits_host, rows * vocab));
    TRY(tok.alloc(rows));
    TRY(lp.alloc(rows));
    TRY(launch_sample(lg.p, (int)rows, (int)vocab, (int)vocab, temperature, top_k, top_p, seed, nullptr, step, tok.p, lp.p, nullptr, nullptr, 0, nullptr));
    CK(cudaMemcpyAsync(tokens_out, tok.p, rows * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    if (logprobs_out) CK(cudaMemcpyAsync(logprobs_out, lp.p, rows * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

// generate() with sampling (:734-802 + :1554-1673): the token is picked ON THE DEVICE from the logits of every step
// (sample_kernel) and fed to the next forward pass without touching the host; one launch of the decode engine per token.
int ti_b200_generate_sampled(ti_model_t h, const int32_t* prompt, int32_t n_prompt, int32_t n_new, float temperature, int32_t top_k, float top_p,
                             uint64_t seed, int32_t stop_on_eos, int32_t* out_tokens, int32_t* n_out, float* logprobs_out, float* decode_ms) {
    TRY(need_init());
    Model* mp = get_model(h);
    if (!mp || !mp->finalized) return fail("invalid or unfinalized model handle");
    Model& m = *mp;
    if (m.cfg.compat_literal) return fail("sampling is not part of the literal benchmark path (top_k = 1 there)");
    if (n_prompt <= 0) return fail("Input tokens cannot be empty");
    if (n_new < 0) return fail("n_new must be >= 0");
    if (temperature <= 0.0f) return fail("Temperature must be positive");
    for (int i = 0; i < n_prompt; ++i)
        if (prompt[i] < 0 || prompt[i] >= m.cfg.vocab) return fail("token id %d out of range", prompt[i]);
    TRY(ti_b200_model_reset(h));
    TRY(check_capacity(m, n_prompt + std::max(0, n_new - 1)));
    const int V = m.cfg.vocab;
    if ((int)m.prompt.n < n_prompt) TRY(m.prompt.alloc(n_prompt));
    CK(cudaMemcpyAsync(m.prompt.p, prompt, n_prompt * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    const int cap = std::max(n_new, 1);
    if ((int)m.smp_tokens.n < cap) { TRY(m.smp_tokens.alloc(cap)); TRY(m.smp_logprobs.alloc(cap)); }
    if (!m.smp_count.p) TRY(m.smp_count.alloc(1));
    StepIO io{};   // the engines' own greedy bookkeeping stays off: the sampler owns the token history
    CK(cudaMemcpyAsync(m.io.p, &io, sizeof(io), cudaMemcpyHostToDevice, g_stream));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    TRY(prompt_pass(m, prompt, n_prompt));
    CK(cudaEventRecord(e0, g_stream));
    if (m.lm_sharded && m.hist.n < (size_t)V) TRY(m.hist.alloc((size_t)V));
    int sampled = 0, passes = 0;   // tokens picked / decode passes run after the prompt
    std::vector<int> seen;
    for (int i = 0; i < n_new; ++i) {
        const float* row = m.logits.p;
        if (m.lm_sharded) {   // every rank samples the same full row with the same uniform: identical tokens on all ranks
            CK(cudaMemcpyAsync(m.hist.p, m.logits.p, (size_t)V * 4, cudaMemcpyDeviceToDevice, g_stream));
            TRY(gather_sharded_logits(m, m.hist.p, (size_t)V));
            row = m.hist.p;
        }
        TRY(launch_sample(row, 1, V, V, temperature, top_k, top_p, seed, nullptr, i, m.smp_tokens.p + i, m.smp_logprobs.p + i, nullptr, nullptr, 0,
                          &m.state.p->token));
        ++sampled;
        if (stop_on_eos && sampled % kEosChunk == 0 && i + 1 < n_new) {   // look at the tokens so far: stop issuing steps after an EOS (:760)
            seen.resize(sampled);
            CK(cudaMemcpyAsync(seen.data(), m.smp_tokens.p, (size_t)sampled * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            if (std::find(seen.begin(), seen.end(), 2) != seen.end()) break;
        }
        if (i + 1 < n_new) { TRY(decode_pass(m)); ++passes; }
    }
    CK(cudaEventRecord(e1, g_stream));
    m.host_pos = n_prompt + passes;
    std::vector<int> toks(cap);
    std::vector<float> lps(cap);
    if (sampled > 0) {
        CK(cudaMemcpyAsync(toks.data(), m.smp_tokens.p, sampled * sizeof(int), cudaMemcpyDeviceToHost, g_stream));
        CK(cudaMemcpyAsync(lps.data(), m.smp_logprobs.p, sampled * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    }
    CK(cudaStreamSynchronize(g_stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (decode_ms) *decode_ms = ms;
    int produced = sampled;
    if (stop_on_eos)
        for (int i = 0; i < sampled; ++i)
            if (toks[i] == 2) { produced = i + 1; break; }  // :760
    for (int i = 0; i < produced; ++i) {
        out_tokens[i] = toks[i];
        if (logprobs_out) logprobs_out[i] = lps[i];
    }
    if (n_out) *n_out = produced;
    return 0;
}

// compute_logprobs (:873-954): forward pass over `tokens`, out[pos] = log softmax(logits[pos])[tokens[pos]] (the reference
// scores a token under the logits of ITS OWN position); -20 for an id outside the vocabulary.
int ti_b200_compute_logprobs(ti_model_t h, const int32_t* tokens, int32_t n, float* out) {
    TRY(need_init());
    Model* mp = get_model(h);
    if (!mp || !mp->finalized) return fail("invalid or unfinalized model handle");
    Model& m = *mp;
    if (m.cfg.compat_literal) return fail("compute_logprobs is not available on the literal benchmark path");
    if (n <= 0) return fail("Input tokens cannot be empty");
    const int V = m.cfg.vocab;
    TRY(ti_b200_model_reset(h));
    TRY(check_capacity(m, n));
    std::vector<int> fed(n);
    for (int i = 0; i < n; ++i) fed[i] = (tokens[i] < 0 || tokens[i] >= V) ? 0 : tokens[i];   // an invalid id is scored -20; it still needs a row to feed
    if ((int)m.prompt.n < n) TRY(m.prompt.alloc(n));
    CK(cudaMemcpyAsync(m.prompt.p, fed.data(), n * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    if (m.hist.n < (size_t)n * V) TRY(m.hist.alloc((size_t)n * V));
    DevBuf<int> tok;
    DevBuf<float> lp;
    TRY(tok.alloc(n));
    TRY(lp.alloc(n));
    CK(cudaMemcpyAsync(tok.p, tokens, n * sizeof(int), cudaMemcpyHostToDevice, g_stream));
    StepIO io{};
    io.hist = m.hist.p;
    io.hist_cap = n;
    CK(cudaMemcpyAsync(m.io.p, &io, sizeof(io), cudaMemcpyHostToDevice, g_stream));
    if (m.use_mega) {
        TRY(run_mega(m, n, n, 0));   // every step runs the lm_head and keeps its logits
    } else {
        for (int i = 0; i < n; ++i) {
            set_token_kernel<<<1, 1, 0, g_stream>>>(m.state.p, m.prompt.p, i);
            ++g_launches;
            TRY(run_step(m, true));
        }
    }
    TRY(gather_sharded_logits(m, m.hist.p, (size_t)n * V));
    logprob_rows_kernel<<<n, 256, 0, g_stream>>>(m.hist.p, V, tok.p, lp.p);
    ++g_launches;
    CK(cudaGetLastError());
    m.host_pos = n;
    CK(cudaMemcpyAsync(out, lp.p, n * sizeof(float), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    return 0;
}

int ti_b200_model_last_prefill_ms(ti_model_t h, float* ms) {
    Model* m = get_model(h);
    if (!m || !m->finalized) return fail("invalid or unfinalized model handle");
    *ms = m->last_prefill_ms;
    return 0;
}

static int debug_timeline_run(ti_model_t h, int32_t token, int64_t* stamps, size_t cap, size_t* n_phases, size_t* n_ctas, bool all) {
    TRY(need_init());
    Model* m = get_model(h);
    if (!m || !m->finalized || !m->use_mega) return fail("timeline needs a finalized model on the persistent-kernel engine");
    TRY(check_capacity(*m, 1));
    const size_t per_cta = (size_t)m->nphases * kStampsPerPhase;
    const size_t n = per_cta * (size_t)g_num_sms;
    const size_t want = all ? n : per_cta;
    if (cap < want) return fail("stamp buffer too small: need %zu", want);
    TRY(m->dbg.alloc(n));
    CK(cudaMemsetAsync(m->dbg.p, 0, n * sizeof(long long), g_stream));
    StepIO io{};
    CK(cudaMemcpyAsync(m->io.p, &io, sizeof(io), cudaMemcpyHostToDevice, g_stream));
    CK(cudaMemcpyAsync(&m->state.p->token, &token, sizeof(int), cudaMemcpyHostToDevice, g_stream));
    m->dbg_on = true;
    m->dbg_all = all;
    int rc = run_mega(*m, 0, 1, 0);
    m->dbg_on = false;
    m->dbg_all = false;
    TRY(rc);
    m->host_pos += 1;
    CK(cudaMemcpyAsync(stamps, m->dbg.p, want * sizeof(long long), cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    *n_phases = m->nphases;
    if (n_ctas) *n_ctas = g_num_sms;
    return 0;
}
int ti_b200_debug_timeline(ti_model_t h, int32_t token, int64_t* stamps, size_t cap, size_t* n_phases) {
    return debug_timeline_run(h, token, stamps, cap, n_phases, nullptr, false);
}
int ti_b200_debug_timeline_all(ti_model_t h, int32_t token, int64_t* stamps, size_t cap, size_t* n_phases, size_t* n_ctas) {
    return debug_timeline_run(h, token, stamps, cap, n_phases, n_ctas, true);
}

int ti_b200_bench_gemv(const ti_qweight_t* ws, size_t n_w, size_t reps, float* ms) {
    TRY(need_init());
    if (n_w == 0 || reps == 0) return fail("nothing to time");
    std::vector<QWeight*> w(n_w);
    size_t maxK = 0, maxN = 0;
    for (size_t i = 0; i < n_w; ++i) {
        w[i] = get_qw(ws[i]);
        if (!w[i]) return fail("invalid qweight handle");
        maxK = std::max<size_t>(maxK, w[i]->L.K);
        maxN = std::max<size_t>(maxN, w[i]->L.N);
    }
    DevBuf<float> x, y;
    TRY(x.alloc(maxK));
    TRY(y.alloc(maxN));
    synth_fill_kernel<<<grid_for(maxK), 256, 0, g_stream>>>(x.p, maxK, 12345, 1.0f);
    ++g_launches;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (size_t i = 0; i < std::min<size_t>(n_w, 3); ++i) TRY(ti_b200_gemv_q_dev(ws[i], x.p, y.p));  // warm-up
    CK(cudaEventRecord(e0, g_stream));
    for (size_t r = 0; r < reps; ++r) TRY(ti_b200_gemv_q_dev(ws[r % n_w], x.p, y.p));
    CK(cudaEventRecord(e1, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    CK(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

int ti_b200_model_bench_gemv(ti_model_t h, int slot, size_t reps, float* ms, double* alg_bytes_per_launch) {
    TRY(need_init());
    Model* m = get_model(h);
    if (!m || !m->finalized) return fail("invalid or unfinalized model handle");
    std::vector<QWeight*> ws;
    if (slot == 4) ws.push_back(m->lm_head.get());
    else
        for (auto& ly : m->layers) {
            QWeight* w = slot == 0 ? ly.qkv.get() : slot == 1 ? ly.o.get() : slot == 2 ? ly.gateup.get() : ly.down.get();
            if (w) ws.push_back(w);
        }
    if (ws.empty() || reps == 0) return fail("nothing to time for slot %d", slot);
    size_t maxK = 0, maxN = 0;
    for (auto* w : ws) { maxK = std::max<size_t>(maxK, w->L.K); maxN = std::max<size_t>(maxN, w->L.N); }
    DevBuf<float> x, y;
    TRY(x.alloc(maxK));
    TRY(y.alloc(maxN));
    synth_fill_kernel<<<grid_for(maxK), 256, 0, g_stream>>>(x.p, maxK, 12345, 1.0f);
    ++g_launches;
    auto run = [&](QWeight* w) {
        GemvArgs a{};
        a.x = x.p;
        a.epi = EPI_STORE;
        a.out = y.p;
        return launch_gemv(*w, a, g_stream);
    };
    for (size_t i = 0; i < std::min<size_t>(ws.size(), 3); ++i) TRY(run(ws[i]));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, g_stream));
    for (size_t r = 0; r < reps; ++r) TRY(run(ws[r % ws.size()]));
    CK(cudaEventRecord(e1, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    CK(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const QWeight& w0 = *ws[0];
    if (alg_bytes_per_launch) *alg_bytes_per_launch = (double)w0.L.K * w0.L.N * w0.L.bits / 8.0 + 4.0 * w0.L.N + 4.0 * (w0.L.K + w0.L.N);
    return 0;
}

}  // extern "C"
