// test_host_api.cpp -- the C++ class surface of the B200 build, written the way the reference's own tests are
// (stand-alone main(), hand-rolled checks; tests/test_tensor.cpp, tests/test_quantization_complete.cpp,
// tests/test_inference_engine.cpp of the reference).  `cpu` mode needs no GPU: host value types and the loud failure
// of every device entry point.  `gpu` mode runs the ops, the quantizer fixtures and a small decoder on cuda:0.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

#include "turboinfer/core/tensor_engine.hpp"
#include "turboinfer/model/inference_engine.hpp"
#include "turboinfer/optimize/quantization.hpp"

using namespace turboinfer;
using core::DataType;
using core::Tensor;
using core::TensorShape;

static int g_failed = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_failed; } } while (0)
#define CHECK_THROWS(expr, type) do { bool ok_ = false; try { (void)(expr); } catch (const type&) { ok_ = true; } catch (...) {} \
    if (!ok_) { std::printf("FAIL %s:%d  %s did not throw %s\n", __FILE__, __LINE__, #expr, #type); ++g_failed; } } while (0)

static Tensor ramp(std::initializer_list<size_t> dims, float a, float b) {
    Tensor t{TensorShape(dims)};
    float* p = t.data_ptr<float>();
    for (size_t i = 0; i < t.shape().total_size(); ++i) p[i] = a * (float)(i % 97) + b;
    return t;
}

static void test_host_types() {
    Tensor t{TensorShape({2, 3, 4})};
    CHECK(t.shape().ndim() == 3 && t.shape().total_size() == 24 && t.byte_size() == 96 && !t.empty());
    for (size_t i = 0; i < 24; ++i) CHECK(t.data_ptr<float>()[i] == 0.0f);            // zero-initialised
    CHECK_THROWS(t.shape().size(3), std::out_of_range);
    CHECK_THROWS(t.data_ptr<int8_t>(), std::runtime_error);                             // sizeof check only
    (void)t.data_ptr<int32_t>();                                                        // same size: allowed, like the reference
    for (size_t i = 0; i < 24; ++i) t.data_ptr<float>()[i] = (float)i;
    Tensor r = t.reshape(TensorShape({6, 4}));
    r.data_ptr<float>()[0] = 99.f;
    CHECK(t.data_ptr<float>()[0] == 0.f);                                               // deep copy
    CHECK_THROWS(t.reshape(TensorShape({5, 5})), std::runtime_error);
    Tensor s = t.slice({0, 1, 1}, {2, 3, 3});
    CHECK(s.shape().dimensions() == (std::vector<size_t>{2, 2, 2}));
    CHECK(s.data_ptr<float>()[0] == 5.f && s.data_ptr<float>()[1] == 6.f && s.data_ptr<float>()[2] == 9.f && s.data_ptr<float>()[7] == 22.f);
    Tensor c = t;
    c.fill<float>(1.f);
    CHECK(t.data_ptr<float>()[5] == 5.f && c.data_ptr<float>()[5] == 1.f);
    CHECK(core::get_dtype_size(DataType::kInt8) == 1 && std::string(core::dtype_to_string(DataType::kInt32)) == "int32");
    Tensor q{TensorShape({4}), DataType::kInt8};
    CHECK(q.byte_size() == 4);
    CHECK(optimize::get_quantization_bits(optimize::QuantizationType::kInt4) == 4);
}

static void test_no_cpu_fallback(bool have_gpu) {
    CHECK_THROWS(core::TensorEngine(core::ComputeDevice::kCPU), std::runtime_error);
    if (!have_gpu) {
        CHECK_THROWS(core::TensorEngine(core::ComputeDevice::kAuto), std::runtime_error);  // no device: fail loudly
        optimize::Quantizer qz;
        CHECK_THROWS(qz.quantize_tensor(ramp({4, 4}, 0.1f, -1.f)), std::runtime_error);
        model::ModelData md;
        md.metadata().vocab_size = 8; md.metadata().hidden_size = 8; md.metadata().num_layers = 0; md.metadata().num_heads = 1;
        CHECK_THROWS(model::InferenceEngine(md), std::runtime_error);
    }
}

static void test_ops_gpu() {
    core::TensorEngine eng(core::ComputeDevice::kGPU);
    CHECK(eng.gpu_available());
    CHECK(eng.device_info().find("B200") != std::string::npos || !eng.device_info().empty());
    Tensor a = ramp({3, 40}, 0.01f, -0.3f), b = ramp({3, 40}, -0.02f, 0.5f);
    Tensor s = eng.add(a, b), m = eng.multiply(a, b), r = eng.relu(a), si = eng.silu(a), sc = eng.scale(a, 2.5f);
    for (size_t i = 0; i < 120; ++i) {
        const float x = a.data_ptr<float>()[i], y = b.data_ptr<float>()[i];
        CHECK(s.data_ptr<float>()[i] == x + y);
        CHECK(m.data_ptr<float>()[i] == x * y);
        CHECK(r.data_ptr<float>()[i] == std::fmax(0.f, x));
        CHECK(std::fabs(si.data_ptr<float>()[i] - x / (1.0f + std::exp(-x))) <= 1e-6f);
        CHECK(sc.data_ptr<float>()[i] == x * 2.5f);
    }
    CHECK_THROWS(eng.add(a, ramp({2, 40}, 1.f, 0.f)), std::runtime_error);
    // matmul [M,K].[K,N] and [B,T,K].[K,N]; integer weights are promoted without scale
    Tensor x = ramp({2, 3, 64}, 0.01f, -0.2f), w = ramp({64, 48}, -0.003f, 0.1f);
    Tensor y = eng.matmul(x, w);
    CHECK(y.shape().dimensions() == (std::vector<size_t>{2, 3, 48}));
    double worst = 0, scale = 0;
    for (size_t i = 0; i < 6; ++i)
        for (size_t n = 0; n < 48; ++n) {
            double acc = 0;
            for (size_t k = 0; k < 64; ++k) acc += (double)x.data_ptr<float>()[i * 64 + k] * w.data_ptr<float>()[k * 48 + n];
            worst = std::fmax(worst, std::fabs(acc - y.data_ptr<float>()[i * 48 + n]));
            scale = std::fmax(scale, std::fabs(acc));
        }
    CHECK(worst <= 1e-5 * scale);
    Tensor wi{TensorShape({64, 48}), DataType::kInt8};
    for (size_t i = 0; i < 64 * 48; ++i) wi.data_ptr<int8_t>()[i] = (int8_t)((int)(i % 15) - 7);
    Tensor yi = eng.matmul(x.reshape(TensorShape({6, 64})), wi);
    double acc0 = 0;
    for (size_t k = 0; k < 64; ++k) acc0 += (double)x.data_ptr<float>()[k] * (double)((int)((k * 48) % 15) - 7);
    CHECK(std::fabs(acc0 - yi.data_ptr<float>()[0]) <= 1e-4 * (1.0 + std::fabs(acc0)));
    CHECK_THROWS(eng.matmul(x, ramp({63, 48}, 1.f, 0.f)), std::runtime_error);
    // softmax rows sum to one; rms_norm of a constant row
    Tensor p = eng.softmax(ramp({4, 10}, 0.3f, -1.f));
    for (size_t rr = 0; rr < 4; ++rr) {
        float sum = 0;
        for (size_t i = 0; i < 10; ++i) sum += p.data_ptr<float>()[rr * 10 + i];
        CHECK(std::fabs(sum - 1.0f) <= 1e-5f);
    }
    Tensor ones{TensorShape({32})};
    ones.fill<float>(1.f);
    Tensor cst{TensorShape({2, 32})};
    cst.fill<float>(3.f);
    Tensor nrm = eng.rms_norm(cst, ones);
    CHECK(std::fabs(nrm.data_ptr<float>()[5] - 3.f / std::sqrt(9.f + 1e-5f)) <= 1e-6f);
    // attention over t identical keys: the output is the mean of the values = the value
    Tensor q = ramp({1, 1, 64}, 0.01f, 0.f), kk{TensorShape({1, 5, 64})}, vv{TensorShape({1, 5, 64})};
    for (size_t t = 0; t < 5; ++t)
        for (size_t h = 0; h < 64; ++h) { kk.data_ptr<float>()[t * 64 + h] = 0.1f; vv.data_ptr<float>()[t * 64 + h] = (float)h; }
    Tensor o1 = eng.attention_fast_incremental(q, kk, vv), o4 = eng.multi_head_attention(q, kk, vv, 4);
    for (size_t h = 0; h < 64; ++h) { CHECK(std::fabs(o1.data_ptr<float>()[h] - (float)h) <= 1e-4f); CHECK(std::fabs(o4.data_ptr<float>()[h] - (float)h) <= 1e-4f); }
    // RoPE at position 0 is the identity; pairs keep their norm at other positions
    Tensor xr = ramp({1, 2, 8}, 0.1f, 0.05f), pos{TensorShape({2})};
    pos.data_ptr<float>()[0] = 0.f; pos.data_ptr<float>()[1] = 3.f;
    Tensor ro = eng.apply_rope(xr, pos);
    for (size_t i = 0; i < 8; ++i) CHECK(std::fabs(ro.data_ptr<float>()[i] - xr.data_ptr<float>()[i]) <= 1e-6f);
    for (size_t i = 0; i < 4; ++i) {
        const float* a0 = xr.data_ptr<float>() + 8 + 2 * i; const float* b0 = ro.data_ptr<float>() + 8 + 2 * i;
        CHECK(std::fabs((a0[0] * a0[0] + a0[1] * a0[1]) - (b0[0] * b0[0] + b0[1] * b0[1])) <= 1e-5f);
    }
    CHECK_THROWS(eng.gelu(a), std::runtime_error);
}

static void test_quantizer_gpu() {
    // the reference's own fixtures (tests/test_quantization_complete.cpp:26-29, :88-91, :141-144): round trip error < scale
    struct Fx { optimize::QuantizationType type; bool sym; float lo, hi; int n; };
    const Fx fx[] = {{optimize::QuantizationType::kInt8, true, -10.f, 10.f, 16}, {optimize::QuantizationType::kInt4, true, -2.f, 2.f, 9},
                     {optimize::QuantizationType::kInt8, false, 1.f, 6.f, 6}};
    for (const Fx& f : fx) {
        Tensor x{TensorShape({(size_t)f.n})};
        for (int i = 0; i < f.n; ++i) x.data_ptr<float>()[i] = f.lo + (f.hi - f.lo) * (float)i / (float)(f.n - 1);
        optimize::QuantizationConfig cfg;
        cfg.type = f.type;
        cfg.symmetric = f.sym;
        optimize::Quantizer qz(cfg);
        optimize::QuantizationInfo info = qz.calculate_quantization_info(x);
        CHECK(info.scales.size() == 1 && info.scales[0] > 0.f);
        if (f.sym) CHECK(info.zero_points[0] == 0.f);
        Tensor q = qz.quantize_tensor(x);
        CHECK(q.dtype() == (f.type == optimize::QuantizationType::kInt8 ? DataType::kInt8 : DataType::kInt32));
        Tensor d = qz.dequantize_tensor(q, info);
        if (f.sym) {
            for (int i = 0; i < f.n; ++i) CHECK(std::fabs(d.data_ptr<float>()[i] - x.data_ptr<float>()[i]) <= info.scales[0]);
        } else {
            // the reference's asymmetric INT8 maps [min, max] onto zp .. zp + 255 and then clamps to int8 (:662-674), so the
            // upper part of the range saturates at 127; the contract here is the reference's formula, not a small error
            const float sc = info.scales[0], zp = info.zero_points[0];
            for (int i = 0; i < f.n; ++i) {
                const float qe = std::fmax(-128.f, std::fmin(127.f, std::round(x.data_ptr<float>()[i] / sc + zp)));
                CHECK(q.data_ptr<int8_t>()[i] == (int8_t)qe);
                CHECK(d.data_ptr<float>()[i] == sc * (qe - zp));
            }
        }
        if (f.type == optimize::QuantizationType::kInt4 && f.sym) {   // scale = max|x| / 7, q = round-half-away(x / scale)
            CHECK(info.scales[0] == 2.0f / 7.0f);
            CHECK(q.data_ptr<int32_t>()[0] == -7 && q.data_ptr<int32_t>()[8] == 7 && q.data_ptr<int32_t>()[4] == 0);
        }
    }
}

static model::ModelData small_model(const char* quant) {
    model::ModelData md;
    auto& m = md.metadata();
    m.name = "cpp-test"; m.architecture = "llama"; m.vocab_size = 96; m.hidden_size = 64; m.num_layers = 2; m.num_heads = 4;
    m.intermediate_size = 128; m.rope_theta = 10000.f;
    m.extra_params["b200.quantization"] = quant;
    uint32_t st = 12345u;
    auto rnd = [&](float amp) { st = st * 1664525u + 1013904223u; return ((float)(st >> 8) / 8388608.0f - 1.0f) * amp; };
    auto mat = [&](size_t r, size_t c, float amp) {
        Tensor t{TensorShape({r, c})};
        for (size_t i = 0; i < r * c; ++i) t.data_ptr<float>()[i] = rnd(amp);
        return t;
    };
    auto vec1 = [&](size_t n) {
        Tensor t{TensorShape({n})};
        for (size_t i = 0; i < n; ++i) t.data_ptr<float>()[i] = 1.0f + rnd(0.1f);
        return t;
    };
    md.add_tensor("token_embeddings.weight", mat(96, 64, 0.1f));
    md.add_tensor("lm_head.weight", mat(64, 96, 0.125f));
    md.add_tensor("norm.weight", vec1(64));
    for (int l = 0; l < 2; ++l) {
        const std::string p = "layers." + std::to_string(l) + ".";
        for (const char* n : {"attention.q_proj.weight", "attention.k_proj.weight", "attention.v_proj.weight", "attention.o_proj.weight"})
            md.add_tensor(p + n, mat(64, 64, 0.125f));
        md.add_tensor(p + "feed_forward.w1.weight", mat(64, 128, 0.125f));
        md.add_tensor(p + "feed_forward.w3.weight", mat(64, 128, 0.125f));
        md.add_tensor(p + "feed_forward.w2.weight", mat(128, 64, 0.09f));
        md.add_tensor(p + "attention_norm.weight", vec1(64));
        md.add_tensor(p + "ffn_norm.weight", vec1(64));
    }
    md.add_tensor("some.unrelated.tensor", mat(2, 2, 1.f));   // ignored, like the reference ignores unknown names
    return md;
}

static void test_engine_gpu() {
    for (const char* quant : {"int8", "int4"}) {
        model::InferenceConfig cfg;
        cfg.top_k = 1;   // greedy
        cfg.max_sequence_length = 64;
        model::InferenceEngine eng(small_model(quant), cfg);
        const std::vector<int> prompt = {1, 15, 25, 35};   // benchmark_inference.cpp:331
        model::GenerationResult a = eng.generate(prompt, 12), b = eng.generate(prompt, 12, true);
        CHECK(a.tokens.size() <= prompt.size() + 12 && a.tokens.size() > prompt.size());
        CHECK(a.tokens == b.tokens);                                       // deterministic
        CHECK(b.logprobs.size() == b.tokens.size() - prompt.size());
        for (float lp : b.logprobs) CHECK(lp <= 0.f);
        CHECK(a.stop_reason == "max_new_tokens" || a.stop_reason == "eos_token");
        // the one-call device loop and the step-by-step path agree token for token
        eng.reset_state();
        std::vector<int> seq = prompt;
        Tensor lg = eng.forward_pass_incremental(prompt);
        for (size_t i = prompt.size(); i < a.tokens.size(); ++i) {
            const float* row = lg.data_ptr<float>() + (lg.shape().size(1) - 1) * 96;
            int best = 0;
            for (int v = 1; v < 96; ++v) if (row[v] > row[best]) best = v;
            CHECK(best == a.tokens[i]);
            seq.push_back(best);
            if (i + 1 < a.tokens.size()) lg = eng.forward_pass_incremental({best});
        }
        {   // forward_pass: the whole sequence from an empty cache, logits of every position; its last row picks generate()'s first token
            Tensor all = eng.forward_pass(prompt);
            CHECK(all.shape().ndim() == 3 && all.shape().size(0) == 1 && all.shape().size(1) == prompt.size() && all.shape().size(2) == 96);
            const float* row = all.data_ptr<float>() + (prompt.size() - 1) * 96;
            int best = 0;
            for (int v = 1; v < 96; ++v) if (row[v] > row[best]) best = v;
            CHECK(best == a.tokens[prompt.size()]);
            Tensor again = eng.forward_pass(prompt);                       // deterministic, and independent of what was cached before
            CHECK(std::memcmp(all.data(), again.data(), all.byte_size()) == 0);
            CHECK_THROWS(eng.forward_pass({}), std::runtime_error);
        }
        CHECK_THROWS(eng.generate({}, 4), std::runtime_error);
        CHECK_THROWS(eng.generate(std::vector<int>(65, 1), 4), std::runtime_error);
        model::GenerationResult longrun = eng.generate(prompt, 1000);      // stops at max_sequence_length (or EOS)
        CHECK(longrun.tokens.size() <= 64 && longrun.finished);
        {   // greedy batches of equal-length prompts advance in lockstep on the device: same tokens as one by one
            std::vector<std::vector<int>> prompts = {prompt, prompt, prompt};
            for (size_t b = 0; b < prompts.size(); ++b) for (int& t : prompts[b]) t = (t + 7 * (int)b) % 96;
            auto lock = eng.generate_batch(prompts, 5);
            CHECK(lock.size() == 3);
            for (size_t b = 0; b < prompts.size(); ++b) CHECK(lock[b].tokens == eng.generate(prompts[b], 5).tokens);
        }
        cfg.top_k = 5;
        cfg.temperature = 0.8f;
        eng.set_config(cfg);
        model::GenerationResult smp = eng.generate(prompt, 6, true);       // on-device sampler (temperature, top-k, top-p)
        CHECK(smp.tokens.size() > prompt.size() && smp.logprobs.size() == smp.tokens.size() - prompt.size());
        for (float lp : smp.logprobs) CHECK(lp <= 0.f);
        {   // a seeded sampler reproduces its generation; compute_logprobs scores a sequence position by position
            model::InferenceEngine e1(small_model(quant), cfg), e2(small_model(quant), cfg);
            e1.set_seed(99);
            e2.set_seed(99);
            CHECK(e1.generate(prompt, 10).tokens == e2.generate(prompt, 10).tokens);
            std::vector<float> lps = e1.compute_logprobs(prompt);
            CHECK(lps.size() == prompt.size());
            for (float lp : lps) CHECK(lp <= 0.f && lp > -30.f);
            std::vector<int> bad_tok = prompt;
            bad_tok[1] = 1000000;                                          // out of vocabulary: the reference's -20 sentinel (:933-936)
            CHECK(e1.compute_logprobs(bad_tok)[1] == -20.0f);
        }
        cfg.top_k = 1;
        eng.set_config(cfg);
        {   // greedy batches with prompts of DIFFERENT lengths also run in lockstep on the device
            const std::vector<std::vector<int>> ragged = {prompt, {3, 4}, {9, 8, 7, 6, 5, 4, 3}};
            auto rb = eng.generate_batch(ragged, 4);
            CHECK(rb.size() == 3);
            for (size_t b = 0; b < ragged.size(); ++b) CHECK(rb[b].tokens == eng.generate(ragged[b], 4).tokens);
        }
        {   // beam search (generate_beam_search :830-871): new tokens only, best first; beam 1 with top_k 1 walks the greedy path
            CHECK_THROWS(eng.generate_beam_search(prompt, 4, 0), std::runtime_error);
            model::InferenceConfig bc = cfg;
            bc.top_k = 1;
            bc.eos_token_id = -1;
            eng.set_config(bc);
            auto one = eng.generate_beam_search(prompt, 6, 1, true);
            model::GenerationResult greedy = eng.generate(prompt, 6);
            CHECK(one.size() == 1 && one[0].tokens.size() == 6 && one[0].finished);
            if (greedy.tokens.size() == prompt.size() + 6)   // (generate() itself still stops on token 2)
                CHECK(std::equal(one[0].tokens.begin(), one[0].tokens.end(), greedy.tokens.begin() + prompt.size()));
            CHECK(one[0].logprobs.size() == 6 && one[0].logprobs[0] == 0.0f);   // the only survivor of top_k = 1 has probability 1
            bc.top_k = 50;
            bc.top_p = 0.95f;
            eng.set_config(bc);
            auto beams = eng.generate_beam_search(prompt, 6, 4, true);
            CHECK(beams.size() == 4);
            for (const auto& r : beams) {
                CHECK(r.tokens.size() == 6 && r.finished && r.logprobs.size() == 6);
                for (float lp : r.logprobs) CHECK(lp <= 0.f && lp == r.logprobs[0]);
            }
            eng.set_config(cfg);
        }
        auto batch = eng.generate_batch({prompt, {3, 4}}, 3);
        CHECK(batch.size() == 2);
        CHECK(eng.memory_usage() > 0 && !eng.performance_stats().empty());
    }
    model::ModelData bad = small_model("int8");
    bad.add_tensor("layers.0.attention.q_proj.weight", Tensor{TensorShape({64, 64}), DataType::kInt8});
    CHECK_THROWS(model::InferenceEngine(bad), std::runtime_error);
}

// ---- .tinq persistence (reference tests/test_quantization_persistence.cpp) -------------------------------------------------
// A quantized model made on the host alone (no device): fixed integers + parameters, one float tensor, one int tensor without
// parameters.
static model::ModelData tinq_fixture(bool int4) {
    model::ModelData md;
    auto& m = md.metadata();
    m.name = "tinq-fixture"; m.architecture = "llama"; m.version = "7"; m.vocab_size = 321; m.hidden_size = 48; m.num_layers = 3;
    m.num_heads = 6; m.intermediate_size = 100; m.rope_theta = 12345.5f;
    const int type = int4 ? (int)optimize::QuantizationType::kInt4 : (int)optimize::QuantizationType::kInt8;
    auto qmat = [&](size_t r, size_t c, int salt) {
        Tensor t{TensorShape({r, c}), int4 ? DataType::kInt32 : DataType::kInt8};
        for (size_t i = 0; i < r * c; ++i) {
            const int v = (int)((i * 37 + salt * 11) % (int4 ? 15 : 255)) - (int4 ? 7 : 127);
            if (int4) t.data_ptr<int32_t>()[i] = v; else t.data_ptr<int8_t>()[i] = (int8_t)v;
        }
        return t;
    };
    md.add_tensor("layers.0.attention.q_proj.weight", qmat(48, 48, 1));
    md.set_quant_params("layers.0.attention.q_proj.weight", {0.00123f, 0.f, type});
    md.add_tensor("lm_head.weight", qmat(48, 321, 2));
    md.set_quant_params("lm_head.weight", {0.0456f, int4 ? 0.f : 3.f, type});
    md.add_tensor("plain.integers", qmat(5, 7, 3));                       // no parameters: stays a plain integer tensor
    md.add_tensor("norm.weight", ramp({48}, 0.01f, 1.f));                 // float32 tensors travel untouched
    return md;
}
static bool same_tensor(const Tensor& a, const Tensor& b) {
    if (a.dtype() != b.dtype() || a.shape().ndim() != b.shape().ndim() || a.byte_size() != b.byte_size()) return false;
    for (size_t i = 0; i < a.shape().ndim(); ++i) if (a.shape().size(i) != b.shape().size(i)) return false;
    return std::memcmp(a.data(), b.data(), a.byte_size()) == 0;
}
static void check_same_model(const model::ModelData& a, const model::ModelData& b, bool params) {
    const auto &ma = a.metadata(), &mb = b.metadata();
    CHECK(ma.name == mb.name && ma.architecture == mb.architecture && ma.version == mb.version);
    CHECK(ma.vocab_size == mb.vocab_size && ma.hidden_size == mb.hidden_size && ma.num_layers == mb.num_layers);
    CHECK(ma.num_heads == mb.num_heads && ma.intermediate_size == mb.intermediate_size && ma.rope_theta == mb.rope_theta);
    CHECK(a.num_tensors() == b.num_tensors());
    for (const std::string& n : a.tensor_names()) {
        const Tensor* tb = b.get_tensor(n);
        CHECK(tb != nullptr);
        if (!tb) continue;
        CHECK(same_tensor(*a.get_tensor(n), *tb));
        const auto *qa = a.quant_params(n), *qb = b.quant_params(n);
        if (params) {
            CHECK((qa == nullptr) == (qb == nullptr));
            if (qa && qb) CHECK(qa->scale == qb->scale && qa->zero_point == qb->zero_point && qa->type == qb->type);
        } else {
            CHECK(qb == nullptr);   // a trailer derived from the integers (the reference's writer) is not taken for parameters
        }
    }
}
static optimize::Quantizer tinq_quantizer(bool int4) {
    optimize::QuantizationConfig qc;
    qc.type = int4 ? optimize::QuantizationType::kInt4 : optimize::QuantizationType::kInt8;
    qc.symmetric = true;
    return optimize::Quantizer(qc);
}
// tinq-write <int8|int4> <path> | tinq-check <int8|int4> <path> <ours|resaved> | tinq-dump <path> <dir> | tinq-errors <dir>
static int tinq_tool(int argc, char** argv) {
    const std::string mode = argv[1];
    if (mode == "tinq-write") {
        tinq_quantizer(std::strcmp(argv[2], "int4") == 0).save_quantized_model(tinq_fixture(std::strcmp(argv[2], "int4") == 0), argv[3]);
    } else if (mode == "tinq-check") {
        const bool int4 = std::strcmp(argv[2], "int4") == 0;
        check_same_model(tinq_fixture(int4), optimize::Quantizer::load_quantized_model(argv[3]), std::strcmp(argv[4], "ours") == 0);
    } else if (mode == "tinq-dump") {   // every tensor's raw bytes + one line of metadata, for a comparison made outside
        model::ModelData md = optimize::Quantizer::load_quantized_model(argv[2]);
        std::ofstream idx(std::string(argv[3]) + "/index.txt");
        const auto& m = md.metadata();
        idx << "meta " << m.name << " " << m.architecture << " " << m.version << " " << m.vocab_size << " " << m.hidden_size << " "
            << m.num_layers << " " << m.num_heads << " " << m.intermediate_size << " " << m.rope_theta << "\n";
        for (const std::string& n : md.tensor_names()) {
            const Tensor* t = md.get_tensor(n);
            idx << "tensor " << n << " " << (int)t->dtype() << " " << (md.quant_params(n) ? 1 : 0);
            for (size_t i = 0; i < t->shape().ndim(); ++i) idx << " " << t->shape().size(i);
            idx << "\n";
            std::ofstream(std::string(argv[3]) + "/" + n + ".bin", std::ios::binary).write(static_cast<const char*>(t->data()), (std::streamsize)t->byte_size());
        }
    } else if (mode == "tinq-errors") {   // the reader's error paths (quantization.cpp:216-231, :289-291, :329-331)
        const std::string dir = argv[2];
        CHECK_THROWS(optimize::Quantizer::load_quantized_model(dir + "/does-not-exist.tinq"), std::runtime_error);
        { std::ofstream f(dir + "/bad-magic.tinq", std::ios::binary); const uint32_t w[2] = {0x12345678u, 1u}; f.write((const char*)w, 8); }
        CHECK_THROWS(optimize::Quantizer::load_quantized_model(dir + "/bad-magic.tinq"), std::runtime_error);
        { std::ofstream f(dir + "/bad-version.tinq", std::ios::binary); const uint32_t w[2] = {0x54494E51u, 2u}; f.write((const char*)w, 8); }
        CHECK_THROWS(optimize::Quantizer::load_quantized_model(dir + "/bad-version.tinq"), std::runtime_error);
        tinq_quantizer(false).save_quantized_model(tinq_fixture(false), dir + "/whole.tinq");
        std::ifstream in(dir + "/whole.tinq", std::ios::binary);
        std::string bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        for (size_t cut : {bytes.size() - 1, bytes.size() / 2, (size_t)40, (size_t)9}) {   // truncated anywhere: an error, never a partial model
            std::ofstream(dir + "/cut.tinq", std::ios::binary).write(bytes.data(), (std::streamsize)cut);
            CHECK_THROWS(optimize::Quantizer::load_quantized_model(dir + "/cut.tinq"), std::runtime_error);
        }
        CHECK_THROWS(tinq_quantizer(false).save_quantized_model(tinq_fixture(false), dir + "/no/such/dir/x.tinq"), std::runtime_error);
    } else {
        std::printf("unknown mode %s\n", mode.c_str());
        return 2;
    }
    std::printf("%s: %d check(s) failed\n", mode.c_str(), g_failed);
    return g_failed == 0 ? 0 : 1;
}

// quantize on the device -> save -> load -> engine: the loaded integers reach the device as they are, so the tokens equal those of
// the engine that quantizes the same float32 projections itself (and the embedding / norm rows, which the engine reads as float32,
// are dequantized from the file)
static void test_tinq_engine_gpu(const char* dir) {
    for (const char* quant : {"int8", "int4"}) {
        const bool int4 = std::strcmp(quant, "int4") == 0;
        model::ModelData fp = small_model(quant);
        optimize::Quantizer qz = tinq_quantizer(int4);
        model::ModelData qm = qz.quantize_model(fp);
        for (const std::string& n : fp.tensor_names()) CHECK(qm.quant_params(n) != nullptr && qm.quant_params(n)->scale > 0.f);
        const std::string path = std::string(dir) + "/small-" + quant + ".tinq";
        qz.save_quantized_model(qm, path);
        model::ModelData back = optimize::Quantizer::load_quantized_model(path);
        check_same_model(qm, back, true);
        // reference engine for the comparison: float32 projections (quantized on the device with the same formulas) over the SAME
        // dequantized embedding / norm rows the file holds
        model::ModelData mixed = small_model(quant);
        for (const std::string& n : fp.tensor_names()) {
            if (n.find("embed") == std::string::npos && n.find("norm") == std::string::npos) continue;
            optimize::QuantizationInfo info;
            info.type = int4 ? optimize::QuantizationType::kInt4 : optimize::QuantizationType::kInt8;
            info.scales = {qm.quant_params(n)->scale};
            info.zero_points = {qm.quant_params(n)->zero_point};
            mixed.add_tensor(n, qz.dequantize_tensor(*qm.get_tensor(n), info));
        }
        model::InferenceConfig cfg;
        cfg.top_k = 1;
        cfg.max_sequence_length = 64;
        model::InferenceEngine from_file(back, cfg), from_float(mixed, cfg);
        const std::vector<int> prompt = {1, 15, 25, 35};
        const auto a = from_file.generate(prompt, 16), b = from_float.generate(prompt, 16);
        CHECK(a.tokens.size() > prompt.size() && a.tokens == b.tokens);
        Tensor la = from_file.forward_pass_incremental({7}), lb = from_float.forward_pass_incremental({7});
        CHECK(same_tensor(la, lb));                                        // same packed integers, same arithmetic: bit-identical logits
        // a file written by the reference carries no usable parameters: refused with a message, not run on bare integers
        model::ModelData no_params;
        no_params.metadata() = back.metadata();
        for (const std::string& n : back.tensor_names()) no_params.add_tensor(n, *back.get_tensor(n));
        CHECK_THROWS(model::InferenceEngine(no_params, cfg), std::runtime_error);
    }
}

int main(int argc, char** argv) {
    if (argc > 2 && std::strncmp(argv[1], "tinq-", 5) == 0) return tinq_tool(argc, argv);
    if (argc > 2 && std::strcmp(argv[1], "gpu-tinq") == 0) {
        test_tinq_engine_gpu(argv[2]);
        std::printf("gpu-tinq: %d check(s) failed\n", g_failed);
        return g_failed == 0 ? 0 : 1;
    }
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    test_host_types();
    test_no_cpu_fallback(gpu);
    if (gpu) {
        test_ops_gpu();
        test_quantizer_gpu();
        test_engine_gpu();
    }
    std::printf("%s: %d check(s) failed\n", gpu ? "gpu" : "cpu", g_failed);
    return g_failed == 0 ? 0 : 1;
}
