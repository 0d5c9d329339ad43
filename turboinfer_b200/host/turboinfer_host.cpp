// turboinfer_host.cpp -- the reference's C++ class surface (Tensor, TensorEngine, Quantizer, InferenceEngine)
// implemented ONLY in terms of the C ABI of libturboinfer_b200.so (include/ti_b200.h).  No CUDA types, no CPU
// arithmetic on the hot path: every op uploads, runs the sm_100a kernel and downloads; InferenceEngine keeps the
// model resident on the device.  A non-zero C status becomes std::runtime_error carrying ti_b200_last_error(),
// which is the exception type the reference throws for shape / dtype / empty-input errors
// (src/core/tensor_engine.cpp:492-494, src/model/inference_engine.cpp:1409-1417).
#include <fstream>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>
#include <random>
#include <sstream>

#include "../../include/ti_b200.h"
#include "../../include/turboinfer/core/tensor_engine.hpp"
#include "../../include/turboinfer/model/inference_engine.hpp"
#include "../../include/turboinfer/optimize/quantization.hpp"

namespace turboinfer {
namespace {

void ck(int rc) {
    if (rc != 0) throw std::runtime_error(ti_b200_last_error());
}
void ensure_device() {
    static bool ready = false;
    if (ready) return;
    ck(ti_b200_init(0));
    ready = true;
}
const float* f32(const core::Tensor& t, const char* what) {
    if (t.dtype() != core::DataType::kFloat32) throw std::runtime_error(std::string(what) + ": only float32 tensors are supported");
    return static_cast<const float*>(t.data());
}
[[noreturn]] void off_path(const char* op) {
    throw std::runtime_error(std::string("TensorEngine::") + op + " is not part of the B200 token-generation hot path");
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// core::Tensor
// ---------------------------------------------------------------------------------------------------------------
namespace core {

size_t get_dtype_size(DataType dtype) {
    switch (dtype) {
        case DataType::kFloat32: case DataType::kInt32: return 4;
        case DataType::kFloat16: case DataType::kInt16: return 2;
        case DataType::kInt8: case DataType::kUInt8: return 1;
    }
    throw std::runtime_error("Unknown data type");
}
const char* dtype_to_string(DataType dtype) {
    switch (dtype) {
        case DataType::kFloat32: return "float32";
        case DataType::kFloat16: return "float16";
        case DataType::kInt32: return "int32";
        case DataType::kInt16: return "int16";
        case DataType::kInt8: return "int8";
        case DataType::kUInt8: return "uint8";
    }
    return "unknown";
}

bool TensorShape::is_broadcastable_with(const TensorShape& o) const noexcept {
    const size_t n = std::max(ndim(), o.ndim());
    for (size_t i = 0; i < n; ++i) {
        const size_t a = i < ndim() ? dimensions_[ndim() - 1 - i] : 1;
        const size_t b = i < o.ndim() ? o.dimensions_[o.ndim() - 1 - i] : 1;
        if (a != b && a != 1 && b != 1) return false;
    }
    return true;
}

Tensor::Tensor(const TensorShape& shape, DataType dtype) : shape_(shape), dtype_(dtype) {
    const size_t bytes = byte_size();
    if (bytes > 0) data_ = std::make_unique<uint8_t[]>(bytes);  // value-initialised: zeros, like the reference
}
Tensor::Tensor(const TensorShape& shape, const void* data, DataType dtype) : Tensor(shape, dtype) {
    if (data && byte_size() > 0) std::memcpy(data_.get(), data, byte_size());
}
Tensor::Tensor(const Tensor& other) : Tensor(other.shape_, other.data_.get(), other.dtype_) {}
Tensor& Tensor::operator=(const Tensor& other) {
    if (this != &other) {
        Tensor tmp(other);
        *this = std::move(tmp);
    }
    return *this;
}
Tensor Tensor::reshape(const TensorShape& new_shape) const {
    if (new_shape.total_size() != shape_.total_size()) throw std::runtime_error("New shape must have same total size");
    return Tensor(new_shape, data_.get(), dtype_);
}
Tensor Tensor::slice(const std::vector<size_t>& start, const std::vector<size_t>& end) const {
    const size_t nd = shape_.ndim();
    if (start.size() != nd || end.size() != nd) throw std::runtime_error("Slice indices must match tensor dimensions");
    std::vector<size_t> dims(nd);
    for (size_t i = 0; i < nd; ++i) {
        if (start[i] >= end[i] || end[i] > shape_.size(i)) throw std::runtime_error("Invalid slice range");
        dims[i] = end[i] - start[i];
    }
    Tensor out{TensorShape(dims), dtype_};
    const size_t es = element_size();
    std::vector<size_t> stride(nd, 1);
    for (size_t i = nd; i-- > 1;) stride[i - 1] = stride[i] * shape_.size(i);
    std::vector<size_t> idx(nd, 0);
    const size_t run = dims[nd - 1];
    const size_t rows = out.shape().total_size() / run;
    uint8_t* dst = static_cast<uint8_t*>(out.data());
    for (size_t r = 0; r < rows; ++r) {
        size_t off = 0;
        for (size_t i = 0; i < nd; ++i) off += (start[i] + idx[i]) * stride[i];
        std::memcpy(dst + r * run * es, data_.get() + off * es, run * es);
        for (size_t i = nd - 1; i-- > 0;) {
            if (++idx[i] < dims[i]) break;
            idx[i] = 0;
        }
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// core::TensorEngine
// ---------------------------------------------------------------------------------------------------------------
TensorEngine::TensorEngine(ComputeDevice device) : device_(device) {
    if (device == ComputeDevice::kCPU)
        throw std::runtime_error("ComputeDevice::kCPU: this library has no CPU fallback (the CPU reference is a separate oracle)");
    ensure_device();
    device_ = ComputeDevice::kGPU;
}
bool TensorEngine::gpu_available() const noexcept {
    int n = 0;
    return ti_b200_device_count(&n) == 0 && n > 0;
}
std::string TensorEngine::device_info() const {
    char buf[512];
    ck(ti_b200_device_info(buf, sizeof(buf)));
    return buf;
}

Tensor TensorEngine::matmul(const Tensor& a, const Tensor& b) {
    if (a.empty() || b.empty()) throw std::runtime_error("Cannot perform matrix multiplication on empty tensors");
    const size_t nda = a.shape().ndim();
    if ((nda != 2 && nda != 3) || b.shape().ndim() != 2)
        throw std::runtime_error("Matrix multiplication on the B200 path supports [M,K] or [B,T,K] times [K,N]");
    const size_t K = a.shape().size(nda - 1), N = b.shape().size(1);
    if (b.shape().size(0) != K) throw std::runtime_error("Matrix dimensions incompatible for multiplication");
    const size_t M = a.shape().total_size() / K;
    std::vector<float> bf;
    const float* bp = nullptr;
    switch (b.dtype()) {   // convert_dtype (:2218-2284): integers are promoted without a scale
        case DataType::kFloat32: bp = static_cast<const float*>(b.data()); break;
        case DataType::kInt8: { auto* p = b.data_ptr<int8_t>(); bf.assign(p, p + K * N); bp = bf.data(); break; }
        case DataType::kUInt8: { auto* p = b.data_ptr<uint8_t>(); bf.assign(p, p + K * N); bp = bf.data(); break; }
        case DataType::kInt32: { auto* p = b.data_ptr<int32_t>(); bf.assign(p, p + K * N); bp = bf.data(); break; }
        default: throw std::runtime_error("Unsupported weight data type for matmul");
    }
    std::vector<size_t> od = a.shape().dimensions();
    od.back() = N;
    Tensor out{TensorShape(od)};
    ck(ti_b200_matmul_f32(f32(a, "matmul"), bp, out.data_ptr<float>(), M, K, N));
    return out;
}

namespace {
Tensor unary(const Tensor& x, int (*fn)(const float*, float*, size_t), const char* what) {
    if (x.empty()) throw std::runtime_error(std::string("Cannot apply ") + what + " to empty tensor");
    Tensor out(x.shape());
    ck(fn(f32(x, what), out.data_ptr<float>(), x.shape().total_size()));
    return out;
}
Tensor binary(const Tensor& a, const Tensor& b, int (*fn)(const float*, const float*, float*, size_t), const char* what) {
    if (a.empty() || b.empty()) throw std::runtime_error(std::string("Cannot ") + what + " empty tensors");
    if (a.shape() != b.shape()) throw std::runtime_error(std::string("Tensor shapes must match for ") + what);
    Tensor out(a.shape());
    ck(fn(f32(a, what), f32(b, what), out.data_ptr<float>(), a.shape().total_size()));
    return out;
}
}  // namespace

Tensor TensorEngine::relu(const Tensor& input) { return unary(input, ti_b200_relu, "relu"); }
Tensor TensorEngine::silu(const Tensor& input) { return unary(input, ti_b200_silu, "silu"); }
Tensor TensorEngine::add(const Tensor& a, const Tensor& b) { return binary(a, b, ti_b200_add, "add"); }
Tensor TensorEngine::multiply(const Tensor& a, const Tensor& b) { return binary(a, b, ti_b200_mul, "multiply"); }
Tensor TensorEngine::scale(const Tensor& input, float s) {
    Tensor k(input.shape());
    k.fill<float>(s);
    return multiply(input, k);
}
Tensor TensorEngine::softmax(const Tensor& input, float temperature) {
    if (input.empty()) throw std::runtime_error("Cannot apply softmax to empty tensor");
    if (temperature <= 0.0f) throw std::runtime_error("Temperature must be positive");
    const size_t n = input.shape().size(input.shape().ndim() - 1);
    Tensor out(input.shape());
    ck(ti_b200_softmax(f32(input, "softmax"), out.data_ptr<float>(), input.shape().total_size() / n, n, temperature));
    return out;
}
Tensor TensorEngine::rms_norm(const Tensor& input, const Tensor& weight, float eps) {
    if (input.empty() || weight.empty()) throw std::runtime_error("Cannot apply RMS normalization to empty tensors");
    const size_t H = input.shape().size(input.shape().ndim() - 1);
    if (weight.shape().total_size() != H) throw std::runtime_error("Weight tensor size must match the last dimension");
    Tensor out(input.shape());
    ck(ti_b200_rms_norm(f32(input, "rms_norm"), f32(weight, "rms_norm"), out.data_ptr<float>(), input.shape().total_size() / H, H, eps));
    return out;
}
Tensor TensorEngine::apply_rope(const Tensor& input, const Tensor& position_ids, float rope_theta) {
    if (input.empty() || position_ids.empty()) throw std::runtime_error("Cannot apply RoPE to empty tensors");
    const size_t nd = input.shape().ndim();
    if (nd != 3 && nd != 4) throw std::runtime_error("RoPE supports 3D or 4D input tensors only");
    const size_t B = input.shape().size(0), nh = nd == 4 ? input.shape().size(1) : 1, T = input.shape().size(nd - 2),
                 D = input.shape().size(nd - 1);
    const int pos_2d = position_ids.shape().ndim() == 2 ? 1 : 0;
    Tensor out(input.shape());
    ck(ti_b200_rope(f32(input, "apply_rope"), f32(position_ids, "apply_rope"), out.data_ptr<float>(), B, nh, T, D, (int)nd, pos_2d, rope_theta));
    return out;
}
Tensor TensorEngine::attention_fast_incremental(const Tensor& q, const Tensor& k, const Tensor& v, const Tensor*) {
    return multi_head_attention(q, k, v, 1, nullptr);   // the mask argument is ignored by the reference too (:1254)
}
Tensor TensorEngine::multi_head_attention(const Tensor& q, const Tensor& k, const Tensor& v, size_t num_heads, const Tensor*) {
    if (q.empty() || k.empty() || v.empty()) throw std::runtime_error("Cannot compute attention with empty tensors");
    if (q.shape().ndim() != 3 || k.shape().ndim() != 3 || v.shape().ndim() != 3)
        throw std::runtime_error("Attention expects [batch, seq, hidden] tensors");
    if (q.shape().size(1) != 1) throw std::runtime_error("The B200 attention entry point is the decode case: query length must be 1");
    const size_t B = q.shape().size(0), H = q.shape().size(2), t = k.shape().size(1);
    if (k.shape().size(0) != B || v.shape() != k.shape() || k.shape().size(2) != H)
        throw std::runtime_error("Attention tensor shapes are inconsistent");
    Tensor out(q.shape());
    ck(ti_b200_attention_decode(f32(q, "attention"), f32(k, "attention"), f32(v, "attention"), out.data_ptr<float>(), B, t, H, num_heads));
    return out;
}
Tensor TensorEngine::batch_matmul(const Tensor&, const Tensor&) { off_path("batch_matmul"); }
Tensor TensorEngine::add_bias(const Tensor&, const Tensor&) { off_path("add_bias"); }
Tensor TensorEngine::gelu(const Tensor&) { off_path("gelu"); }
Tensor TensorEngine::attention(const Tensor&, const Tensor&, const Tensor&, const Tensor*) { off_path("attention (prefill)"); }
Tensor TensorEngine::layer_norm(const Tensor&, const Tensor&, const Tensor&, float) { off_path("layer_norm"); }

}  // namespace core

// ---------------------------------------------------------------------------------------------------------------
// optimize::Quantizer
// ---------------------------------------------------------------------------------------------------------------
namespace optimize {
namespace {
int c_qtype(QuantizationType t) {
    if (t == QuantizationType::kInt8) return TI_Q_INT8;
    if (t == QuantizationType::kInt4) return TI_Q_INT4;
    throw std::runtime_error("Unsupported quantization type");
}
}  // namespace

Quantizer::Quantizer(const QuantizationConfig& config) : config_(config) {}

QuantizationInfo Quantizer::calculate_quantization_info(const core::Tensor& input) {
    if (input.empty()) throw std::runtime_error("Cannot quantize an empty tensor");
    ensure_device();
    float scale = 0.f, zp = 0.f;
    const size_t n = input.shape().total_size();
    ck(ti_b200_quant_info(f32(input, "quantize"), n, c_qtype(config_.type), config_.symmetric ? 1 : 0, &scale, &zp));
    QuantizationInfo info;
    info.type = config_.type;
    info.scales = {scale};
    info.zero_points = {zp};
    info.original_size_bytes = n * sizeof(float);
    info.quantized_size_bytes = config_.type == QuantizationType::kInt8 ? n : n * sizeof(int32_t);  // INT4 lives in int32
    info.compression_ratio = (float)info.original_size_bytes / (float)info.quantized_size_bytes;
    return info;
}
core::Tensor Quantizer::quantize_tensor(const core::Tensor& input) {
    const QuantizationInfo info = calculate_quantization_info(input);
    const size_t n = input.shape().total_size();
    if (config_.type == QuantizationType::kInt8) {
        core::Tensor out(input.shape(), core::DataType::kInt8);
        quantize_to_int8(static_cast<const float*>(input.data()), out.data_ptr<int8_t>(), n, info);
        return out;
    }
    core::Tensor out(input.shape(), core::DataType::kInt32);
    quantize_to_int4(static_cast<const float*>(input.data()), out.data_ptr<int32_t>(), n, info);
    return out;
}
core::Tensor Quantizer::dequantize_tensor(const core::Tensor& q, const QuantizationInfo& info) {
    if (q.empty()) throw std::runtime_error("Cannot dequantize an empty tensor");
    core::Tensor out(q.shape(), core::DataType::kFloat32);
    const size_t n = q.shape().total_size();
    if (info.type == QuantizationType::kInt8) dequantize_from_int8(q.data_ptr<int8_t>(), out.data_ptr<float>(), n, info);
    else if (info.type == QuantizationType::kInt4) dequantize_from_int4(q.data_ptr<int32_t>(), out.data_ptr<float>(), n, info);
    else throw std::runtime_error("Unsupported quantization type for dequantization");
    return out;
}
model::ModelData Quantizer::quantize_model(const model::ModelData& model_data) {
    model::ModelData out;
    out.metadata() = model_data.metadata();
    for (const std::string& name : model_data.tensor_names()) {
        const core::Tensor* t = model_data.get_tensor(name);
        if (t->dtype() == core::DataType::kFloat32 && !t->empty()) {   // :89-118
            const QuantizationInfo info = calculate_quantization_info(*t);
            out.add_tensor(name, quantize_tensor(*t));
            out.set_quant_params(name, {info.scales[0], info.zero_points[0], (int)config_.type});   // kept (the reference drops them, R8)
        } else {
            out.add_tensor(name, *t);
        }
    }
    return out;
}

namespace {
// little-endian scalar IO with the reference's field widths (quantization.cpp:120-333 writes raw object bytes on x86-64:
// enum = 4, bool = 1, size_t = 8, float = 4; strings = u32 length + bytes, :714-735)
template <typename T> void put(std::ofstream& f, T v) { f.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <typename T> T get(std::ifstream& f) {
    T v{};
    f.read(reinterpret_cast<char*>(&v), sizeof(T));
    if (!f) throw std::runtime_error("unexpected end of file");
    return v;
}
void put_string(std::ofstream& f, const std::string& s) {
    put<uint32_t>(f, (uint32_t)s.size());
    if (!s.empty()) f.write(s.data(), (std::streamsize)s.size());
}
std::string get_string(std::ifstream& f) {
    const uint32_t n = get<uint32_t>(f);
    if (n > (1u << 20)) throw std::runtime_error("implausible string length");
    std::string s(n, '\0');
    if (n) f.read(&s[0], n);
    if (!f) throw std::runtime_error("unexpected end of file");
    return s;
}
// true when (scale, zp) is what the reference's writer computes from the stored integers alone (quantization.cpp:737-816)
bool derived_from_integers(const core::Tensor& t, float scale, float zp) {
    const size_t n = t.shape().total_size();
    if (n == 0) return true;
    long lo, hi;
    float full, fallback;
    if (t.dtype() == core::DataType::kInt8) {
        const int8_t* d = static_cast<const int8_t*>(t.data());
        lo = hi = d[0];
        for (size_t i = 1; i < n; ++i) { lo = std::min<long>(lo, d[i]); hi = std::max<long>(hi, d[i]); }
        full = 255.0f; fallback = 1.0f / 127.0f;
    } else {
        const int32_t* d = static_cast<const int32_t*>(t.data());
        lo = hi = d[0];
        for (size_t i = 1; i < n; ++i) { lo = std::min<long>(lo, d[i]); hi = std::max<long>(hi, d[i]); }
        lo = std::max(lo, -8L); hi = std::min(hi, 7L);
        full = 15.0f; fallback = 1.0f / 7.0f;
    }
    const float range = (float)(hi - lo);
    if (range > 0) return scale == range / full && zp == (float)(-lo);
    return scale == fallback && zp == 0.0f;
}
size_t dtype_bytes(core::DataType d) {
    switch (d) {
        case core::DataType::kFloat32: case core::DataType::kInt32: return 4;
        case core::DataType::kFloat16: case core::DataType::kInt16: return 2;
        default: return 1;
    }
}
}  // namespace

void Quantizer::save_quantized_model(const model::ModelData& qm, const std::string& output_path) {
    std::ofstream f(output_path, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Failed to open file for writing: " + output_path);
    put<uint32_t>(f, 0x54494E51u);                       // "TINQ" (:128)
    put<uint32_t>(f, 1u);                                // version (:132)
    put<int32_t>(f, (int32_t)config_.type);              // :136-138
    put<uint8_t>(f, config_.symmetric ? 1 : 0);
    put<uint8_t>(f, config_.per_channel ? 1 : 0);
    const model::ModelMetadata& md = qm.metadata();      // :141-150
    put_string(f, md.name);
    put_string(f, md.architecture);
    put_string(f, md.version);
    put<uint64_t>(f, md.vocab_size);
    put<uint64_t>(f, md.hidden_size);
    put<uint64_t>(f, md.num_layers);
    put<uint64_t>(f, md.num_heads);
    put<uint64_t>(f, md.intermediate_size);
    put<float>(f, md.rope_theta);
    const std::vector<std::string> names = qm.tensor_names();
    put<uint32_t>(f, (uint32_t)names.size());            // :153-155
    for (const std::string& name : names) {
        const core::Tensor* t = qm.get_tensor(name);
        put_string(f, name);                             // :163-181
        put<uint32_t>(f, (uint32_t)t->dtype());
        put<uint32_t>(f, (uint32_t)t->shape().ndim());
        for (size_t i = 0; i < t->shape().ndim(); ++i) put<uint64_t>(f, t->shape().size(i));
        put<uint64_t>(f, t->byte_size());
        f.write(static_cast<const char*>(t->data()), (std::streamsize)t->byte_size());
        if (t->dtype() == core::DataType::kInt8 || t->dtype() == core::DataType::kInt32) {   // the trailer (:184-205)
            const model::ModelData::QuantParams* q = qm.quant_params(name);
            put<uint32_t>(f, q ? 1u : 0u);
            if (q) put<float>(f, q->scale);
            put<uint32_t>(f, q ? 1u : 0u);
            if (q) put<float>(f, q->zero_point);
            const uint64_t orig = t->shape().total_size() * sizeof(float);
            put<uint64_t>(f, orig);
            put<uint64_t>(f, t->byte_size());
            put<float>(f, (float)orig / (float)t->byte_size());
        }
    }
    if (!f) throw std::runtime_error("Failed to save quantized model: write error");
}

model::ModelData Quantizer::load_quantized_model(const std::string& model_path) {
    std::ifstream f(model_path, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Failed to open file for reading: " + model_path);
    try {
        if (get<uint32_t>(f) != 0x54494E51u) throw std::runtime_error("Invalid file format - not a TurboInfer quantized model");   // :222-224
        const uint32_t version = get<uint32_t>(f);
        if (version != 1) throw std::runtime_error("Unsupported quantized model version: " + std::to_string(version));          // :229-231
        const int32_t type = get<int32_t>(f);
        (void)get<uint8_t>(f);   // symmetric
        (void)get<uint8_t>(f);   // per_channel
        model::ModelData md;
        model::ModelMetadata& m = md.metadata();
        m.name = get_string(f);
        m.architecture = get_string(f);
        m.version = get_string(f);
        m.vocab_size = get<uint64_t>(f);
        m.hidden_size = get<uint64_t>(f);
        m.num_layers = get<uint64_t>(f);
        m.num_heads = get<uint64_t>(f);
        m.intermediate_size = get<uint64_t>(f);
        m.rope_theta = get<float>(f);
        const uint32_t count = get<uint32_t>(f);
        for (uint32_t i = 0; i < count; ++i) {
            const std::string name = get_string(f);
            const uint32_t dt = get<uint32_t>(f);
            if (dt > (uint32_t)core::DataType::kUInt8) throw std::runtime_error("unknown tensor data type in '" + name + "'");
            const core::DataType dtype = (core::DataType)dt;
            const uint32_t ndim = get<uint32_t>(f);
            if (ndim == 0 || ndim > 8) throw std::runtime_error("implausible tensor rank in '" + name + "'");
            std::vector<size_t> dims;
            size_t total = 1;
            for (uint32_t d = 0; d < ndim; ++d) { dims.push_back(get<uint64_t>(f)); total *= dims.back(); }
            const uint64_t bytes = get<uint64_t>(f);
            if (bytes != total * dtype_bytes(dtype)) throw std::runtime_error("Tensor size mismatch for: " + name);                // :289-291
            core::Tensor t(core::TensorShape(dims), dtype);
            f.read(static_cast<char*>(t.data()), (std::streamsize)bytes);
            if (!f) throw std::runtime_error("unexpected end of file");
            if (dtype == core::DataType::kInt8 || dtype == core::DataType::kInt32) {
                std::vector<float> scales(get<uint32_t>(f));
                for (float& v : scales) v = get<float>(f);
                std::vector<float> zps(get<uint32_t>(f));
                for (float& v : zps) v = get<float>(f);
                (void)get<uint64_t>(f);
                (void)get<uint64_t>(f);
                (void)get<float>(f);
                // The reference's writer re-derives (scale, zero_point) from the INTEGER range (:737-816): such a trailer carries no
                // information about the original values, so the tensor then loads as plain integers, without parameters.
                if (scales.size() == 1 && zps.size() == 1 && !derived_from_integers(t, scales[0], zps[0]))
                    md.set_quant_params(name, {scales[0], zps[0], type});
            }
            md.add_tensor(name, std::move(t));
        }
        return md;
    } catch (const std::exception& e) {
        throw std::runtime_error("Failed to load quantized model: " + std::string(e.what()));   // :329-331
    }
}

const char* quantization_type_to_string(QuantizationType type) {
    switch (type) {
        case QuantizationType::kInt8: return "int8";
        case QuantizationType::kInt4: return "int4";
        case QuantizationType::kFloat16: return "float16";
        case QuantizationType::kNone: return "none";
    }
    return "unknown";
}
size_t get_quantization_bits(QuantizationType type) {
    switch (type) {
        case QuantizationType::kInt8: return 8;
        case QuantizationType::kInt4: return 4;
        case QuantizationType::kFloat16: return 16;
        case QuantizationType::kNone: return 32;
    }
    return 32;
}

namespace {
void info_params(const QuantizationInfo& info, float& scale, float& zp) {
    if (info.scales.empty() || info.zero_points.empty()) throw std::runtime_error("Quantization info has no scale / zero point");
    scale = info.scales[0];
    zp = info.zero_points[0];
}
}  // namespace
void quantize_to_int8(const float* input, int8_t* output, size_t count, const QuantizationInfo& info) {
    float s, z;
    info_params(info, s, z);
    ensure_device();
    ck(ti_b200_quantize(input, count, TI_Q_INT8, s, z, output));
}
void quantize_to_int4(const float* input, int32_t* output, size_t count, const QuantizationInfo& info) {
    float s, z;
    info_params(info, s, z);
    ensure_device();
    ck(ti_b200_quantize(input, count, TI_Q_INT4, s, z, output));
}
void dequantize_from_int8(const int8_t* input, float* output, size_t count, const QuantizationInfo& info) {
    float s, z;
    info_params(info, s, z);
    ensure_device();
    ck(ti_b200_dequantize(input, count, TI_Q_INT8, s, z, output));
}
void dequantize_from_int4(const int32_t* input, float* output, size_t count, const QuantizationInfo& info) {
    float s, z;
    info_params(info, s, z);
    ensure_device();
    ck(ti_b200_dequantize(input, count, TI_Q_INT4, s, z, output));
}

}  // namespace optimize

// ---------------------------------------------------------------------------------------------------------------
// model::InferenceEngine
// ---------------------------------------------------------------------------------------------------------------
namespace model {

struct InferenceEngine::Stats {
    size_t generations = 0, tokens = 0;
    double time_ms = 0.0, peak_tps = 0.0;
    std::mt19937 rng{(uint32_t)std::chrono::steady_clock::now().time_since_epoch().count()};  // time-seeded like the reference (:472)
    bool seed_set = false;
    uint64_t seed = 0;
};

InferenceEngine::InferenceEngine(const ModelData& model_data, const InferenceConfig& config)
    : model_metadata_(model_data.metadata()), config_(config), stats_(std::make_unique<Stats>()) {
    if (config.device == core::ComputeDevice::kCPU)
        throw std::runtime_error("InferenceConfig::device = kCPU: this library has no CPU fallback");
    ensure_device();
    const ModelMetadata& md = model_metadata_;
    auto param = [&](const char* key, const char* dflt) {
        auto it = md.extra_params.find(key);
        return it == md.extra_params.end() ? std::string(dflt) : it->second;
    };
    ti_model_config cfg{};
    cfg.vocab = (int32_t)md.vocab_size;
    cfg.hidden = (int32_t)md.hidden_size;
    cfg.layers = (int32_t)md.num_layers;
    cfg.heads = (int32_t)std::max<size_t>(md.num_heads, 1);
    cfg.inter = (int32_t)md.intermediate_size;
    cfg.rope_theta = md.rope_theta;
    cfg.rms_eps = 1e-5f;
    // a quantized ModelData (Quantizer::quantize_model, a .tinq file) names its own type; otherwise the extra_params key decides
    std::string quant_default = "int8";
    for (const std::string& name : model_data.tensor_names())
        if (const ModelData::QuantParams* q = model_data.quant_params(name)) {
            quant_default = q->type == (int)optimize::QuantizationType::kInt4 ? "int4" : "int8";
            break;
        }
    const std::string quant = param("b200.quantization", quant_default.c_str());
    cfg.qtype = quant == "int4" ? TI_Q_INT4 : (quant == "none" ? TI_Q_NONE : TI_Q_INT8);
    // "literal": the path benchmarks/benchmark_inference runs on the reference today (placeholder embeddings, attention
    // fall-back, unscaled integer weights; quantization "none" = its FP32 variant) -- BASELINE.json configs[0]
    cfg.compat_literal = param("b200.compat", "") == "literal" ? 1 : 0;
    cfg.attn_mode = param("b200.attention", "multi_head") == "single_head" ? 0 : 1;
    const std::string rope = param("b200.rope", "per_head");
    cfg.rope_mode = rope == "none" ? 0 : (rope == "hidden" ? 2 : 1);
    cfg.max_seq = (int32_t)config.max_sequence_length;
    ti_model_t h = 0;
    ck(ti_b200_model_new(&cfg, &h));
    handle_ = h;
    try {
        for (const std::string& name : model_data.tensor_names()) {
            const core::Tensor* t = model_data.get_tensor(name);
            if (t->empty()) continue;
            const size_t nd = t->shape().ndim();
            const size_t cols = t->shape().size(nd - 1), rows = t->shape().total_size() / cols;
            int rc;
            if (t->dtype() == core::DataType::kFloat32) {
                rc = ti_b200_model_set_tensor(h, name.c_str(), static_cast<const float*>(t->data()), rows, cols);
            } else if ((t->dtype() == core::DataType::kInt8 || t->dtype() == core::DataType::kInt32) && model_data.quant_params(name)) {
                // an already quantized tensor (Quantizer::quantize_model, or a .tinq file): its integers go to the device as they are
                const ModelData::QuantParams* q = model_data.quant_params(name);
                const bool i8 = t->dtype() == core::DataType::kInt8;
                const bool row_data = name.find("embed") != std::string::npos || name.find("norm") != std::string::npos;
                if (row_data) {
                    // embedding rows and norm weights are read as float32 by the engine: dequantize them here with the reference's
                    // formulas (dequantize_from_int8 / _int4, quantization.cpp:695-713)
                    std::vector<float> x(rows * cols);
                    for (size_t i = 0; i < x.size(); ++i)
                        x[i] = i8 ? q->scale * ((float)static_cast<const int8_t*>(t->data())[i] - q->zero_point)
                                  : q->scale * ((float)static_cast<const int32_t*>(t->data())[i] + q->zero_point);
                    rc = ti_b200_model_set_tensor(h, name.c_str(), x.data(), rows, cols);
                } else {
                    rc = ti_b200_model_set_tensor_q(h, name.c_str(), t->data(), rows, cols, i8 ? TI_Q_INT8 : TI_Q_INT4, q->scale, q->zero_point);
                }
            } else {
                throw std::runtime_error("tensor '" + name + "': the B200 engine takes float32 weights (quantized on the device, metadata "
                                         "extra_params[\"b200.quantization\"]) or int8 / int32 tensors with their quantization parameters");
            }
            if (rc != 0) {
                const std::string err = ti_b200_last_error();
                if (err.rfind("unknown tensor name", 0) == 0) continue;   // the reference ignores tensors it has no slot for
                throw std::runtime_error(err);
            }
        }
        ck(ti_b200_model_finalize(h));
    } catch (...) {
        ti_b200_model_free(h);
        handle_ = 0;
        throw;
    }
}
InferenceEngine::~InferenceEngine() {
    if (handle_) ti_b200_model_free(handle_);
}
InferenceEngine::InferenceEngine(InferenceEngine&& o) noexcept
    : model_metadata_(std::move(o.model_metadata_)), config_(o.config_), handle_(o.handle_), stats_(std::move(o.stats_)) {
    o.handle_ = 0;
}
InferenceEngine& InferenceEngine::operator=(InferenceEngine&& o) noexcept {
    if (this != &o) {
        if (handle_) ti_b200_model_free(handle_);
        model_metadata_ = std::move(o.model_metadata_);
        config_ = o.config_;
        handle_ = o.handle_;
        stats_ = std::move(o.stats_);
        o.handle_ = 0;
    }
    return *this;
}

void InferenceEngine::validate_input_tokens(const std::vector<int>& tokens) const {
    if (tokens.empty()) throw std::runtime_error("Input tokens cannot be empty");
    if (tokens.size() > config_.max_sequence_length) throw std::runtime_error("Input sequence length exceeds maximum allowed length");
}
void InferenceEngine::reset_state() { ck(ti_b200_model_reset(handle_)); }
void InferenceEngine::set_seed(uint64_t seed) { stats_->seed_set = true; stats_->seed = seed; }

std::vector<float> InferenceEngine::compute_logprobs(const std::vector<int>& tokens) {
    validate_input_tokens(tokens);
    std::vector<float> out(tokens.size());
    // the reference swallows failures and answers with a sentinel (-18: computation error, :947-952)
    if (ti_b200_compute_logprobs(handle_, tokens.data(), (int32_t)tokens.size(), out.data()) != 0) return std::vector<float>(tokens.size(), -18.0f);
    return out;
}

core::Tensor InferenceEngine::forward_pass_incremental(const std::vector<int>& tokens) {
    if (tokens.empty()) throw std::runtime_error("Cannot perform incremental forward pass with empty token sequence");
    const size_t V = model_metadata_.vocab_size;
    core::Tensor logits{core::TensorShape({1, tokens.size(), V})};
    for (size_t i = 0; i < tokens.size(); ++i)
        ck(ti_b200_decode_step(handle_, tokens[i], logits.data_ptr<float>() + i * V, nullptr));
    return logits;
}

core::Tensor InferenceEngine::forward_pass(const std::vector<int>& tokens) {
    validate_input_tokens(tokens);
    reset_state();
    return forward_pass_incremental(tokens);
}

GenerationResult InferenceEngine::generate(const std::vector<int>& input_tokens, size_t max_new_tokens, bool include_logprobs) {
    validate_input_tokens(input_tokens);
    const auto t0 = std::chrono::high_resolution_clock::now();
    GenerationResult result;
    result.tokens = input_tokens;
    const size_t V = model_metadata_.vocab_size;
    // the loop of the reference (:752-775) stops at max_sequence_length tokens in total
    const size_t room = config_.max_sequence_length > input_tokens.size() ? config_.max_sequence_length - input_tokens.size() : 0;
    const size_t budget = std::min(max_new_tokens, std::max<size_t>(room, max_new_tokens > 0 ? 1 : 0));
    if (config_.top_k == 1 && budget > 0) {
        // greedy: prefill + decode + arg-max entirely on the device, one call
        std::vector<int32_t> out(budget);
        std::vector<float> logits;
        if (include_logprobs) logits.resize(budget * V);
        int32_t n_out = 0;
        ck(ti_b200_generate_greedy(handle_, input_tokens.data(), (int32_t)input_tokens.size(), (int32_t)budget, 1, out.data(), &n_out,
                                   include_logprobs ? logits.data() : nullptr, nullptr));
        for (int32_t i = 0; i < n_out; ++i) {
            result.tokens.push_back(out[i]);
            if (include_logprobs) {   // log-softmax of the picked token at temperature T
                const float* l = logits.data() + (size_t)i * V;
                double mx = l[0];
                for (size_t j = 1; j < V; ++j) mx = std::max<double>(mx, l[j]);
                double s = 0.0;
                for (size_t j = 0; j < V; ++j) s += std::exp((l[j] - mx) / config_.temperature);
                result.logprobs.push_back((float)((l[out[i]] - mx) / config_.temperature - std::log(s)));
            }
        }
        if (n_out > 0 && out[n_out - 1] == 2) { result.finished = true; result.stop_reason = "eos_token"; }
        else if (result.tokens.size() >= config_.max_sequence_length) { result.finished = true; result.stop_reason = "max_length"; }
    } else if (budget > 0) {
        // sampling: the token is picked on the device from every step's logits and fed to the next forward pass there
        if (config_.temperature <= 0.0f) throw std::runtime_error("Temperature must be positive");
        std::vector<int32_t> out(budget);
        std::vector<float> lps(budget);
        int32_t n_out = 0;
        const uint64_t seed = stats_->seed_set ? stats_->seed + stats_->generations : stats_->rng();
        ck(ti_b200_generate_sampled(handle_, input_tokens.data(), (int32_t)input_tokens.size(), (int32_t)budget, config_.temperature,
                                    (int32_t)std::min<size_t>(config_.top_k, 0x7fffffff), config_.top_p, seed, 1, out.data(), &n_out,
                                    include_logprobs ? lps.data() : nullptr, nullptr));
        for (int32_t i = 0; i < n_out; ++i) {
            result.tokens.push_back(out[i]);
            if (include_logprobs) result.logprobs.push_back(lps[i]);
        }
        if (n_out > 0 && out[n_out - 1] == 2) { result.finished = true; result.stop_reason = "eos_token"; }
        else if (result.tokens.size() >= config_.max_sequence_length) { result.finished = true; result.stop_reason = "max_length"; }
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    result.total_time_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
    const size_t generated = result.tokens.size() - input_tokens.size();
    result.tokens_per_second = result.total_time_ms > 0.f ? generated / (result.total_time_ms / 1000.0f) : 0.f;
    if (!result.finished) result.stop_reason = "max_new_tokens";
    stats_->generations++;
    stats_->tokens += generated;
    stats_->time_ms += result.total_time_ms;
    stats_->peak_tps = std::max<double>(stats_->peak_tps, result.tokens_per_second);
    return result;
}

std::vector<GenerationResult> InferenceEngine::generate_beam_search(const std::vector<int>& input_tokens, size_t max_new_tokens, size_t beam_size,
                                                                    bool include_logprobs) {
    if (beam_size == 0) throw std::runtime_error("Beam size must be greater than 0");   // :836-838
    validate_input_tokens(input_tokens);
    const auto t0 = std::chrono::high_resolution_clock::now();
    const size_t cap = std::max<size_t>(max_new_tokens, 1);
    std::vector<int32_t> toks(beam_size * cap), lens(beam_size), fin(beam_size);
    std::vector<float> lp(beam_size), score(beam_size);
    int32_t n = 0;
    ck(ti_b200_beam_search(handle_, input_tokens.data(), (int32_t)input_tokens.size(), (int32_t)max_new_tokens, (int32_t)beam_size, config_.temperature,
                           (int32_t)config_.top_k, config_.top_p, config_.length_penalty, config_.eos_token_id, toks.data(), lens.data(), lp.data(),
                           score.data(), fin.data(), &n));
    const float ms = std::chrono::duration<float, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    std::vector<GenerationResult> out;
    for (int32_t i = 0; i < n; ++i) {
        GenerationResult r;
        r.tokens.assign(toks.begin() + i * max_new_tokens, toks.begin() + i * max_new_tokens + lens[i]);
        r.finished = fin[i] != 0;
        if (include_logprobs) r.logprobs.assign(r.tokens.size(), lp[i] / (float)r.tokens.size());   // :862-865
        r.total_time_ms = ms;
        out.push_back(std::move(r));
    }
    stats_->generations++;
    stats_->time_ms += ms;
    return out;
}

std::vector<GenerationResult> InferenceEngine::generate_batch(const std::vector<std::vector<int>>& batch, size_t max_new_tokens,
                                                              bool include_logprobs) {
    if (batch.empty()) throw std::runtime_error("Batch size cannot be zero");
    if (batch.size() > config_.max_batch_size) throw std::runtime_error("Batch size exceeds maximum allowed batch size");
    std::vector<GenerationResult> out;
    // Greedy, no log-probabilities (prompts of any lengths): the sequences advance in lockstep on the device, the weights
    // are read once per step for the whole batch (ti_b200_generate_batch_greedy).  Tokens after a sequence's first EOS
    // are dropped here, which is where the reference's per-sequence loop would have stopped (:760).
    bool lockstep = config_.top_k == 1 && !include_logprobs && batch.size() > 1 && max_new_tokens > 0;
    size_t P = 0;
    for (const auto& tokens : batch) {
        validate_input_tokens(tokens);
        P = std::max(P, tokens.size());   // prompts may differ in length: the sequences are left-aligned on the device
    }
    lockstep = lockstep && P + max_new_tokens - 1 <= config_.max_sequence_length;
    if (lockstep) {
        const auto t0 = std::chrono::high_resolution_clock::now();
        const size_t B = batch.size();
        std::vector<int32_t> prompts(B * P, 0), lens(B), toks(B * max_new_tokens);
        for (size_t b = 0; b < B; ++b) {
            std::copy(batch[b].begin(), batch[b].end(), prompts.begin() + b * P);
            lens[b] = (int32_t)batch[b].size();
        }
        const int rc = ti_b200_generate_batch_ragged(handle_, prompts.data(), lens.data(), (int32_t)B, (int32_t)P, (int32_t)max_new_tokens, toks.data(), nullptr);
        if (rc == 0) {
            const auto t1 = std::chrono::high_resolution_clock::now();
            const float ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
            for (size_t b = 0; b < B; ++b) {
                GenerationResult r;
                r.tokens = batch[b];
                for (size_t i = 0; i < max_new_tokens; ++i) {
                    r.tokens.push_back(toks[b * max_new_tokens + i]);
                    if (toks[b * max_new_tokens + i] == 2) { r.finished = true; r.stop_reason = "eos_token"; break; }
                }
                if (!r.finished && r.tokens.size() >= config_.max_sequence_length) { r.finished = true; r.stop_reason = "max_length"; }
                if (!r.finished) r.stop_reason = "max_new_tokens";
                r.total_time_ms = ms;
                const size_t generated = r.tokens.size() - batch[b].size();
                r.tokens_per_second = ms > 0.f ? generated / (ms / 1000.0f) : 0.f;
                stats_->generations++;
                stats_->tokens += generated;
                out.push_back(std::move(r));
            }
            stats_->time_ms += ms;
            return out;
        }
        // models the lockstep path does not cover (missing projections, literal mode, tensor parallel): one by one
    }
    for (const auto& tokens : batch) out.push_back(generate(tokens, max_new_tokens, include_logprobs));   // :804-828
    return out;
}

size_t InferenceEngine::memory_usage() const {
    double w = 0, kv = 0;
    if (ti_b200_model_step_bytes(handle_, (int32_t)config_.max_sequence_length, &w, &kv) != 0) return 0;
    return (size_t)(w + kv);
}
std::string InferenceEngine::performance_stats() const {
    std::ostringstream os;
    os << "generations: " << stats_->generations << ", tokens: " << stats_->tokens << ", time: " << stats_->time_ms << " ms, average: "
       << (stats_->time_ms > 0 ? stats_->tokens / (stats_->time_ms / 1000.0) : 0.0) << " tok/s, peak: " << stats_->peak_tps << " tok/s";
    return os.str();
}

}  // namespace model
}  // namespace turboinfer
