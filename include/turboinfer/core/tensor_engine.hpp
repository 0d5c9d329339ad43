// turboinfer/core/tensor_engine.hpp -- TensorEngine of the B200 build: the ops of the token-generation hot path
// (reference include/turboinfer/core/tensor_engine.hpp:36-318) as value-in / value-out calls that upload, run the
// sm_100a kernel behind the C ABI (include/ti_b200.h) and download.  They exist for API compatibility and parity
// tests; the fast path (InferenceEngine) keeps everything resident on the device.
// ComputeDevice::kCPU throws: there is no CPU fallback in this library (the CPU reference is the separate oracle).
#pragma once
#include <string>
#include <vector>

#include "tensor.hpp"

namespace turboinfer {
namespace core {

enum class ComputeDevice { kCPU, kGPU, kAuto };

class TensorEngine {
public:
    explicit TensorEngine(ComputeDevice device = ComputeDevice::kAuto);
    ~TensorEngine() = default;

    ComputeDevice device() const noexcept { return device_; }
    bool gpu_available() const noexcept;
    std::string device_info() const;

    // y = a . b with b [K, N] ([in, out]); a is [M, K] or [B, T, K].  fp32 b reproduces the reference build's order
    // of roundings; int8 / int32 b is promoted WITHOUT scale, like the reference's convert_dtype.
    Tensor matmul(const Tensor& a, const Tensor& b);
    Tensor relu(const Tensor& input);
    Tensor silu(const Tensor& input);
    Tensor softmax(const Tensor& input, float temperature = 1.0f);   // scalar branch of the reference
    Tensor attention_fast_incremental(const Tensor& query, const Tensor& key, const Tensor& value, const Tensor* mask = nullptr);
    Tensor multi_head_attention(const Tensor& query, const Tensor& key, const Tensor& value, size_t num_heads,
                                const Tensor* mask = nullptr);      // q_len 1 (decode) only
    Tensor rms_norm(const Tensor& input, const Tensor& weight, float eps = 1e-5f);
    Tensor apply_rope(const Tensor& input, const Tensor& position_ids, float rope_theta = 10000.0f);
    Tensor add(const Tensor& a, const Tensor& b);
    Tensor multiply(const Tensor& a, const Tensor& b);
    Tensor scale(const Tensor& input, float scale);

    // Not on the decode path: these throw std::runtime_error("... not part of the B200 hot path").
    Tensor batch_matmul(const Tensor& a, const Tensor& b);
    Tensor add_bias(const Tensor& input, const Tensor& bias);
    Tensor gelu(const Tensor& input);
    Tensor attention(const Tensor& query, const Tensor& key, const Tensor& value, const Tensor* mask = nullptr);
    Tensor layer_norm(const Tensor& input, const Tensor& weight, const Tensor& bias, float eps = 1e-5f);

private:
    ComputeDevice device_;
};

}  // namespace core
}  // namespace turboinfer
