// turboinfer/optimize/quantization.hpp -- the quantizer of the B200 build (reference include/turboinfer/optimize/
// quantization.hpp:24-271).  Integers and scale / zero-point are bit-exact with the reference; all arithmetic runs in
// the sm_100a kernels behind the C ABI (ti_b200_quant_info / quantize / dequantize).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../core/tensor.hpp"
#include "../model/model_loader.hpp"

namespace turboinfer {
namespace optimize {

enum class QuantizationType { kInt8, kInt4, kFloat16, kNone };

struct QuantizationConfig {
    QuantizationType type = QuantizationType::kInt8;
    bool symmetric = true;
    bool per_channel = true;          // never read by the reference either: scales are per tensor
    float calibration_ratio = 0.1f;
    std::string calibration_dataset;
};

struct QuantizationInfo {
    QuantizationType type;
    std::vector<float> scales;
    std::vector<float> zero_points;
    size_t original_size_bytes;
    size_t quantized_size_bytes;
    float compression_ratio;
};

class Quantizer {
public:
    explicit Quantizer(const QuantizationConfig& config = QuantizationConfig{});
    const QuantizationConfig& config() const noexcept { return config_; }
    void set_config(const QuantizationConfig& config) { config_ = config; }

    core::Tensor quantize_tensor(const core::Tensor& input);   // kInt8 -> int8 tensor, kInt4 -> int32 tensor
    core::Tensor dequantize_tensor(const core::Tensor& quantized, const QuantizationInfo& info);
    model::ModelData quantize_model(const model::ModelData& model_data);
    QuantizationInfo calculate_quantization_info(const core::Tensor& input);
    // .tinq files (reference quantization.cpp:120-333): the same byte layout -- magic "TINQ", version 1, config, metadata,
    // tensors with their (scales, zero_points, sizes) trailer -- so files travel both ways between the two implementations.
    // The trailer holds the tensor's REAL scale / zero-point (ModelData::quant_params) where the reference writes values
    // re-derived from the integer range (:737-816); a file without usable parameters loads as plain integer tensors.
    void save_quantized_model(const model::ModelData& quantized_model, const std::string& output_path);
    static model::ModelData load_quantized_model(const std::string& model_path);

private:
    QuantizationConfig config_;
};

const char* quantization_type_to_string(QuantizationType type);
size_t get_quantization_bits(QuantizationType type);

void quantize_to_int8(const float* input, int8_t* output, size_t count, const QuantizationInfo& info);
void quantize_to_int4(const float* input, int32_t* output, size_t count, const QuantizationInfo& info);
void dequantize_from_int8(const int8_t* input, float* output, size_t count, const QuantizationInfo& info);
void dequantize_from_int4(const int32_t* input, float* output, size_t count, const QuantizationInfo& info);

}  // namespace optimize
}  // namespace turboinfer
