// turboinfer/model/model_loader.hpp -- the weight hand-off types (reference include/turboinfer/model/model_loader.hpp:35-153).
// File parsing (GGUF / SafeTensors / ...) is out of scope of the B200 hot path; models are built in memory.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "../core/tensor.hpp"

namespace turboinfer {
namespace model {

struct ModelMetadata {
    std::string name;
    std::string architecture;
    std::string version;
    size_t vocab_size = 0;
    size_t hidden_size = 0;
    size_t num_layers = 0;
    size_t num_heads = 0;
    size_t intermediate_size = 0;
    float rope_theta = 10000.0f;
    // B200 build reads: "b200.quantization" = "int4" | "int8" (default int8), "b200.attention" = "multi_head" |
    // "single_head" (default multi_head), "b200.rope" = "per_head" | "none" | "hidden" (default per_head)
    std::unordered_map<std::string, std::string> extra_params;
};

class ModelData {
public:
    ModelData() = default;
    const ModelMetadata& metadata() const noexcept { return metadata_; }
    ModelMetadata& metadata() noexcept { return metadata_; }
    const core::Tensor* get_tensor(const std::string& name) const {
        auto it = tensors_.find(name);
        return it == tensors_.end() ? nullptr : &it->second;
    }
    core::Tensor* get_tensor(const std::string& name) {
        auto it = tensors_.find(name);
        return it == tensors_.end() ? nullptr : &it->second;
    }
    void add_tensor(const std::string& name, core::Tensor tensor) { tensors_.insert_or_assign(name, std::move(tensor)); }
    std::vector<std::string> tensor_names() const {
        std::vector<std::string> n;
        for (auto& kv : tensors_) n.push_back(kv.first);
        return n;
    }
    bool has_tensor(const std::string& name) const { return tensors_.count(name) != 0; }
    size_t num_tensors() const noexcept { return tensors_.size(); }

    // B200 build: the REAL (scale, zero_point) of a quantized (int8 / int32) tensor, as Quantizer::quantize_model computed
    // them.  The reference drops them (its engine multiplies by the bare integers, SURVEY R8, and its .tinq writer re-invents
    // scales from the integer range, quantization.cpp:737-816); keeping them is what lets a quantized model -- in memory or
    // loaded from a .tinq file -- go to the device as integers, packed once, without re-quantizing.
    struct QuantParams { float scale = 0.f, zero_point = 0.f; int type = 0; };   // type: optimize::QuantizationType as int
    void set_quant_params(const std::string& name, QuantParams q) { qparams_[name] = q; }
    const QuantParams* quant_params(const std::string& name) const {
        auto it = qparams_.find(name);
        return it == qparams_.end() ? nullptr : &it->second;
    }

private:
    ModelMetadata metadata_;
    std::unordered_map<std::string, core::Tensor> tensors_;
    std::unordered_map<std::string, QuantParams> qparams_;
};

}  // namespace model
}  // namespace turboinfer
