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

private:
    ModelMetadata metadata_;
    std::unordered_map<std::string, core::Tensor> tensors_;
};

}  // namespace model
}  // namespace turboinfer
