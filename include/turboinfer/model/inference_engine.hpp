// turboinfer/model/inference_engine.hpp -- InferenceEngine of the B200 build (reference include/turboinfer/model/
// inference_engine.hpp:25-372): weights are quantized, packed and uploaded once at construction; generate() runs the
// persistent decode kernel.  Sampling runs on the device for every setting: greedy (top_k = 1) as an arg-max inside the
// decode kernel, everything else through the on-device sampler (temperature -> top-k -> softmax -> top-p -> inverse CDF,
// sample_next_token :1554-1673) with a seeded counter-based RNG (set_seed; the reference seeds from the clock, :472).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../core/tensor_engine.hpp"
#include "model_loader.hpp"

namespace turboinfer {
namespace model {

struct InferenceConfig {
    size_t max_sequence_length = 2048;
    size_t max_batch_size = 32;
    float temperature = 1.0f;
    float top_p = 0.9f;
    size_t top_k = 50;
    float length_penalty = 1.0f;
    int eos_token_id = 2;          // generate() stops on token 2 regardless, like the reference (:760)
    bool use_cache = true;
    core::ComputeDevice device = core::ComputeDevice::kAuto;
};

struct GenerationResult {
    std::vector<int> tokens;
    std::vector<float> logprobs;
    float total_time_ms = 0.f;
    float tokens_per_second = 0.f;
    bool finished = false;
    std::string stop_reason;
};

class InferenceEngine {
public:
    explicit InferenceEngine(const ModelData& model_data, const InferenceConfig& config = InferenceConfig{});
    ~InferenceEngine();
    InferenceEngine(const InferenceEngine&) = delete;
    InferenceEngine& operator=(const InferenceEngine&) = delete;
    InferenceEngine(InferenceEngine&&) noexcept;
    InferenceEngine& operator=(InferenceEngine&&) noexcept;

    const ModelMetadata& model_metadata() const noexcept { return model_metadata_; }
    const InferenceConfig& config() const noexcept { return config_; }
    void set_config(const InferenceConfig& config) { config_ = config; }

    GenerationResult generate(const std::vector<int>& input_tokens, size_t max_new_tokens, bool include_logprobs = false);
    std::vector<GenerationResult> generate_batch(const std::vector<std::vector<int>>& input_tokens_batch, size_t max_new_tokens,
                                                 bool include_logprobs = false);
    /// generate_beam_search (:830-871): one result per beam, best normalised score first; tokens = the NEW tokens only (:853-857),
    /// logprobs (when asked) = the average log-probability repeated per token (:862-865).  On the device the beams share the
    /// prompt's KV pages and fork by page table (ti_b200_beam_search).  throws std::runtime_error for beam_size == 0.
    std::vector<GenerationResult> generate_beam_search(const std::vector<int>& input_tokens, size_t max_new_tokens, size_t beam_size = 4,
                                                       bool include_logprobs = false);
    /// compute_logprobs (:873-954): log softmax(logits[pos])[tokens[pos]] per position; the reference's sentinels on failure
    std::vector<float> compute_logprobs(const std::vector<int>& tokens);
    /// seed of the sampler's counter-based RNG (a generation with the same seed reproduces its tokens); default: from the clock
    void set_seed(uint64_t seed);
    void reset_state();
    size_t memory_usage() const;
    std::string performance_stats() const;

    // exposed for tests (private in the reference, :248-255)
    /// forward_pass_incremental (:1493-1552): appends the tokens to the cached sequence; logits [1, tokens, vocab]
    core::Tensor forward_pass_incremental(const std::vector<int>& tokens);
    /// forward_pass (:1429-1491): the whole sequence from an empty cache; logits of every position, [1, tokens, vocab]
    core::Tensor forward_pass(const std::vector<int>& tokens);

private:
    void validate_input_tokens(const std::vector<int>& tokens) const;

    ModelMetadata model_metadata_;
    InferenceConfig config_;
    uint64_t handle_ = 0;
    struct Stats;
    std::unique_ptr<Stats> stats_;
};

}  // namespace model
}  // namespace turboinfer
