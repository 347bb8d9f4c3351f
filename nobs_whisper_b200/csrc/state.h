// Internal definitions of the opaque C-ABI handles.
#pragma once
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../include/whisper_b200.h"
#include "engine.h"
#include "model.h"

struct whisper_context {
    nobs::HostModel model;  // tensors are dropped after upload; hparams / vocab / filters stay
    std::unique_ptr<nobs::Engine> engine;
    whisper_context_params params{};
    whisper_b200_logits_hook logits_hook = nullptr;   // scripted-logits test hook (whisper_b200_set_logits_hook)
    void* logits_hook_user = nullptr;
};

namespace nobs {
struct Segment {
    int64_t t0 = 0, t1 = 0;
    std::string text;
    float no_speech_prob = 0;
    std::vector<whisper_token_data> tokens;
    bool speaker_turn_next = false;
};
}  // namespace nobs

struct whisper_state {
    whisper_context* ctx = nullptr;
    nobs::DeviceMel mel;
    int audio_slot = -1;
    int encoded_seek = -1;          // window currently held in the audio slot (-1: none)
    std::vector<int> kv_slots;      // leased self-KV slots
    std::vector<nobs::Segment> result_all;
    std::vector<whisper_token> prompt_past;
    std::mt19937 rng[WHISPER_MAX_DECODERS];  // per-decoder generators, seeded 0 like the reference path
    float no_speech_prob = 0.0f;
    int lang_id = 0;
    std::vector<float> logits;      // last-token logits of the last whisper_decode_with_state
    whisper_b200_stats stats{};
    whisper_state() { for (auto& r : rng) r = std::mt19937(0); }
};

namespace nobs {
void set_last_error(const std::string& e);
// the batched `full` driver (full.cpp)
int full_batch(whisper_context* ctx, whisper_state* const* states, int n, const whisper_full_params& params, const float* const* samples,
               const int* n_samples, int* rc, const char* const* initial_prompts = nullptr);
bool ensure_state_slots(whisper_context* ctx, whisper_state* st, int n_kv);
}  // namespace nobs
