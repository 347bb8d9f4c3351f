// Host-side model description: hyper-parameters, vocabulary, tokenizer, language table and
// the ggml `.bin` reader.  Serves WhisperContext::new_with_params (reference
// src-tauri/src/whisper.rs:36-52) for files named `ggml-<id>.bin` (reference lib.rs:29,
// config.rs:144, model.rs:50-188).  File layout: SURVEY.md §8a row a1.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace nobs {

struct HParams {
    int32_t n_vocab = 0, n_audio_ctx = 0, n_audio_state = 0, n_audio_head = 0, n_audio_layer = 0;
    int32_t n_text_ctx = 0, n_text_state = 0, n_text_head = 0, n_text_layer = 0, n_mels = 0, ftype = 0;
};

struct Vocab {
    int n_vocab = 0;
    std::vector<std::string> id_to_token;
    std::unordered_map<std::string, int> token_to_id;
    size_t max_token_len = 0;
    int token_eot = 50256, token_sot = 50257, token_translate = 50357, token_transcribe = 50358;
    int token_solm = 50359, token_prev = 50360, token_nosp = 50361, token_not = 50362, token_beg = 50363;
    int token_blank = -1;  // id of " " (suppress_blank), -1 if absent
    bool is_multilingual() const { return n_vocab >= 51865; }
    int num_languages() const { return n_vocab - 51765 - (is_multilingual() ? 1 : 0); }
    int token_lang(int lang_id) const { return token_sot + 1 + lang_id; }
};

struct HostTensor {
    std::vector<int64_t> shape;  // torch order (outermost first)
    std::vector<float> data;     // always widened to f32 on the host
    int ttype = 0;               // on-disk ggml type (0 f32, 1 f16, 2 q4_0, 3 q4_1, 6 q5_0, 7 q5_1, 8 q8_0)
};

struct HostModel {
    HParams hp;
    Vocab vocab;
    std::vector<float> filters;  // [n_mels][201]
    std::unordered_map<std::string, HostTensor> tensors;
    int mtype = 0;  // 1 tiny .. 5 large (by n_audio_layer), 0 unknown
    const HostTensor& get(const std::string& name) const;
};

// Returns false and fills `err` on any malformed / truncated / inconsistent file.
bool load_ggml_model(const std::string& path, HostModel& m, std::string& err);

// GPT-2 style word split + greedy longest match against the vocabulary (the tokenizer the
// reference path applies to initial_prompt, whisper.rs:98-109).
std::vector<int> tokenize(const Vocab& vocab, const std::string& text);

constexpr int kNumLangs = 100;
int lang_id(const char* lang);       // code ("en") or full name ("english"); -1 if unknown
const char* lang_str(int id);        // code
const char* lang_str_full(int id);   // full name

}  // namespace nobs
