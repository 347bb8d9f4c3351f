// C++ host-side mirror of the reference's engine wrapper (reference src-tauri/src/whisper.rs:16-260)
// written above the C ABI (include/whisper_b200.h).  The reference's host language is Rust; no
// Rust toolchain exists in this image, so the wrapper logic — prompt assembly, the fixed
// parameter block, segment concatenation, hallucination filter, chunk chaining — is restated
// here with the same names, argument meaning and error behaviour.  The equivalent Rust FFI
// crate source is in rust/ and INTEGRATION.md.
#pragma once
#include <optional>
#include <string>
#include <vector>

#include "../../../include/whisper_b200.h"

namespace nobs {

// reference whisper.rs:5-14  enum WhisperError { LoadError(String), TranscriptionError(String), NoModel }
struct WhisperError {
    enum Kind { None = 0, LoadError = 1, TranscriptionError = 2, NoModel = 3 } kind = None;
    std::string message;
    std::string to_string() const;  // same wording as the reference's #[error(...)] strings
};

// reference whisper.rs:233-260 (+ phrase list :202-230)
std::string filter_hallucinations(const std::string& text);

class WhisperEngine {
public:
    WhisperEngine() = default;                      // whisper.rs:22-27  new()
    ~WhisperEngine();
    WhisperEngine(const WhisperEngine&) = delete;
    WhisperEngine& operator=(const WhisperEngine&) = delete;

    WhisperError load_model(const std::string& model_path);   // whisper.rs:36-52
    void unload_model();                                      // whisper.rs:55-59
    bool is_loaded() const { return ctx_ != nullptr; }         // whisper.rs:62-64

    // whisper.rs:66-148.  nullopt == Rust None.
    WhisperError transcribe(const float* audio, int n, const std::optional<std::string>& language,
                            const std::optional<std::string>& vocabulary, const std::optional<std::string>& context, std::string& out) const;
    // whisper.rs:152-197: sequential, previous non-empty text becomes the next chunk's context
    WhisperError transcribe_chunked(const std::vector<std::vector<float>>& chunks, const std::optional<std::string>& language,
                                    const std::optional<std::string>& vocabulary, std::string& out) const;
    // SURVEY.md §8f N4 — the same result as transcribe_chunked, with the chunks decoded data-parallel.  Chunk k's prompt needs the
    // text of the last non-empty chunk before it, which is only known once those chunks are decoded; so: decode a window of chunks
    // together, each with the context known so far (speculation), accept the longest prefix whose context turned out to be right,
    // re-decode the rest; the window halves when a round only confirmed one chunk and doubles when it confirmed all (window 1 is
    // exactly the sequential algorithm).  abort_on_error: whisper.rs:182-185 (true) or state.rs:773-775 (false: skip the chunk).
    struct ChainStats { int n_decodes = 0, n_rounds = 0; };
    WhisperError transcribe_chunked_parallel(const std::vector<const float*>& chunks, const std::vector<int>& n, const std::optional<std::string>& language,
                                             const std::optional<std::string>& vocabulary, bool abort_on_error, std::string& out, ChainStats* stats = nullptr) const;
    // B200 addition: independent windows decoded together, no context chaining.
    // beam_size > 0 selects BeamSearch{beam_size, patience:-1}, otherwise Greedy{best_of:1}.
    WhisperError transcribe_batch(const std::vector<const float*>& audios, const std::vector<int>& n, const std::optional<std::string>& language,
                                  const std::optional<std::string>& vocabulary, int beam_size, std::vector<std::string>& out) const;
    // one context per audio (nullopt: none); per-audio return codes instead of failing the whole batch when rc_out is given
    WhisperError transcribe_batch_contexts(const std::vector<const float*>& audios, const std::vector<int>& n, const std::optional<std::string>& language,
                                           const std::optional<std::string>& vocabulary, const std::vector<std::optional<std::string>>& contexts,
                                           std::vector<std::string>& out, std::vector<int>* rc_out) const;

    // state.rs:757-792: what the reference does with the audio left in the buffer when a recording stops — longer
    // than 30 s: cut at silences (audio.rs find_silence_boundaries + split_at_silences), then transcribe the pieces
    // in order, each with the previous non-empty text as context; a failing piece is skipped (state.rs:773-775),
    // results are joined with " " and trimmed.  parallel: 0 sequential (the reference's loop), 1 the pieces together without
    // context chaining, 2 together WITH the reference's chaining (transcribe_chunked_parallel: same text as 0).
    WhisperError transcribe_recording(const float* audio, size_t n, const std::optional<std::string>& language,
                                      const std::optional<std::string>& vocabulary, int parallel, std::string& out) const;

    whisper_context* raw_context() const { return ctx_; }
    // counters of the last transcribe / transcribe_batch call, summed over its audios
    const whisper_b200_stats& last_stats() const { return last_stats_; }

private:
    whisper_full_params make_params(const std::optional<std::string>& language, const std::string* initial_prompt, int beam_size) const;
    static std::optional<std::string> build_prompt(const std::optional<std::string>& vocabulary, const std::optional<std::string>& context);
    void accumulate_stats(whisper_state* st, bool first) const;
    whisper_context* ctx_ = nullptr;
    std::string model_path_;
    mutable whisper_b200_stats last_stats_{};
};

}  // namespace nobs
