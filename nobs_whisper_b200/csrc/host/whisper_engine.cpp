#include "whisper_engine.h"

#include <algorithm>
#include <cstring>

namespace nobs {

std::string WhisperError::to_string() const {
    switch (kind) {
        case LoadError: return "Failed to load model: " + message;
        case TranscriptionError: return "Transcription failed: " + message;
        case NoModel: return "No model loaded";
        default: return "";
    }
}

namespace {

// reference whisper.rs:202-230 (data: the phrases whose exact match is discarded)
const char* const kHallucinationPhrases[] = {
    "thank you for watching", "thanks for watching", "thank you for listening", "thanks for listening",
    "subscribe to my channel", "please subscribe", "like and subscribe", "see you in the next video",
    "see you next time", "please like and subscribe", "don't forget to subscribe", "hit the bell",
    "leave a comment", "check out my other videos", "thanks for tuning in",
    "시청해 주셔서 감사합니다", "구독과 좋아요", "구독 부탁드립니다",
    "ご視聴ありがとうございました",
    "感谢收看", "谢谢观看",
    "you", "MBC 뉴스 이덕영입니다",
};

// decode one UTF-8 scalar starting at s[i]; returns its byte length (1 on malformed input)
size_t utf8_len(const std::string& s, size_t i) {
    const unsigned char c = (unsigned char)s[i];
    size_t n = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 1;
    if (i + n > s.size()) n = 1;
    return n;
}
bool is_ascii_punct(unsigned char c) { return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126); }
// Rust char::is_whitespace for the characters that can realistically appear (ASCII + common Unicode spaces)
size_t whitespace_len_at(const std::string& s, size_t i) {
    const unsigned char c = (unsigned char)s[i];
    if (c == ' ' || (c >= 9 && c <= 13)) return 1;
    if (c == 0xC2 && i + 1 < s.size() && ((unsigned char)s[i + 1] == 0x85 || (unsigned char)s[i + 1] == 0xA0)) return 2;
    if (c == 0xE2 && i + 2 < s.size()) {
        const unsigned char b1 = (unsigned char)s[i + 1], b2 = (unsigned char)s[i + 2];
        if (b1 == 0x80 && ((b2 >= 0x80 && b2 <= 0x8A) || b2 == 0xA8 || b2 == 0xA9 || b2 == 0xAF)) return 3;
        if (b1 == 0x81 && b2 == 0x9F) return 3;
    }
    if (c == 0xE3 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x80 && (unsigned char)s[i + 2] == 0x80) return 3;
    if (c == 0xE1 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x9A && (unsigned char)s[i + 2] == 0x80) return 3;
    return 0;
}
std::string trim(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e) { size_t l = whitespace_len_at(s, b); if (!l) break; b += l; }
    while (e > b) {
        // step back one scalar
        size_t p = e - 1;
        while (p > b && ((unsigned char)s[p] & 0xC0) == 0x80) --p;
        if (whitespace_len_at(s, p) == e - p && e - p > 0) e = p; else break;
    }
    return s.substr(b, e - b);
}
// symbols the reference treats like punctuation: '…' U+2026, '♪' U+266A, U+266B, U+266C
bool is_music_or_ellipsis(const std::string& s, size_t i, size_t n, bool with_beamed) {
    if (n != 3) return false;
    const unsigned char a = (unsigned char)s[i], b = (unsigned char)s[i + 1], c = (unsigned char)s[i + 2];
    if (a == 0xE2 && b == 0x80 && c == 0xA6) return true;               // …
    if (a == 0xE2 && b == 0x99 && c == 0xAA) return true;               // ♪
    if (with_beamed && a == 0xE2 && b == 0x99 && (c == 0xAB || c == 0xAC)) return true;
    return false;
}
// ASCII lower-casing plus the two-byte Latin-1/Greek/Cyrillic upper-case ranges is all the phrase list needs
// (the CJK / Hangul phrases have no case).
std::string to_lower(const std::string& s) {
    std::string o = s;
    for (auto& ch : o) if (ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
    return o;
}

}  // namespace

std::string filter_hallucinations(const std::string& text) {
    const std::string trimmed = trim(text);
    if (trimmed.empty()) return std::string();
    // punctuation / symbol only output ("...", "♪") is discarded   (whisper.rs:240-243)
    {
        bool all = true;
        for (size_t i = 0; i < trimmed.size();) {
            const size_t n = utf8_len(trimmed, i);
            if (!((n == 1 && is_ascii_punct((unsigned char)trimmed[i])) || is_music_or_ellipsis(trimmed, i, n, true))) { all = false; break; }
            i += n;
        }
        if (all) return std::string();
    }
    const std::string lower = to_lower(trimmed);
    // strip trailing punctuation (ASCII, '…', '♪') then compare with each phrase   (whisper.rs:248-257)
    size_t e = lower.size();
    while (e > 0) {
        size_t p = e - 1;
        while (p > 0 && ((unsigned char)lower[p] & 0xC0) == 0x80) --p;
        const size_t n = e - p;
        if ((n == 1 && is_ascii_punct((unsigned char)lower[p])) || is_music_or_ellipsis(lower, p, n, false)) e = p; else break;
    }
    const std::string stripped = lower.substr(0, e);
    for (const char* phrase : kHallucinationPhrases)
        if (stripped == to_lower(phrase)) return std::string();
    return trimmed;
}

WhisperEngine::~WhisperEngine() { unload_model(); }

WhisperError WhisperEngine::load_model(const std::string& model_path) {
    whisper_context_params cp = whisper_context_default_params();
    cp.use_gpu = true;  // whisper.rs:40
    whisper_context* c = whisper_init_from_file_with_params_no_state(model_path.c_str(), cp);
    if (!c) {
        const char* e = whisper_b200_last_error();
        return WhisperError{WhisperError::LoadError, (e && *e) ? e : "failed to initialise whisper context"};
    }
    unload_model();
    ctx_ = c;
    model_path_ = model_path;
    return WhisperError{};
}

void WhisperEngine::unload_model() {
    if (ctx_) whisper_free(ctx_);
    ctx_ = nullptr;
    model_path_.clear();
}

std::optional<std::string> WhisperEngine::build_prompt(const std::optional<std::string>& vocabulary, const std::optional<std::string>& context) {
    // whisper.rs:98-105
    if (vocabulary && context && !vocabulary->empty()) return *vocabulary + " " + *context;
    if (vocabulary && !context && !vocabulary->empty()) return *vocabulary;
    if (context) return *context;
    return std::nullopt;
}

whisper_full_params WhisperEngine::make_params(const std::optional<std::string>& language, const std::string* initial_prompt, int beam_size) const {
    whisper_full_params p;
    if (beam_size > 0) {
        p = whisper_full_default_params(WHISPER_SAMPLING_BEAM_SEARCH);
        p.beam_search.beam_size = beam_size;
        p.beam_search.patience = -1.0f;
    } else {
        p = whisper_full_default_params(WHISPER_SAMPLING_GREEDY);
        p.greedy.best_of = 1;  // whisper.rs:88
    }
    p.language = language ? language->c_str() : nullptr;                 // whisper.rs:91-95 (None => auto-detect)
    if (initial_prompt) p.initial_prompt = initial_prompt->c_str();      // whisper.rs:106-107
    p.print_special = false;                                             // whisper.rs:112-118
    p.print_progress = false;
    p.print_realtime = false;
    p.print_timestamps = false;
    p.translate = false;
    p.no_context = false;
    p.single_segment = false;
    p.suppress_blank = true;                                             // whisper.rs:121-124
    p.no_speech_thold = 0.6f;
    p.entropy_thold = 2.4f;
    p.logprob_thold = -1.0f;
    return p;
}

static std::string collect_text(whisper_state* st) {
    // whisper.rs:132-141: concatenate segment texts with no separator
    std::string result;
    const int n = whisper_full_n_segments_from_state(st);
    for (int i = 0; i < n; ++i) {
        const char* t = whisper_full_get_segment_text_from_state(st, i);
        if (t) result += t;
    }
    return result;
}

void WhisperEngine::accumulate_stats(whisper_state* st, bool first) const {
    whisper_b200_stats s{};
    if (whisper_b200_get_stats(st, &s) != 0) return;
    if (first) { last_stats_ = s; return; }
    last_stats_.n_windows += s.n_windows;
    last_stats_.n_decode_rows += s.n_decode_rows;
    last_stats_.n_sample_rows += s.n_sample_rows;
    last_stats_.n_fallbacks += s.n_fallbacks;
    last_stats_.n_decode_rounds = std::max(last_stats_.n_decode_rounds, s.n_decode_rounds);
}

WhisperError WhisperEngine::transcribe(const float* audio, int n, const std::optional<std::string>& language,
                                       const std::optional<std::string>& vocabulary, const std::optional<std::string>& context,
                                       std::string& out) const {
    out.clear();
    if (!ctx_) return WhisperError{WhisperError::NoModel, ""};                      // whisper.rs:73
    whisper_state* st = whisper_init_state(ctx_);                                    // whisper.rs:83-85
    if (!st) return WhisperError{WhisperError::TranscriptionError, whisper_b200_last_error()};
    const std::optional<std::string> prompt = build_prompt(vocabulary, context);
    const whisper_full_params p = make_params(language, prompt ? &*prompt : nullptr, 0);
    if (n <= 0 || !audio) {  // whisper-rs rejects an empty slice before the FFI call
        whisper_free_state(st);
        return WhisperError{WhisperError::TranscriptionError, "Input sample buffer was empty."};
    }
    const int rc = whisper_full_with_state(ctx_, st, p, audio, n);                   // whisper.rs:127-129
    if (rc != 0) {
        whisper_free_state(st);
        return WhisperError{WhisperError::TranscriptionError, "whisper_full_with_state returned " + std::to_string(rc) + ": " + whisper_b200_last_error()};
    }
    out = filter_hallucinations(trim(collect_text(st)));                             // whisper.rs:143-144
    accumulate_stats(st, true);
    whisper_free_state(st);
    return WhisperError{};
}

WhisperError WhisperEngine::transcribe_chunked(const std::vector<std::vector<float>>& chunks, const std::optional<std::string>& language,
                                               const std::optional<std::string>& vocabulary, std::string& out) const {
    out.clear();
    std::vector<std::string> results;
    std::optional<std::string> last_context;
    for (const auto& chunk : chunks) {
        std::string text;
        WhisperError e = transcribe(chunk.data(), (int)chunk.size(), language, vocabulary, last_context, text);
        if (e.kind != WhisperError::None) return e;  // abort on the first failing chunk (whisper.rs:182-185)
        if (!text.empty()) {
            last_context = text;
            results.push_back(text);
        }
    }
    for (size_t i = 0; i < results.size(); ++i) {
        if (i) out += " ";
        out += results[i];
    }
    return WhisperError{};
}

WhisperError WhisperEngine::transcribe_recording(const float* audio, size_t n, const std::optional<std::string>& language,
                                                 const std::optional<std::string>& vocabulary, int parallel, std::string& out) const {
    out.clear();
    if (!ctx_) return WhisperError{WhisperError::NoModel, ""};
    if (n <= 1600) return WhisperError{WhisperError::None, ""};   // state.rs:749: only audio longer than 0.1 s is transcribed
    std::vector<std::pair<size_t, size_t>> pieces;
    const size_t whisper_max_samples = 30 * 16000;   // state.rs:758
    if (n > whisper_max_samples) {
        std::vector<size_t> boundaries(n / 16000 + 2), ranges;
        size_t nb = 0, nc = 0;
        if (nobs_find_silence_boundaries(audio, n, 16000, boundaries.data(), boundaries.size(), &nb) != 0)
            return WhisperError{WhisperError::TranscriptionError, whisper_b200_last_error()};
        nb = std::min(nb, boundaries.size());
        ranges.resize(2 * (nb + 1));
        if (nobs_split_at_silences(n, boundaries.data(), nb, ranges.data(), &nc) != 0)
            return WhisperError{WhisperError::TranscriptionError, "split_at_silences failed"};
        for (size_t k = 0; k < nc; ++k) pieces.emplace_back(ranges[2 * k], ranges[2 * k + 1]);
    } else {
        pieces.emplace_back(0, n);
    }
    std::vector<std::string> results;
    if (parallel == 2 && pieces.size() > 1) {
        // the reference's loop (context chaining, failing pieces skipped), decoded data-parallel
        std::vector<const float*> ptrs;
        std::vector<int> lens;
        for (const auto& p : pieces) { ptrs.push_back(audio + p.first); lens.push_back((int)(p.second - p.first)); }
        const WhisperError e = transcribe_chunked_parallel(ptrs, lens, language, vocabulary, /*abort_on_error=*/false, out);
        if (e.kind != WhisperError::None) return e;
        const size_t b0 = out.find_first_not_of(" \t\r\n"), e0 = out.find_last_not_of(" \t\r\n");
        out = b0 == std::string::npos ? std::string() : out.substr(b0, e0 - b0 + 1);
        return WhisperError{};
    }
    if (parallel == 1 && pieces.size() > 1) {
        std::vector<const float*> ptrs;
        std::vector<int> lens;
        for (const auto& p : pieces) { ptrs.push_back(audio + p.first); lens.push_back((int)(p.second - p.first)); }
        std::vector<std::string> texts;
        WhisperError e = transcribe_batch(ptrs, lens, language, vocabulary, 0, texts);
        if (e.kind != WhisperError::None) return e;
        for (auto& t : texts) if (!t.empty()) results.push_back(std::move(t));
    } else {
        for (const auto& p : pieces) {
            std::string text;
            const std::optional<std::string> prev = results.empty() ? std::nullopt : std::optional<std::string>(results.back());
            const WhisperError e = transcribe(audio + p.first, (int)(p.second - p.first), language, vocabulary, prev, text);
            if (e.kind != WhisperError::None) continue;   // state.rs:773-775: log and go on
            if (!text.empty()) results.push_back(std::move(text));
        }
    }
    for (size_t i = 0; i < results.size(); ++i) {
        if (i) out += " ";
        out += results[i];
    }
    const size_t b = out.find_first_not_of(" \t\r\n"), e2 = out.find_last_not_of(" \t\r\n");   // .trim()
    out = b == std::string::npos ? std::string() : out.substr(b, e2 - b + 1);
    return WhisperError{};
}

WhisperError WhisperEngine::transcribe_batch_contexts(const std::vector<const float*>& audios, const std::vector<int>& n,
                                                      const std::optional<std::string>& language, const std::optional<std::string>& vocabulary,
                                                      const std::vector<std::optional<std::string>>& contexts, std::vector<std::string>& out,
                                                      std::vector<int>* rc_out) const {
    out.clear();
    if (!ctx_) return WhisperError{WhisperError::NoModel, ""};
    const int cnt = (int)audios.size();
    if (cnt == 0) return WhisperError{};
    if (n.size() != audios.size() || contexts.size() != audios.size()) return WhisperError{WhisperError::TranscriptionError, "audios / lengths / contexts size mismatch"};
    std::vector<whisper_state*> states(cnt, nullptr);
    auto free_states = [&]() { for (auto* s : states) whisper_free_state(s); };
    std::vector<std::optional<std::string>> prompts(cnt);
    std::vector<const char*> prompt_ptrs(cnt, nullptr);
    for (int i = 0; i < cnt; ++i) {
        if (!audios[i] || n[i] <= 0) { free_states(); return WhisperError{WhisperError::TranscriptionError, "Input sample buffer was empty."}; }
        states[i] = whisper_init_state(ctx_);
        if (!states[i]) { free_states(); return WhisperError{WhisperError::TranscriptionError, whisper_b200_last_error()}; }
        prompts[i] = build_prompt(vocabulary, contexts[i]);                           // whisper.rs:98-105, per chunk
        prompt_ptrs[i] = prompts[i] ? prompts[i]->c_str() : nullptr;
    }
    const whisper_full_params p = make_params(language, nullptr, 0);
    std::vector<int> rc(cnt, 0);
    const int r = whisper_b200_full_batch_prompts(ctx_, states.data(), cnt, p, prompt_ptrs.data(), audios.data(), n.data(), rc.data());
    if (r != 0 && !rc_out) {
        free_states();
        return WhisperError{WhisperError::TranscriptionError, "whisper_b200_full_batch returned " + std::to_string(r) + ": " + whisper_b200_last_error()};
    }
    if (!rc_out) {
        for (int i = 0; i < cnt; ++i)
            if (rc[i] != 0) {
                free_states();
                return WhisperError{WhisperError::TranscriptionError, "whisper_b200_full_batch returned " + std::to_string(rc[i]) + ": " + whisper_b200_last_error()};
            }
    } else {
        *rc_out = rc;
        if (r != 0) for (auto& v : *rc_out) if (v == 0) v = r;
    }
    out.resize(cnt);
    for (int i = 0; i < cnt; ++i) {
        if (rc[i] == 0 && r == 0) out[i] = filter_hallucinations(trim(collect_text(states[i])));
        accumulate_stats(states[i], i == 0);
    }
    free_states();
    return WhisperError{};
}

WhisperError WhisperEngine::transcribe_chunked_parallel(const std::vector<const float*>& chunks, const std::vector<int>& n,
                                                        const std::optional<std::string>& language, const std::optional<std::string>& vocabulary,
                                                        bool abort_on_error, std::string& out, ChainStats* stats) const {
    out.clear();
    if (!ctx_) return WhisperError{WhisperError::NoModel, ""};
    const int cnt = (int)chunks.size();
    if (n.size() != chunks.size()) return WhisperError{WhisperError::TranscriptionError, "chunks / lengths size mismatch"};
    ChainStats cs;
    std::vector<std::string> text(cnt);                       // current transcript of every chunk
    std::vector<int> status(cnt, -1);                         // -1 never decoded, 0 ok, > 0 failed
    std::vector<std::optional<std::string>> used(cnt);        // the context chunk k was last decoded with
    std::vector<std::string> err_text(cnt);
    int final_upto = 0;                                       // chunks [0, final_upto) carry their final text
    std::optional<std::string> final_context;                 // last non-empty text among the final chunks (whisper.rs:175-180)
    int window = cnt;
    while (final_upto < cnt) {
        // ---- speculate: the context of chunk k is the last non-empty text before it, as far as currently known (a chunk that was
        // never decoded counts as empty; a transcript made with a context that has since changed is still the best guess there is)
        const int hi = std::min(cnt, final_upto + std::max(1, window));
        std::vector<int> todo;
        std::vector<std::optional<std::string>> guess(cnt);
        {
            std::optional<std::string> c = final_context;
            for (int k = final_upto; k < hi; ++k) {
                guess[k] = c;                                          // chunk final_upto always gets its true context: every round confirms it
                if (status[k] < 0 || used[k] != c) todo.push_back(k);
                if (status[k] == 0 && !text[k].empty()) c = text[k];
            }
        }
        if (!todo.empty()) {
            std::vector<const float*> a;
            std::vector<int> ns;
            std::vector<std::optional<std::string>> ctxs;
            for (int k : todo) { a.push_back(chunks[k]); ns.push_back(n[k]); ctxs.push_back(guess[k]); }
            std::vector<std::string> texts;
            std::vector<int> rc;
            const WhisperError e = transcribe_batch_contexts(a, ns, language, vocabulary, ctxs, texts, &rc);
            if (e.kind != WhisperError::None) return e;
            cs.n_rounds += 1;
            cs.n_decodes += (int)todo.size();
            for (size_t i = 0; i < todo.size(); ++i) {
                const int k = todo[i];
                used[k] = guess[k];
                status[k] = rc[i] == 0 ? 0 : 1;
                text[k] = rc[i] == 0 ? texts[i] : std::string();
                if (rc[i] != 0) err_text[k] = "whisper_full_with_state returned " + std::to_string(rc[i]) + ": " + whisper_b200_last_error();
            }
        }
        // ---- accept the longest prefix whose context was right
        int confirmed = 0;
        while (final_upto < cnt && status[final_upto] >= 0 && used[final_upto] == final_context) {
            const int k = final_upto;
            if (status[k] > 0 && abort_on_error) return WhisperError{WhisperError::TranscriptionError, err_text[k]};
            if (status[k] == 0 && !text[k].empty()) final_context = text[k];
            ++final_upto;
            ++confirmed;
        }
        if (confirmed <= 1) window = std::max(1, window / 2);        // speculation past the first chunk was wasted: narrow it (1 = sequential)
        else window = std::min(cnt, window * 2);
    }
    std::vector<std::string> results;
    for (int k = 0; k < cnt; ++k) if (status[k] == 0 && !text[k].empty()) results.push_back(text[k]);
    for (size_t i = 0; i < results.size(); ++i) {
        if (i) out += " ";
        out += results[i];
    }
    if (stats) *stats = cs;
    return WhisperError{};
}

WhisperError WhisperEngine::transcribe_batch(const std::vector<const float*>& audios, const std::vector<int>& n,
                                             const std::optional<std::string>& language, const std::optional<std::string>& vocabulary,
                                             int beam_size, std::vector<std::string>& out) const {
    out.clear();
    if (!ctx_) return WhisperError{WhisperError::NoModel, ""};
    const int cnt = (int)audios.size();
    if (cnt == 0) return WhisperError{};
    if (n.size() != audios.size()) return WhisperError{WhisperError::TranscriptionError, "audios / lengths size mismatch"};
    std::vector<whisper_state*> states(cnt, nullptr);
    auto free_states = [&]() { for (auto* s : states) whisper_free_state(s); };
    for (int i = 0; i < cnt; ++i) {
        if (!audios[i] || n[i] <= 0) { free_states(); return WhisperError{WhisperError::TranscriptionError, "Input sample buffer was empty."}; }
        states[i] = whisper_init_state(ctx_);
        if (!states[i]) { free_states(); return WhisperError{WhisperError::TranscriptionError, whisper_b200_last_error()}; }
    }
    const std::optional<std::string> prompt = build_prompt(vocabulary, std::nullopt);
    const whisper_full_params p = make_params(language, prompt ? &*prompt : nullptr, beam_size);
    std::vector<int> rc(cnt, 0);
    const int r = whisper_b200_full_batch(ctx_, states.data(), cnt, p, audios.data(), n.data(), rc.data());
    int bad = r;
    for (int i = 0; i < cnt && bad == 0; ++i) bad = rc[i];
    if (bad != 0) {
        free_states();
        return WhisperError{WhisperError::TranscriptionError, "whisper_b200_full_batch returned " + std::to_string(bad) + ": " + whisper_b200_last_error()};
    }
    out.resize(cnt);
    for (int i = 0; i < cnt; ++i) {
        out[i] = filter_hallucinations(trim(collect_text(states[i])));
        accumulate_stats(states[i], i == 0);
    }
    free_states();
    return WhisperError{};
}

}  // namespace nobs

// ------------------------------------------------------------------------------------------
// C shims (declared in include/whisper_b200.h) so C / ctypes callers reach the same wrapper
// ------------------------------------------------------------------------------------------
struct nobs_engine {
    nobs::WhisperEngine engine;
    std::string last_error;
    std::string text;
    std::vector<std::string> texts;
};

extern "C" {

struct nobs_engine* nobs_engine_new(void) { return new nobs_engine(); }
void nobs_engine_free(struct nobs_engine* e) { delete e; }
int nobs_engine_load_model(struct nobs_engine* e, const char* path) {
    const nobs::WhisperError err = e->engine.load_model(path ? path : "");
    e->last_error = err.to_string();
    return (int)err.kind;
}
void nobs_engine_unload_model(struct nobs_engine* e) { e->engine.unload_model(); }
int nobs_engine_is_loaded(struct nobs_engine* e) { return e->engine.is_loaded() ? 1 : 0; }

static std::optional<std::string> opt(const char* s) { return s ? std::optional<std::string>(s) : std::nullopt; }

int nobs_engine_transcribe(struct nobs_engine* e, const float* audio, int n, const char* language, const char* vocabulary, const char* context,
                           const char** out) {
    const nobs::WhisperError err = e->engine.transcribe(audio, n, opt(language), opt(vocabulary), opt(context), e->text);
    e->last_error = err.to_string();
    if (out) *out = e->text.c_str();
    return (int)err.kind;
}
int nobs_engine_transcribe_chunked(struct nobs_engine* e, const float* const* chunks, const int* n, int n_chunks, const char* language,
                                   const char* vocabulary, const char** out) {
    std::vector<std::vector<float>> cs;
    for (int i = 0; i < n_chunks; ++i) cs.emplace_back(chunks[i], chunks[i] + std::max(0, n[i]));
    const nobs::WhisperError err = e->engine.transcribe_chunked(cs, opt(language), opt(vocabulary), e->text);
    e->last_error = err.to_string();
    if (out) *out = e->text.c_str();
    return (int)err.kind;
}
int nobs_engine_transcribe_recording(struct nobs_engine* e, const float* audio, size_t n, const char* language, const char* vocabulary, int parallel,
                                     const char** out) {
    const nobs::WhisperError err = e->engine.transcribe_recording(audio, n, opt(language), opt(vocabulary), parallel, e->text);
    e->last_error = err.to_string();
    if (out) *out = e->text.c_str();
    return (int)err.kind;
}
int nobs_engine_transcribe_chunked_parallel(struct nobs_engine* e, const float* const* chunks, const int* n, int n_chunks, const char* language,
                                            const char* vocabulary, int abort_on_error, const char** out, int* n_decodes, int* n_rounds) {
    std::vector<const float*> a(chunks, chunks + n_chunks);
    std::vector<int> ns(n, n + n_chunks);
    nobs::WhisperEngine::ChainStats cs;
    const nobs::WhisperError err = e->engine.transcribe_chunked_parallel(a, ns, opt(language), opt(vocabulary), abort_on_error != 0, e->text, &cs);
    e->last_error = err.to_string();
    if (out) *out = e->text.c_str();
    if (n_decodes) *n_decodes = cs.n_decodes;
    if (n_rounds) *n_rounds = cs.n_rounds;
    return (int)err.kind;
}
int nobs_engine_transcribe_batch(struct nobs_engine* e, const float* const* audios, const int* n, int n_audios, const char* language,
                                 const char* vocabulary, int beam_size, const char** texts) {
    std::vector<const float*> a(audios, audios + n_audios);
    std::vector<int> ns(n, n + n_audios);
    const nobs::WhisperError err = e->engine.transcribe_batch(a, ns, opt(language), opt(vocabulary), beam_size, e->texts);
    e->last_error = err.to_string();
    if (err.kind == nobs::WhisperError::None && texts)
        for (int i = 0; i < n_audios; ++i) texts[i] = e->texts[i].c_str();
    return (int)err.kind;
}
const char* nobs_engine_last_error(struct nobs_engine* e) { return e->last_error.c_str(); }
int nobs_engine_last_stats(struct nobs_engine* e, whisper_b200_stats* out) {
    if (!e || !out) return -1;
    *out = e->engine.last_stats();
    return 0;
}
struct whisper_context* nobs_engine_context(struct nobs_engine* e) { return e ? e->engine.raw_context() : nullptr; }
const char* nobs_filter_hallucinations(const char* text) {
    static thread_local std::string buf;
    buf = nobs::filter_hallucinations(text ? text : "");
    return buf.c_str();
}

}  // extern "C"
