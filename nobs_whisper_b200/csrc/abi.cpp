// extern "C" surface of libnobswhisper_b200.so (declared in include/whisper_b200.h).
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include <memory>

#include "state.h"

namespace nobs {
static thread_local std::string g_last_error;
void set_last_error(const std::string& e) {
    if (!e.empty()) g_last_error = e;
}
}  // namespace nobs

using namespace nobs;

namespace {
Precision resolve_precision(int precision) {
    if (precision == WHISPER_B200_PRECISION_FP32) return Precision::FP32;
    if (precision == WHISPER_B200_PRECISION_BF16) return Precision::BF16;
    const char* e = getenv("NOBS_WHISPER_PRECISION");
    if (e && (!strcmp(e, "fp32") || !strcmp(e, "f32") || !strcmp(e, "FP32"))) return Precision::FP32;
    return Precision::BF16;
}
const Segment* seg(whisper_state* st, int i) {
    if (!st || i < 0 || i >= (int)st->result_all.size()) return nullptr;
    return &st->result_all[i];
}
}  // namespace

extern "C" {

struct whisper_context_params whisper_context_default_params(void) {
    whisper_context_params p{};
    p.use_gpu = true;
    p.flash_attn = false;
    p.gpu_device = 0;
    if (const char* dev = getenv("NOBS_WHISPER_DEVICE")) p.gpu_device = atoi(dev);  // one process per GPU: LOCAL_RANK
    p.dtw_token_timestamps = false;
    p.dtw_aheads_preset = WHISPER_AHEADS_NONE;
    p.dtw_n_top = -1;
    p.dtw_aheads.n_heads = 0;
    p.dtw_aheads.heads = nullptr;
    p.dtw_mem_size = 1024 * 1024 * 128;
    return p;
}

// No exception may cross the C ABI (the callers are Rust / ctypes): every entry point that can allocate or parse
// reports failure through its return value and whisper_b200_last_error().
struct whisper_context* whisper_b200_init_from_file(const char* path_model, struct whisper_context_params params, int precision) {
    if (!path_model) { set_last_error("null model path"); return nullptr; }
    std::unique_ptr<whisper_context> ctx;
    try {
        ctx.reset(new whisper_context());
        ctx->params = params;
        std::string err;
        if (!load_ggml_model(path_model, ctx->model, err)) {
            set_last_error("failed to load model: " + err);
            return nullptr;
        }
        ctx->engine.reset(Engine::create(ctx->model, params.gpu_device, resolve_precision(precision), err));
        if (!ctx->engine) {
            set_last_error("failed to initialise the GPU engine: " + err);
            return nullptr;
        }
        ctx->model.tensors.clear();  // weights now live in HBM
    } catch (const std::exception& e) {
        set_last_error(std::string("failed to load model: ") + e.what());
        return nullptr;
    } catch (...) {
        set_last_error("failed to load model: unknown exception");
        return nullptr;
    }
    return ctx.release();
}

struct whisper_context* whisper_b200_init_host_only(const char* path_model) {
    if (!path_model) { set_last_error("null model path"); return nullptr; }
    std::unique_ptr<whisper_context> ctx;
    try {
        ctx.reset(new whisper_context());
        ctx->params = whisper_context_default_params();
        std::string err;
        if (!load_ggml_model(path_model, ctx->model, err)) {
            set_last_error("failed to load model: " + err);
            return nullptr;
        }
        ctx->model.tensors.clear();
    } catch (const std::exception& e) {
        set_last_error(std::string("failed to load model: ") + e.what());
        return nullptr;
    } catch (...) {
        set_last_error("failed to load model: unknown exception");
        return nullptr;
    }
    return ctx.release();  // no engine: every compute entry point fails on this handle
}

struct whisper_context* whisper_init_from_file_with_params_no_state(const char* path_model, struct whisper_context_params params) {
    return whisper_b200_init_from_file(path_model, params, WHISPER_B200_PRECISION_DEFAULT);
}

void whisper_free(struct whisper_context* ctx) { delete ctx; }

int whisper_b200_precision(struct whisper_context* ctx) {
    return ctx && ctx->engine ? (int)ctx->engine->precision() : 0;
}

int whisper_b200_decode_lanes(struct whisper_context* ctx) { return ctx && ctx->engine ? ctx->engine->n_lanes() : 0; }

struct whisper_state* whisper_init_state(struct whisper_context* ctx) {
    if (!ctx || !ctx->engine) { set_last_error("null context"); return nullptr; }
    auto* st = new whisper_state();
    st->ctx = ctx;
    return st;
}

void whisper_free_state(struct whisper_state* st) {
    if (!st) return;
    if (st->ctx && st->ctx->engine) {
        Engine& e = *st->ctx->engine;
        std::lock_guard<std::mutex> lock(e.mu);
        e.release_audio_slot(st->audio_slot);
        for (int s : st->kv_slots) e.release_kv_slot(s);
        e.free_mel(st->mel);
    }
    delete st;
}

struct whisper_full_params whisper_full_default_params(enum whisper_sampling_strategy strategy) {
    whisper_full_params p{};
    p.strategy = strategy;
    p.n_threads = std::min(4, (int)std::max(1u, std::thread::hardware_concurrency()));
    p.n_max_text_ctx = 16384;
    p.offset_ms = 0;
    p.duration_ms = 0;
    p.translate = false;
    p.no_context = true;
    p.no_timestamps = false;
    p.single_segment = false;
    p.print_special = false;
    p.print_progress = true;
    p.print_realtime = false;
    p.print_timestamps = true;
    p.token_timestamps = false;
    p.thold_pt = 0.01f;
    p.thold_ptsum = 0.01f;
    p.max_len = 0;
    p.split_on_word = false;
    p.max_tokens = 0;
    p.debug_mode = false;
    p.audio_ctx = 0;
    p.tdrz_enable = false;
    p.suppress_regex = nullptr;
    p.initial_prompt = nullptr;
    p.prompt_tokens = nullptr;
    p.prompt_n_tokens = 0;
    p.language = "en";
    p.detect_language = false;
    p.suppress_blank = true;
    p.suppress_nst = false;
    p.temperature = 0.0f;
    p.max_initial_ts = 1.0f;
    p.length_penalty = -1.0f;
    p.temperature_inc = 0.2f;
    p.entropy_thold = 2.4f;
    p.logprob_thold = -1.0f;
    p.no_speech_thold = 0.6f;
    p.greedy.best_of = -1;
    p.beam_search.beam_size = -1;
    p.beam_search.patience = -1.0f;
    p.grammar_penalty = 100.0f;
    p.vad = false;
    p.vad_model_path = nullptr;
    p.vad_params.threshold = 0.5f;
    p.vad_params.min_speech_duration_ms = 250;
    p.vad_params.min_silence_duration_ms = 100;
    p.vad_params.max_speech_duration_s = FLT_MAX;
    p.vad_params.speech_pad_ms = 30;
    p.vad_params.samples_overlap = 0.1f;
    switch (strategy) {
        case WHISPER_SAMPLING_GREEDY: p.greedy.best_of = 5; break;
        case WHISPER_SAMPLING_BEAM_SEARCH: p.beam_search.beam_size = 5; break;
    }
    return p;
}

int whisper_b200_full_batch(struct whisper_context* ctx, struct whisper_state* const* states, int n, struct whisper_full_params params,
                            const float* const* samples, const int* n_samples, int* rc) {
    if (!ctx || !states || !samples || !n_samples || !rc || n <= 0) { set_last_error("bad arguments"); return -100; }
    if (!ctx->engine) { set_last_error("this context has no GPU engine (host-only handle); there is no CPU fallback"); return -100; }
    for (int i = 0; i < n; ++i) {
        if (!states[i] || states[i]->ctx != ctx || (n_samples[i] > 0 && !samples[i])) { set_last_error("bad state/audio in batch"); return -100; }
        for (int k = 0; k < i; ++k) if (states[k] == states[i]) { set_last_error("a state appears twice in the batch"); return -100; }
    }
    try {
        return full_batch(ctx, states, n, params, samples, n_samples, rc);
    } catch (const std::exception& e) {
        set_last_error(std::string("full: ") + e.what());
    } catch (...) {
        set_last_error("full: unknown exception");
    }
    for (int i = 0; i < n; ++i) rc[i] = -100;
    return -100;
}

int whisper_b200_full_batch_prompts(struct whisper_context* ctx, struct whisper_state* const* states, int n, struct whisper_full_params params,
                                    const char* const* initial_prompts, const float* const* samples, const int* n_samples, int* rc) {
    if (!ctx || !states || !samples || !n_samples || !rc || n <= 0) { set_last_error("bad arguments"); return -100; }
    if (!ctx->engine) { set_last_error("this context has no GPU engine (host-only handle); there is no CPU fallback"); return -100; }
    for (int i = 0; i < n; ++i) {
        if (!states[i] || states[i]->ctx != ctx || (n_samples[i] > 0 && !samples[i])) { set_last_error("bad state/audio in batch"); return -100; }
        for (int k = 0; k < i; ++k) if (states[k] == states[i]) { set_last_error("a state appears twice in the batch"); return -100; }
    }
    try {
        return full_batch(ctx, states, n, params, samples, n_samples, rc, initial_prompts);
    } catch (const std::exception& e) {
        set_last_error(std::string("full: ") + e.what());
    } catch (...) {
        set_last_error("full: unknown exception");
    }
    for (int i = 0; i < n; ++i) rc[i] = -100;
    return -100;
}

int whisper_full_with_state(struct whisper_context* ctx, struct whisper_state* state, struct whisper_full_params params,
                            const float* samples, int n_samples) {
    int rc = 0;
    const int r = whisper_b200_full_batch(ctx, &state, 1, params, &samples, &n_samples, &rc);
    return r != 0 ? r : rc;
}

int whisper_full_n_segments_from_state(struct whisper_state* st) { return st ? (int)st->result_all.size() : 0; }
const char* whisper_full_get_segment_text_from_state(struct whisper_state* st, int i) { const Segment* s = seg(st, i); return s ? s->text.c_str() : nullptr; }
int64_t whisper_full_get_segment_t0_from_state(struct whisper_state* st, int i) { const Segment* s = seg(st, i); return s ? s->t0 : 0; }
int64_t whisper_full_get_segment_t1_from_state(struct whisper_state* st, int i) { const Segment* s = seg(st, i); return s ? s->t1 : 0; }
bool whisper_full_get_segment_speaker_turn_next_from_state(struct whisper_state* st, int i) { const Segment* s = seg(st, i); return s && s->speaker_turn_next; }
float whisper_full_get_segment_no_speech_prob_from_state(struct whisper_state* st, int i) { const Segment* s = seg(st, i); return s ? s->no_speech_prob : 0.0f; }
int whisper_full_n_tokens_from_state(struct whisper_state* st, int i) { const Segment* s = seg(st, i); return s ? (int)s->tokens.size() : 0; }
whisper_token whisper_full_get_token_id_from_state(struct whisper_state* st, int i, int t) {
    const Segment* s = seg(st, i);
    return (s && t >= 0 && t < (int)s->tokens.size()) ? s->tokens[t].id : -1;
}
whisper_token_data whisper_full_get_token_data_from_state(struct whisper_state* st, int i, int t) {
    const Segment* s = seg(st, i);
    if (s && t >= 0 && t < (int)s->tokens.size()) return s->tokens[t];
    whisper_token_data z{};
    return z;
}
float whisper_full_get_token_p_from_state(struct whisper_state* st, int i, int t) {
    const Segment* s = seg(st, i);
    return (s && t >= 0 && t < (int)s->tokens.size()) ? s->tokens[t].p : 0.0f;
}
const char* whisper_full_get_token_text_from_state(struct whisper_context* ctx, struct whisper_state* st, int i, int t) {
    return whisper_token_to_str(ctx, whisper_full_get_token_id_from_state(st, i, t));
}
int whisper_full_lang_id_from_state(struct whisper_state* st) { return st ? st->lang_id : -1; }

int whisper_n_vocab(struct whisper_context* c) { return c->model.hp.n_vocab; }
int whisper_n_text_ctx(struct whisper_context* c) { return c->model.hp.n_text_ctx; }
int whisper_n_audio_ctx(struct whisper_context* c) { return c->model.hp.n_audio_ctx; }
int whisper_is_multilingual(struct whisper_context* c) { return c->model.vocab.is_multilingual() ? 1 : 0; }
int whisper_model_n_vocab(struct whisper_context* c) { return c->model.hp.n_vocab; }
int whisper_model_n_audio_ctx(struct whisper_context* c) { return c->model.hp.n_audio_ctx; }
int whisper_model_n_audio_state(struct whisper_context* c) { return c->model.hp.n_audio_state; }
int whisper_model_n_audio_head(struct whisper_context* c) { return c->model.hp.n_audio_head; }
int whisper_model_n_audio_layer(struct whisper_context* c) { return c->model.hp.n_audio_layer; }
int whisper_model_n_text_ctx(struct whisper_context* c) { return c->model.hp.n_text_ctx; }
int whisper_model_n_text_state(struct whisper_context* c) { return c->model.hp.n_text_state; }
int whisper_model_n_text_head(struct whisper_context* c) { return c->model.hp.n_text_head; }
int whisper_model_n_text_layer(struct whisper_context* c) { return c->model.hp.n_text_layer; }
int whisper_model_n_mels(struct whisper_context* c) { return c->model.hp.n_mels; }
int whisper_model_ftype(struct whisper_context* c) { return c->model.hp.ftype; }
int whisper_model_type(struct whisper_context* c) { return c->model.mtype; }
const char* whisper_token_to_str(struct whisper_context* c, whisper_token t) {
    if (!c || t < 0 || t >= (int)c->model.vocab.id_to_token.size()) return "";
    return c->model.vocab.id_to_token[t].c_str();
}
whisper_token whisper_token_eot(struct whisper_context* c) { return c->model.vocab.token_eot; }
whisper_token whisper_token_sot(struct whisper_context* c) { return c->model.vocab.token_sot; }
whisper_token whisper_token_solm(struct whisper_context* c) { return c->model.vocab.token_solm; }
whisper_token whisper_token_prev(struct whisper_context* c) { return c->model.vocab.token_prev; }
whisper_token whisper_token_nosp(struct whisper_context* c) { return c->model.vocab.token_nosp; }
whisper_token whisper_token_not(struct whisper_context* c) { return c->model.vocab.token_not; }
whisper_token whisper_token_beg(struct whisper_context* c) { return c->model.vocab.token_beg; }
whisper_token whisper_token_lang(struct whisper_context* c, int lang_id) { return c->model.vocab.token_lang(lang_id); }
whisper_token whisper_token_translate(struct whisper_context* c) { return c->model.vocab.token_translate; }
whisper_token whisper_token_transcribe(struct whisper_context* c) { return c->model.vocab.token_transcribe; }

int whisper_tokenize(struct whisper_context* c, const char* text, whisper_token* tokens, int n_max_tokens) {
    if (!c || !text) return 0;
    try {
        const auto res = tokenize(c->model.vocab, text);
        if (n_max_tokens < (int)res.size()) return -(int)res.size();
        for (size_t i = 0; i < res.size(); ++i) tokens[i] = res[i];
        return (int)res.size();
    } catch (...) {
        set_last_error("tokenize: out of memory");
        return 0;
    }
}
int whisper_token_count(struct whisper_context* c, const char* text) { return -whisper_tokenize(c, text, nullptr, 0); }
int whisper_lang_max_id(void) { return kNumLangs - 1; }
int whisper_lang_id(const char* lang) { return lang_id(lang); }
const char* whisper_lang_str(int id) { return lang_str(id); }
const char* whisper_lang_str_full(int id) { return lang_str_full(id); }
const char* whisper_print_system_info(void) { return "NOBS_WHISPER_B200 = 1 | CUDA sm_100a = 1 | CPU_FALLBACK = 0"; }
const char* whisper_version(void) { return "nobs-whisper-b200 0.1 (whisper.h 1.7.x surface)"; }

// ---------------------------------------------------------------- stage-level entry points
int whisper_pcm_to_mel_with_state(struct whisper_context* ctx, struct whisper_state* st, const float* samples, int n_samples, int) {
    if (!ctx || !ctx->engine || !st || !samples || n_samples <= 0) return -1;
    Engine& e = *ctx->engine;
    std::lock_guard<std::mutex> lock(e.mu);
    std::vector<MelRequest> r{MelRequest{samples, n_samples, &st->mel}};
    st->encoded_seek = -1;
    if (!e.compute_mel(r)) { set_last_error(e.last_error()); return -1; }
    return 0;
}
int whisper_n_len_from_state(struct whisper_state* st) { return st ? st->mel.n_len_org : 0; }

int whisper_encode_with_state(struct whisper_context* ctx, struct whisper_state* st, int offset, int) {
    if (!ctx || !ctx->engine || !st || !st->mel.raw) return -1;
    Engine& e = *ctx->engine;
    std::lock_guard<std::mutex> lock(e.mu);
    if (!ensure_state_slots(ctx, st, 1)) return -1;
    std::vector<EncodeRequest> r{EncodeRequest{&st->mel, offset, st->audio_slot}};
    if (!e.encode(r)) { set_last_error(e.last_error()); return -1; }
    st->encoded_seek = offset;
    return 0;
}

int whisper_decode_with_state(struct whisper_context* ctx, struct whisper_state* st, const whisper_token* tokens, int n_tokens, int n_past, int) {
    if (!ctx || !ctx->engine || !st || !tokens || n_tokens <= 0 || st->audio_slot < 0) return -1;
    Engine& e = *ctx->engine;
    std::lock_guard<std::mutex> lock(e.mu);
    if (!ensure_state_slots(ctx, st, 1)) return -1;
    std::vector<RowDesc> rows;
    for (int i = 0; i < n_tokens; ++i) rows.push_back(RowDesc{tokens[i], n_past + i, st->kv_slots[0], st->audio_slot});
    std::vector<int> samp{n_tokens - 1};
    SampleParams sp{};
    sp.ts_initial_limit = ctx->model.hp.n_vocab;
    std::vector<SampleParams> sps{sp};
    std::vector<SampleResult> res;
    st->logits.resize(ctx->model.hp.n_vocab);
    if (!e.decode(rows, samp, sps, res, st->logits.data())) { set_last_error(e.last_error()); return -1; }
    return 0;
}
int whisper_b200_decode_batch(struct whisper_context* ctx, struct whisper_state* const* states, int n, const whisper_token* tokens, const int* n_tokens,
                              const int* n_past, int lane, float* logits_out) {
    if (!ctx || !ctx->engine || !states || !tokens || !n_tokens || !n_past || !logits_out || n <= 0) { set_last_error("decode_batch: bad arguments"); return -1; }
    Engine& e = *ctx->engine;
    if (lane < 0 || lane >= e.n_lanes()) { set_last_error("decode_batch: no such lane"); return -1; }
    try {
        std::lock_guard<std::mutex> lock(e.mu);
        std::vector<RowDesc> rows;
        std::vector<int> samp;
        std::vector<SampleParams> sps;
        const whisper_token* t = tokens;
        for (int i = 0; i < n; ++i) {
            whisper_state* st = states[i];
            if (!st || st->ctx != ctx || st->audio_slot < 0 || n_tokens[i] <= 0 || n_past[i] < 0) { set_last_error("decode_batch: bad state / token count"); return -1; }
            if (!ensure_state_slots(ctx, st, 1)) return -1;
            for (int k = 0; k < n_tokens[i]; ++k) rows.push_back(RowDesc{t[k], n_past[i] + k, st->kv_slots[0], st->audio_slot});
            t += n_tokens[i];
            samp.push_back((int)rows.size() - 1);
            SampleParams sp{};
            sp.ts_initial_limit = ctx->model.hp.n_vocab;
            sps.push_back(sp);
        }
        std::vector<SampleResult> res;
        if (!e.decode_submit(lane, rows, samp, sps, logits_out) || !e.decode_collect(lane, res)) { set_last_error(e.last_error()); return -1; }
    } catch (const std::exception& ex) {
        set_last_error(std::string("decode_batch: ") + ex.what());
        return -1;
    }
    return 0;
}
void whisper_b200_set_logits_hook(struct whisper_context* ctx, whisper_b200_logits_hook hook, void* user) {
    if (!ctx) return;
    std::unique_lock<std::mutex> lock;
    if (ctx->engine) lock = std::unique_lock<std::mutex>(ctx->engine->mu);
    ctx->logits_hook = hook;
    ctx->logits_hook_user = user;
}
float* whisper_get_logits_from_state(struct whisper_state* st) { return st && !st->logits.empty() ? st->logits.data() : nullptr; }

int whisper_lang_auto_detect_with_state(struct whisper_context* ctx, struct whisper_state* st, int offset_ms, int n_threads, float* lang_probs) {
    if (!ctx || !ctx->engine || !st || !st->mel.raw) return -1;
    const int seek = offset_ms / 10;
    if (seek < 0 || seek >= st->mel.n_len_org) return -1;
    if (!ctx->model.vocab.is_multilingual()) return -2;
    if (whisper_encode_with_state(ctx, st, seek, n_threads) != 0) return -6;
    const whisper_token sot = ctx->model.vocab.token_sot;
    if (whisper_decode_with_state(ctx, st, &sot, 1, 0, n_threads) != 0) return -7;
    Engine& e = *ctx->engine;
    std::lock_guard<std::mutex> lock(e.mu);
    int best = -1;
    if (!e.lang_probs(0, 0, lang_probs, &best)) { set_last_error(e.last_error()); return -7; }
    return best;
}

// ---------------------------------------------------------------- inspection hooks
int whisper_b200_get_mel(struct whisper_state* st, float* out, size_t cap) {
    if (!st || !st->mel.raw || !st->ctx || !st->ctx->engine) return 0;
    const size_t need = (size_t)st->mel.n_len * st->mel.n_mel;
    if (!out || cap < need) return -(int)need;
    Engine& e = *st->ctx->engine;
    std::lock_guard<std::mutex> lock(e.mu);
    if (!e.export_mel(st->mel, out)) { set_last_error(e.last_error()); return 0; }
    return st->mel.n_len;
}
int whisper_b200_get_encoder_output(struct whisper_context* ctx, struct whisper_state* st, float* out, size_t cap) {
    if (!ctx || !ctx->engine || !st || st->audio_slot < 0) return -1;
    const size_t need = (size_t)ctx->model.hp.n_audio_ctx * ctx->model.hp.n_audio_state;
    if (!out || cap < need) return -(int)need;
    Engine& e = *ctx->engine;
    std::lock_guard<std::mutex> lock(e.mu);
    if (!e.export_encoder_output(st->audio_slot, out)) { set_last_error(e.last_error()); return -1; }
    return 0;
}
int whisper_b200_get_cross_kv(struct whisper_context* ctx, struct whisper_state* st, int layer, float* k, float* v, size_t cap) {
    if (!ctx || !ctx->engine || !st || st->audio_slot < 0 || !k || !v) return -1;
    const size_t need = (size_t)ctx->model.hp.n_audio_ctx * ctx->model.hp.n_text_state;
    if (cap < need) return -(int)need;
    Engine& e = *ctx->engine;
    std::lock_guard<std::mutex> lock(e.mu);
    if (!e.export_cross_kv(st->audio_slot, layer, k, v)) { set_last_error(e.last_error()); return -1; }
    return 0;
}

int whisper_b200_process_logits(struct whisper_context* ctx, struct whisper_full_params params, const float* logits, const whisper_token* hist,
                                int n_hist, int has_ts, int seek_delta, float temperature, int mode, double u, int k, float* logprobs_out,
                                float* probs_out, whisper_b200_sample_result* result) {
    if (!ctx || !ctx->engine || !logits || !result) { set_last_error("no GPU engine on this context"); return -1; }
    const Vocab& v = ctx->model.vocab;
    SampleParams sp{};
    sp.temperature = temperature;
    sp.is_initial = n_hist == 0;
    sp.last_was_ts = n_hist > 0 && hist[n_hist - 1] >= v.token_beg;
    sp.penult_was_ts = n_hist < 2 || hist[n_hist - 2] >= v.token_beg;
    sp.has_ts = has_ts;
    sp.suppress_blank = params.suppress_blank;
    sp.no_timestamps = params.no_timestamps;
    sp.ts_initial_limit = ctx->model.hp.n_vocab;
    if (sp.is_initial && params.max_initial_ts > 0.0f) {
        const float precision = float(WHISPER_CHUNK_SIZE) / ctx->model.hp.n_audio_ctx;
        sp.ts_initial_limit = v.token_beg + (int)std::round(params.max_initial_ts / precision) + 1;
    }
    sp.ts_min = v.token_beg + seek_delta / 2;
    sp.mode = mode;
    sp.k = k;
    sp.u = u;
    sp.want_nosp = 1;
    SampleResult r{};
    Engine& e = *ctx->engine;
    std::lock_guard<std::mutex> lock(e.mu);
    if (!e.process_logits_host(logits, sp, r, logprobs_out, probs_out)) { set_last_error(e.last_error()); return -1; }
    result->id = r.id; result->tid = r.tid; result->p = r.p; result->plog = r.plog; result->pt = r.pt; result->ptsum = r.ptsum;
    result->no_speech_prob = r.no_speech_prob; result->n_topk = r.n_topk;
    for (int i = 0; i < WHISPER_MAX_DECODERS; ++i) { result->topk_id[i] = r.topk_id[i]; result->topk_plog[i] = r.topk_plog[i]; result->topk_p[i] = r.topk_p[i]; }
    return 0;
}

int whisper_b200_get_stats(struct whisper_state* st, whisper_b200_stats* out) {
    if (!st || !out) return -1;
    *out = st->stats;
    return 0;
}

void whisper_b200_set_profiling(struct whisper_context* ctx, int on) {
    if (ctx && ctx->engine) ctx->engine->profiling = on != 0;
}

int whisper_b200_event_record(struct whisper_context* ctx, int slot) {
    if (!ctx || !ctx->engine) return -1;
    std::lock_guard<std::mutex> lock(ctx->engine->mu);
    return ctx->engine->event_record(slot) ? 0 : -1;
}
double whisper_b200_event_elapsed_ms(struct whisper_context* ctx, int slot_a, int slot_b) {
    if (!ctx || !ctx->engine) return -1.0;
    std::lock_guard<std::mutex> lock(ctx->engine->mu);
    return ctx->engine->event_elapsed_ms(slot_a, slot_b);
}

int whisper_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char* whisper_b200_last_error(void) { return g_last_error.c_str(); }

}  // extern "C"
