// Host control flow of `WhisperState::full` (reference src-tauri/src/whisper.rs:127-129),
// restructured as a batch of per-audio state machines so that independent windows share every
// GPU launch: one encode batch, then decoder rounds in which each live sequence contributes
// its next token row (or its prompt rows).  Semantics per audio are those of the reference
// path: language auto-detect, prompt assembly,
// temperature fallback, timestamp-driven seek, segment splitting.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>

#include "state.h"

namespace nobs {

namespace {

constexpr int kChunk = WHISPER_CHUNK_SIZE;
// Minimum audio (mel frames, 10 ms each) whisper_full still processes: whisper.cpp v1.7.x `delta_min` = 100 ms.  The reference app
// hands over every clip longer than 1600 samples (state.rs:749 'at least 0.1 second').
constexpr int kDeltaMin = 10;

struct Sequence {
    std::vector<whisper_token_data> tokens;
    int result_len = 0;
    double sum_logprobs_all = 0.0, sum_logprobs = -INFINITY, avg_logprobs = -INFINITY, entropy = 0.0, score = -INFINITY;
};

struct Dec {
    Sequence seq;
    int seek_delta = 0;
    bool failed = false, completed = false, has_ts = false;
    SampleResult res{};  // sampling result waiting to be applied at the next step
};

struct Candidate {
    int decoder_idx;
    int seek_delta;
    bool has_ts;
    Sequence seq;
};

enum class Phase { LangDetect, Window, Prefill, Step, Finished };

// The first fallback pass of a window, decoded speculatively NEXT TO pass 0 (see Driver::shadow_* below).
struct Shadow {
    bool on = false;      // exists for the current window
    bool done = false;    // finished (completed / failed / token budget): waiting for pass 0's verdict
    Dec dec;
    std::vector<whisper_token> prompt;
    int step = 0;
    int sample = -1;      // index of its sample in the lane's round in flight (-1: none)
    bool prefill = false; // that sample is the pass's first (sampled from the prompt)
    float no_speech_prob = 0.0f;
    std::mt19937 rng0;    // decoder 0's generator as it was before the shadow drew from it
};

struct Job {
    whisper_state* st = nullptr;
    whisper_full_params p{};
    int rc = 0;
    Phase phase = Phase::Window;
    std::string lang;
    bool need_lang = false;
    int seek_start = 0, seek_end = 0, seek = 0;
    std::vector<float> temps;
    int it = 0, n_decoders = 1, n_cur = 1, step = 0, n_max = 0;
    std::vector<whisper_token> prompt, prompt_init;
    Dec dec[WHISPER_MAX_DECODERS];
    int best = 0;
    int lane = 0;  // decode lane this audio belongs to
    // bookkeeping of the round in flight
    int first_sample = -1, n_samples = 0;
    std::vector<int> live;  // decoder index of each sample of this round
    // the window this audio is about to decode is still with the encoder: it joins the lane's rounds again at round enc_round
    bool enc_wait = false;
    long enc_ticket = 0;
    long enc_round = 0;
    bool speculate = false; // this audio may run a shadow pass (greedy, best_of 1, pass 0 is argmax, a fallback temperature exists)
    Shadow sh;
};

whisper_token_data make_token(const SampleResult& r, int id, float p, float plog, int default_tid, int token_beg) {
    whisper_token_data t{};
    t.id = id;
    t.tid = r.tid >= 0 ? r.tid : default_tid;
    t.p = p;
    t.plog = plog;
    t.pt = r.pt;
    t.ptsum = r.ptsum;
    t.t0 = t.t1 = t.t_dtw = -1;
    t.vlen = 0.0f;
    if (id >= token_beg) { t.tid = id; t.pt = p; }
    return t;
}

void sequence_score(const whisper_full_params& p, Sequence& s) {
    if (s.result_len == 0) return;
    double result = 0.0;
    for (int i = 0; i < s.result_len; ++i) result += s.tokens[i].plog;
    s.sum_logprobs = result;
    s.avg_logprobs = result / s.result_len;
    double penalty = s.result_len;
    if (p.length_penalty > 0.0f) penalty = pow((5.0 + penalty) / 6.0, p.length_penalty);
    s.score = result / penalty;
    // entropy of the last 32 token ids (repetition detector)
    int cnt = 0;
    std::map<whisper_token, int> counts;
    for (int i = std::max(0, s.result_len - 32); i < s.result_len; ++i) { counts[s.tokens[i].id]++; cnt++; }
    double entropy = 0.0;
    for (const auto& kv : counts) { const double q = kv.second / (double)cnt; entropy -= q * log(q); }
    s.entropy = entropy;
}

bool same_tokens(const Sequence& a, const Sequence& b) {
    if (a.tokens.size() != b.tokens.size()) return false;
    for (size_t i = a.tokens.size(); i-- > 0;) if (a.tokens[i].id != b.tokens[i].id) return false;
    return true;
}

class Driver {
    struct LaneState {           // host side of one engine decode lane
        std::vector<int> jobs;   // indices into jobs_
        std::vector<RowDesc> rows;
        std::vector<int> samp;
        std::vector<SampleParams> sp;
        std::vector<SampleResult> res;
        std::vector<KvCopy> kv_pairs_a, kv_pairs_b;
        std::vector<EncodeRequest> enc;        // windows of this lane's audios that are due for the encoder
        std::string error;                     // error text raised on this lane's thread (handed to the caller's thread at the end)
        std::vector<float> inject;             // scripted-logits test hook: replacement logits per sample of the round
        std::vector<unsigned char> inject_mask;
        bool inflight = false;
        long round = 0;                        // rounds queued on this lane so far
        bool stalled = false;                  // the last turn found no rows to decode: wait for every window that is with the encoder
    };

public:
    Driver(whisper_context* ctx) : ctx_(ctx), eng_(*ctx->engine), vocab_(ctx->model.vocab), hp_(ctx->model.hp) {}

    int run(whisper_state* const* states, int n, const whisper_full_params& params, const float* const* samples, const int* n_samples, int* rc,
            const char* const* initial_prompts) {
        const long launches0 = kernel_launch_count();
        eng_.stats = EngineStats();
        jobs_.assign(n, Job());
        // ---- mel for every audio in one launch
        std::vector<MelRequest> mel_reqs;
        for (int i = 0; i < n; ++i) {
            Job& j = jobs_[i];
            j.st = states[i];
            j.p = params;
            if (initial_prompts && initial_prompts[i]) j.p.initial_prompt = initial_prompts[i];   // per-audio prompt (chunk chaining under data parallelism)
            j.st->result_all.clear();
            j.st->stats = whisper_b200_stats{};
            j.st->encoded_seek = -1;
            if (n_samples[i] > 0) mel_reqs.push_back(MelRequest{samples[i], n_samples[i], &j.st->mel});
        }
        if (!eng_.compute_mel(mel_reqs)) return fail_all(rc, n, -2);
        // ---- per-audio setup
        int need_kv = 0;
        for (int i = 0; i < n; ++i) need_kv += setup_job(jobs_[i]);
        {
            int need_audio = 0, kv_missing = 0;
            for (auto& j : jobs_) {
                if (j.phase == Phase::Finished) continue;
                if (j.st->audio_slot < 0) ++need_audio;
                const int want = kv_needed(j);
                if ((int)j.st->kv_slots.size() < want) kv_missing += want - (int)j.st->kv_slots.size();
            }
            (void)need_kv;
            if (!eng_.reserve_slots(need_audio, kv_missing)) return fail_all(rc, n, -100);
            for (auto& j : jobs_) {
                if (j.phase == Phase::Finished) continue;
                if (!ensure_state_slots(ctx_, j.st, kv_needed(j))) return fail_all(rc, n, -100);
            }
        }
        // ---- decode lanes: contiguous blocks of live audios per lane (an audio never changes lane, so its KV
        // slots are only ever touched by one stream)
        {
            std::vector<int> live;
            for (int i = 0; i < n; ++i) if (jobs_[i].phase != Phase::Finished) live.push_back(i);
            n_lanes_ = std::max(1, std::min(eng_.n_lanes(), (int)live.size()));
            for (size_t k = 0; k < live.size(); ++k) jobs_[live[k]].lane = (int)(k * n_lanes_ / live.size());
            lanes_.assign(n_lanes_, LaneState());
            for (int i = 0; i < n; ++i) lanes_[jobs_[i].lane].jobs.push_back(i);
        }
        // ---- rounds: every lane alternates "collect results, advance its audios, queue the next round".  With several lanes each
        // lane is driven by its own host thread: a round is ~400 launches (~2 ms of host time), and a single thread that is busy
        // issuing one lane's round leaves every other lane that finishes meanwhile idle until it gets back to it.  The lanes share
        // nothing but the encoder (Engine::encode is serialised inside the engine) and read-only weights; audios never change lane.
        // The scripted-logits test hook calls back into the test in a fixed order: it keeps the single-threaded loop.
        static const bool threads_ok = [] { const char* v = getenv("NOBS_WHISPER_HOST_THREADS"); return !(v && *v == '0'); }();
        if (n_lanes_ > 1 && threads_ok && !ctx_->logits_hook) {
            std::atomic<int> failed{0};
            std::vector<std::thread> workers;
            for (int ln = 0; ln < n_lanes_; ++ln) {
                workers.emplace_back([this, ln, &failed] {
                    while (!failed.load(std::memory_order_relaxed)) {
                        bool active = false;
                        const int code = lane_step(ln, active);
                        if (code != 0) { int expected = 0; failed.compare_exchange_strong(expected, code); break; }
                        if (!active) break;
                    }
                });
            }
            for (auto& w : workers) w.join();
            if (failed.load()) return fail_all(rc, n, failed.load());
        } else {
            while (true) {
                bool active = false;
                for (int ln = 0; ln < n_lanes_; ++ln) {
                    const int code = lane_step(ln, active);
                    if (code != 0) return fail_all(rc, n, code);
                }
                if (!active) break;
            }
        }
        for (const LaneState& L : lanes_) if (!L.error.empty()) set_last_error(L.error);
        const long launches = kernel_launch_count() - launches0;
        for (int i = 0; i < n; ++i) {
            rc[i] = jobs_[i].rc;
            whisper_b200_stats& s = jobs_[i].st->stats;
            s.n_kernel_launches = launches;
            s.gpu_ms_mel = eng_.stats.ms_mel;
            s.gpu_ms_encode = eng_.stats.ms_encode;
            s.gpu_ms_decode = eng_.stats.ms_decode;
            s.gpu_ms_enc_gemm = eng_.stats.ms_enc_gemm;
            s.gpu_ms_enc_attn = eng_.stats.ms_enc_attn;
            s.n_enc_gemm = eng_.stats.n_enc_gemm;
            s.n_enc_attn = eng_.stats.n_enc_attn;
            s.gpu_ms_dec_cross = eng_.stats.ms_dec_cross;
            s.n_dec_cross = eng_.stats.n_dec_cross;
            s.dec_cross_bytes = eng_.stats.dec_cross_bytes;
        }
        return 0;
    }

private:
    // One turn of a lane: take the results of its round in flight (if any), advance its audios, encode the windows that
    // became due, and queue the next round.  Returns 0, or the whisper_full error code; sets `active` if a round was queued.
    int lane_step(int ln, bool& active) {
        LaneState& L = lanes_[ln];
        if (L.inflight) {
            L.inflight = false;
            if (!eng_.decode_collect(ln, L.res)) return -8;
            for (int i : L.jobs) consume(jobs_[i]);
            if (!L.kv_pairs_a.empty()) {
                if (!eng_.kv_copy(ln, L.kv_pairs_a)) return -8;
                if (!L.kv_pairs_b.empty() && !eng_.kv_copy(ln, L.kv_pairs_b)) return -8;
            }
        }
        while (true) {
            if (!encode_round(L)) return -6;
            L.rows.clear(); L.samp.clear(); L.sp.clear();
            L.kv_pairs_a.clear(); L.kv_pairs_b.clear();
            L.inject_mask.clear();
            bool any = false;
            for (int i : L.jobs) any |= emit_rows(jobs_[i]);
            if (any) {
                const bool scripted = !L.inject_mask.empty();
                if (!eng_.decode_submit(ln, L.rows, L.samp, L.sp, nullptr, scripted ? L.inject.data() : nullptr, scripted ? L.inject_mask.data() : nullptr))
                    return -8;
                L.inflight = true;
                L.round++;
                active = true;
                return 0;
            }
            bool pending = false;
            for (int i : L.jobs) pending |= (jobs_[i].phase == Phase::Window) || jobs_[i].enc_wait;
            if (!pending) return 0;  // otherwise: windows that still have to be encoded, or are with the encoder
            L.stalled = true;
        }
    }

    int fail_all(int* rc, int n, int code) {
        const std::string first_error = eng_.last_error();
        // other lanes may still have a round in flight: wait for them (results are dropped) so that no kernel writes the
        // states' KV slots after this call returns and no lane is left "busy" for the next call on the context
        for (size_t ln = 0; ln < lanes_.size(); ++ln) {
            if (!lanes_[ln].inflight) continue;
            std::vector<SampleResult> dropped;
            eng_.decode_collect((int)ln, dropped);
            lanes_[ln].inflight = false;
        }
        for (Job& j : jobs_) if (j.enc_wait) { eng_.encode_wait(j.enc_ticket); j.enc_wait = false; }   // the encoder reads the states' mel buffers
        set_last_error(first_error);
        for (int i = 0; i < n; ++i) rc[i] = (jobs_.size() > (size_t)i && jobs_[i].rc != 0) ? jobs_[i].rc : code;
        return code;
    }
    int kv_needed(const Job& j) const { return j.p.strategy == WHISPER_SAMPLING_BEAM_SEARCH ? 2 * j.n_decoders : j.n_decoders + (j.speculate ? 1 : 0); }

    // Everything whisper_full does before its main loop.  Returns the number of KV slots needed.
    int setup_job(Job& j) {
        whisper_state* st = j.st;
        whisper_full_params& p = j.p;
        const char* lang = p.language;
        j.need_lang = (lang == nullptr || !*lang || !strcmp(lang, "auto") || p.detect_language);
        if (!j.need_lang) j.lang = lang;
        j.seek_start = p.offset_ms / 10;
        j.seek_end = p.duration_ms == 0 ? st->mel.n_len_org : j.seek_start + p.duration_ms / 10;
        if (st->mel.raw == nullptr || j.seek_end < j.seek_start + kDeltaMin) {  // under 100 ms of audio: nothing to do
            j.phase = Phase::Finished;
            return 0;
        }
        if (p.temperature_inc > 0.0f) {
            for (float t = p.temperature; t < 1.0f + 1e-6f; t += p.temperature_inc) j.temps.push_back(t);
        } else {
            j.temps.push_back(p.temperature);
        }
        j.n_decoders = p.strategy == WHISPER_SAMPLING_GREEDY ? p.greedy.best_of : std::max(p.greedy.best_of, p.beam_search.beam_size);
        j.n_decoders = std::max(1, j.n_decoders);
        if (j.n_decoders > WHISPER_MAX_DECODERS) { j.rc = -4; j.phase = Phase::Finished; return 0; }
        // Speculative first fallback (shadow pass): with the reference's parameters (greedy, best_of 1, temperature ladder 0, 0.2, ...)
        // pass 0 is an argmax decode that never touches the decoder's generator, so pass 1 — should it be needed — starts from a
        // generator state that is known up front.  It is decoded in the same rounds as pass 0 from its own KV slot: if pass 0 is
        // accepted the shadow is dropped and the generator restored, otherwise pass 1 is already (partly) done.  Token for token
        // the result is the sequential one; a round carries up to two rows per audio that share one cross-KV panel stream.
        const char* spec_env = getenv("NOBS_WHISPER_SPECULATE");   // read per call: tests and A/B runs toggle it between calls
        const bool spec_ok = !(spec_env && *spec_env == '0');
        j.speculate = spec_ok && !ctx_->logits_hook && p.strategy == WHISPER_SAMPLING_GREEDY && p.greedy.best_of <= 1 && j.temps.size() >= 2 &&
                      j.temps[0] < 1e-6f && j.temps[1] > 1e-6f && j.n_decoders + 1 <= WHISPER_MAX_DECODERS;
        if (p.strategy == WHISPER_SAMPLING_BEAM_SEARCH && p.beam_search.beam_size > kMaxTopK) { j.rc = -4; j.phase = Phase::Finished; return 0; }
        if (p.audio_ctx > hp_.n_audio_ctx) { j.rc = -5; j.phase = Phase::Finished; return 0; }
        {
            // Parameters that change the transcript upstream and are not implemented here are refused, never silently ignored
            // (the reference sets none of them, whisper.rs:88-124).  Observer callbacks (new_segment / progress) do not change
            // results and stay ignored.
            const char* what = nullptr;
            if (p.audio_ctx != 0) what = "audio_ctx != 0";
            else if (p.vad) what = "vad";
            else if (p.suppress_nst) what = "suppress_nst";
            else if (p.suppress_regex && *p.suppress_regex) what = "suppress_regex";
            else if (p.logits_filter_callback) what = "logits_filter_callback";
            else if (p.encoder_begin_callback) what = "encoder_begin_callback";
            else if (p.abort_callback) what = "abort_callback";
            else if (p.grammar_rules && p.n_grammar_rules > 0) what = "grammar_rules";
            else if (p.tdrz_enable) what = "tdrz_enable";
            if (what) {
                set_last_error(std::string("whisper_full: parameter not supported by this engine: ") + what);
                j.rc = -100; j.phase = Phase::Finished;
                return 0;
            }
        }
        if (p.no_context) st->prompt_past.clear();
        {
            std::vector<whisper_token> pt;
            if (p.prompt_tokens && p.prompt_n_tokens > 0) pt.assign(p.prompt_tokens, p.prompt_tokens + p.prompt_n_tokens);
            else if (p.initial_prompt) pt = tokenize(vocab_, p.initial_prompt);
            if (!pt.empty()) {  // prepended to whatever context the state already carries
                st->prompt_past.insert(st->prompt_past.end(), pt.begin(), pt.end());
                std::rotate(st->prompt_past.begin(), st->prompt_past.end() - pt.size(), st->prompt_past.end());
            }
        }
        {
            const bool is_distil = hp_.n_text_layer == 2 && hp_.n_vocab != 51866;
            if (is_distil && !p.no_timestamps) p.no_timestamps = true;
        }
        j.n_max = hp_.n_text_ctx / 2 - 4;
        j.seek = j.seek_start;
        if (j.need_lang && vocab_.is_multilingual()) j.phase = Phase::LangDetect;
        else { j.phase = Phase::Window; if (!finish_lang(j)) return 0; }
        return kv_needed(j);
    }

    bool finish_lang(Job& j) {
        j.prompt_init = {vocab_.token_sot};
        if (vocab_.is_multilingual()) {
            const int lid = lang_id(j.lang.c_str());
            if (lid < 0) {
                j.rc = -3; j.phase = Phase::Finished;
                const std::string msg = "unknown language '" + j.lang + "'";
                if (j.lane >= 0 && j.lane < (int)lanes_.size()) lanes_[j.lane].error = msg;   // may run on a lane's thread: the error text is thread-local
                set_last_error(msg);
                return false;
            }
            j.st->lang_id = lid;
            j.prompt_init.push_back(vocab_.token_lang(lid));
            j.prompt_init.push_back(j.p.translate ? vocab_.token_translate : vocab_.token_transcribe);
        }
        if (j.p.no_timestamps) j.prompt_init.push_back(vocab_.token_not);
        return true;
    }

    // Encode every window that is due (language-detect windows included), as one batch.  The batch is queued on the encoder's own
    // stream; the lane does not stand still for it: audios that have rows keep decoding, and an audio whose window is with the
    // encoder joins again a fixed number of rounds later (kEncDelay; by then the ~10 ms a window takes have normally passed, and
    // if not the lane waits).  The join round depends on the audios' states only, never on timing, so rounds are composed the same
    // way in every run.  When no audio of the lane has rows (the first windows of a batch), the lane waits for the encoder at once.
    bool encode_round(LaneState& L) {
        static const int delay = [] { const char* v = getenv("NOBS_WHISPER_ENC_DELAY"); return (v && *v) ? std::max(0, atoi(v)) : 3; }();
        std::vector<EncodeRequest>& enc_ = L.enc;
        enc_.clear();
        std::vector<Job*> queued;
        for (int ji : L.jobs) {
            Job& j = jobs_[ji];
            if (j.phase == Phase::Window && j.seek + kDeltaMin >= j.seek_end) j.phase = Phase::Finished;  // under 100 ms left
            if (j.phase != Phase::Window && j.phase != Phase::LangDetect) continue;
            const int seek = j.phase == Phase::LangDetect ? 0 : j.seek;   // whisper_lang_auto_detect_with_state(ctx, state, 0, ...): always offset 0
            if (j.st->encoded_seek != seek) {
                enc_.push_back(EncodeRequest{&j.st->mel, seek, j.st->audio_slot});
                j.st->encoded_seek = seek;
                j.st->stats.n_windows++;
                queued.push_back(&j);
            }
        }
        if (!enc_.empty()) {
            long ticket = 0;
            const bool ok = eng_.encode_async(enc_, &ticket);
            for (Job* j : queued) { j->enc_wait = true; j->enc_ticket = ticket; j->enc_round = L.round + delay; }
            if (!ok) return false;
        }
        bool others_have_rows = false;
        for (int ji : L.jobs) {
            const Job& j = jobs_[ji];
            others_have_rows |= !j.enc_wait && (j.phase == Phase::Prefill || j.phase == Phase::Step);
        }
        for (int ji : L.jobs) {
            Job& j = jobs_[ji];
            if (j.enc_wait && (delay == 0 || L.stalled || !others_have_rows || L.round >= j.enc_round)) {
                if (!eng_.encode_wait(j.enc_ticket)) return false;
                j.enc_wait = false;
            }
            if (j.phase == Phase::Window && !j.enc_wait) {
                if (j.seek > j.seek_start && j.seek + 500 >= j.seek_end) j.st->prompt_past.clear();
                j.it = 0;
                j.best = 0;
                j.phase = Phase::Prefill;
            }
        }
        L.stalled = false;
        return true;
    }

    void fill_sample_params(const Job& j, const Dec& d, float t_cur, bool prefill, int decoder_idx, SampleParams& sp) {
        const auto& toks = d.seq.tokens;
        sp = SampleParams{};
        sp.temperature = t_cur;
        sp.is_initial = toks.empty();
        sp.last_was_ts = !toks.empty() && toks.back().id >= vocab_.token_beg;
        sp.penult_was_ts = toks.size() < 2 || toks[toks.size() - 2].id >= vocab_.token_beg;
        sp.has_ts = d.has_ts;
        sp.suppress_blank = j.p.suppress_blank;
        sp.no_timestamps = j.p.no_timestamps;
        sp.ts_initial_limit = hp_.n_vocab;
        if (sp.is_initial && j.p.max_initial_ts > 0.0f) {
            const float precision = float(kChunk) / hp_.n_audio_ctx;
            sp.ts_initial_limit = vocab_.token_beg + (int)std::round(j.p.max_initial_ts / precision) + 1;
        }
        sp.ts_min = vocab_.token_beg + d.seek_delta / 2;
        sp.want_nosp = prefill ? 1 : 0;
        if (j.p.strategy == WHISPER_SAMPLING_BEAM_SEARCH) {
            sp.mode = 2;
            sp.k = j.p.beam_search.beam_size;
        } else if (t_cur < 1e-6f) {
            sp.mode = 0;
        } else {
            sp.mode = 1;  // one draw from the decoder's own generator, exactly as the reference distribution consumes it
            sp.u = std::generate_canonical<double, 53>(j.st->rng[decoder_idx]);
        }
    }

    // Scripted-logits test hook (whisper_b200_set_logits_hook): called for the sample just appended to the lane's round.
    void offer_logits(LaneState& L, const Job& j, int step, int decoder) {
        if (!ctx_->logits_hook) return;
        const size_t nv = (size_t)hp_.n_vocab, i = L.samp.size() - 1;
        if (L.inject.size() < (i + 1) * nv) L.inject.resize((i + 1) * nv);
        L.inject_mask.resize(i + 1, 0);
        L.inject_mask[i] = ctx_->logits_hook(ctx_->logits_hook_user, j.seek, j.it, step, decoder, (int)j.prompt.size(), hp_.n_vocab, L.inject.data() + i * nv) ? 1 : 0;
    }

    // Append this job's rows for the next decoder round.  Returns false if it has none.
    bool emit_rows(Job& j) {
        LaneState& L = lanes_[j.lane];
        std::vector<RowDesc>& rows_ = L.rows;
        std::vector<int>& samp_ = L.samp;
        std::vector<SampleParams>& sp_ = L.sp;
        j.first_sample = (int)samp_.size();
        j.n_samples = 0;
        j.live.clear();
        whisper_state* st = j.st;
        if (j.enc_wait) return false;   // its window is still with the encoder (encode_round)
        if (j.phase == Phase::LangDetect) {
            rows_.push_back(RowDesc{vocab_.token_sot, 0, st->kv_slots[0], st->audio_slot});
            samp_.push_back((int)rows_.size() - 1);
            SampleParams sp{};
            sp.ts_initial_limit = hp_.n_vocab;
            sp.mode = 0;
            sp_.push_back(sp);
            j.n_samples = 1;
            st->stats.n_decode_rows += 1;
            st->stats.n_sample_rows += 1;
            st->stats.n_decode_rounds++;
            return true;
        }
        if (j.phase == Phase::Prefill) {
            const float t_cur = j.temps[j.it];
            j.n_cur = 1;
            if (j.p.strategy == WHISPER_SAMPLING_GREEDY) { if (t_cur > 0.0f) j.n_cur = j.p.greedy.best_of; }
            else { j.n_cur = t_cur > 0.0f ? j.p.greedy.best_of : j.p.beam_search.beam_size; }
            j.n_cur = std::max(1, j.n_cur);
            for (int k = 0; k < j.n_cur; ++k) {
                Dec& d = j.dec[k];
                d.seq = Sequence();
                d.seek_delta = 100 * kChunk;
                d.failed = d.completed = d.has_ts = false;
            }
            j.prompt.clear();
            if (!st->prompt_past.empty() && t_cur < 0.5f && j.p.n_max_text_ctx > 0) {
                const int n_take = std::min(std::min(j.p.n_max_text_ctx, hp_.n_text_ctx / 2), (int)st->prompt_past.size());
                j.prompt.push_back(vocab_.token_prev);
                j.prompt.insert(j.prompt.end(), st->prompt_past.end() - n_take, st->prompt_past.end());
            }
            j.prompt.insert(j.prompt.end(), j.prompt_init.begin(), j.prompt_init.end());
            for (size_t i = 0; i < j.prompt.size(); ++i) rows_.push_back(RowDesc{j.prompt[i], (int)i, st->kv_slots[0], st->audio_slot});
            const int last = (int)rows_.size() - 1;
            for (int k = 0; k < j.n_cur; ++k) {  // every decoder samples its first token from the prompt's logits
                samp_.push_back(last);
                SampleParams sp;
                fill_sample_params(j, j.dec[k], t_cur, /*prefill=*/k == 0, k, sp);
                sp_.push_back(sp);
                j.live.push_back(k);
                offer_logits(L, j, /*step=*/0, k);
            }
            j.n_samples = j.n_cur;
            j.step = 0;
            st->stats.n_decode_rows += (int64_t)j.prompt.size();
            st->stats.n_sample_rows += j.n_cur;
            st->stats.n_decode_rounds++;
            j.sh.sample = -1;
            if (j.it == 0) { j.sh = Shadow(); if (j.speculate) shadow_prefill(j); }
            return true;
        }
        if (j.phase == Phase::Step) {
            const float t_cur = j.temps[j.it];
            const int n_past = (int)j.prompt.size() + j.step;
            for (int k = 0; k < j.n_cur; ++k) {
                Dec& d = j.dec[k];
                if (d.failed || d.completed) continue;
                rows_.push_back(RowDesc{d.seq.tokens.back().id, n_past, st->kv_slots[k], st->audio_slot});
                samp_.push_back((int)rows_.size() - 1);
                SampleParams sp;
                fill_sample_params(j, d, t_cur, false, k, sp);
                sp_.push_back(sp);
                j.live.push_back(k);
                offer_logits(L, j, j.step + 1, k);
            }
            j.n_samples = (int)j.live.size();
            st->stats.n_decode_rows += j.n_samples;
            st->stats.n_sample_rows += j.n_samples;
            st->stats.n_decode_rounds++;
            j.sh.sample = -1;
            if (j.it == 0 && j.sh.on && !j.sh.done && j.n_samples > 0) shadow_step(j);
            return j.n_samples > 0;
        }
        return false;
    }

    // ---- shadow pass (see setup_job): rows of temperature pass 1 queued next to pass 0's
    void shadow_prefill(Job& j) {
        LaneState& L = lanes_[j.lane];
        whisper_state* st = j.st;
        Shadow& sh = j.sh;
        const float t1 = j.temps[1];
        sh.on = true;
        sh.rng0 = st->rng[0];
        sh.dec.seq = Sequence();
        sh.dec.seek_delta = 100 * kChunk;
        sh.dec.failed = sh.dec.completed = sh.dec.has_ts = false;
        sh.prompt.clear();
        if (!st->prompt_past.empty() && t1 < 0.5f && j.p.n_max_text_ctx > 0) {
            const int n_take = std::min(std::min(j.p.n_max_text_ctx, hp_.n_text_ctx / 2), (int)st->prompt_past.size());
            sh.prompt.push_back(vocab_.token_prev);
            sh.prompt.insert(sh.prompt.end(), st->prompt_past.end() - n_take, st->prompt_past.end());
        }
        sh.prompt.insert(sh.prompt.end(), j.prompt_init.begin(), j.prompt_init.end());
        const int slot = st->kv_slots[j.n_decoders];
        for (size_t i = 0; i < sh.prompt.size(); ++i) L.rows.push_back(RowDesc{sh.prompt[i], (int)i, slot, st->audio_slot});
        L.samp.push_back((int)L.rows.size() - 1);
        SampleParams sp;
        fill_sample_params(j, sh.dec, t1, /*prefill=*/true, 0, sp);
        L.sp.push_back(sp);
        sh.sample = (int)L.samp.size() - 1;
        sh.prefill = true;
        sh.step = 0;
        st->stats.n_decode_rows += (int64_t)sh.prompt.size();
        st->stats.n_sample_rows += 1;
    }
    void shadow_step(Job& j) {
        LaneState& L = lanes_[j.lane];
        whisper_state* st = j.st;
        Shadow& sh = j.sh;
        L.rows.push_back(RowDesc{sh.dec.seq.tokens.back().id, (int)sh.prompt.size() + sh.step, st->kv_slots[j.n_decoders], st->audio_slot});
        L.samp.push_back((int)L.rows.size() - 1);
        SampleParams sp;
        fill_sample_params(j, sh.dec, j.temps[1], false, 0, sp);
        L.sp.push_back(sp);
        sh.sample = (int)L.samp.size() - 1;
        sh.prefill = false;
        st->stats.n_decode_rows += 1;
        st->stats.n_sample_rows += 1;
    }
    // apply the shadow's sample of the round just collected (before pass 0 is advanced: its verdict may adopt the shadow)
    void shadow_consume(Job& j) {
        Shadow& sh = j.sh;
        if (!sh.on || sh.sample < 0) return;
        sh.dec.res = lanes_[j.lane].res[sh.sample];
        sh.sample = -1;
        if (sh.prefill) { sh.no_speech_prob = sh.dec.res.no_speech_prob; sh.step = 0; } else sh.step += 1;
        Dec& d = sh.dec;
        d.seq.tokens.push_back(make_token(d.res, d.res.id, d.res.p, d.res.plog, 0, vocab_.token_beg));
        d.seq.sum_logprobs_all += d.res.plog;
        settle(j, d, sh.step);
        if (d.completed || d.failed || sh.step == j.n_max - 1) sh.done = true;
    }
    // pass 0 was rejected: the shadow becomes the current pass (temperature index 1)
    void shadow_adopt(Job& j) {
        whisper_state* st = j.st;
        Shadow& sh = j.sh;
        std::swap(st->kv_slots[0], st->kv_slots[j.n_decoders]);
        j.it = 1;
        j.n_cur = 1;
        j.dec[0] = sh.dec;
        j.prompt = sh.prompt;
        j.step = sh.step;
        st->no_speech_prob = sh.no_speech_prob;
        const bool done = sh.done;
        sh = Shadow();
        if (done) finish_temperature(j);
        else j.phase = Phase::Step;
    }

    // Take this job's results of the round and advance its state machine.
    void consume(Job& j) {
        if (j.n_samples == 0) return;
        whisper_state* st = j.st;
        LaneState& L = lanes_[j.lane];
        const std::vector<SampleResult>& res_ = L.res;
        std::vector<KvCopy>& kv_pairs_a_ = L.kv_pairs_a;
        if (j.phase == Phase::LangDetect) {
            int best = 0;
            if (!eng_.lang_probs(j.lane, j.first_sample, nullptr, &best)) { j.rc = -3; j.phase = Phase::Finished; return; }
            j.lang = lang_str(best);
            if (j.p.detect_language) { st->lang_id = best; j.phase = Phase::Finished; return; }
            j.phase = Phase::Window;
            finish_lang(j);
            return;
        }
        for (int s = 0; s < j.n_samples; ++s) j.dec[j.live[s]].res = res_[j.first_sample + s];
        shadow_consume(j);
        if (j.phase == Phase::Prefill) {
            st->no_speech_prob = res_[j.first_sample].no_speech_prob;
            for (int k = 1; k < j.n_cur; ++k) kv_pairs_a_.push_back(KvCopy{st->kv_slots[0], st->kv_slots[k], (int)j.prompt.size()});
            j.step = 0;
        } else {
            j.step += 1;
        }
        advance(j);
    }

    // The reference's per-decoder rules after token i has been appended: timestamp bookkeeping, completion, failure.
    void settle(const Job& j, Dec& d, int i) {
        int& result_len = d.seq.result_len;
        const whisper_token_data& tok = d.seq.tokens.back();
        if (tok.id > vocab_.token_beg) {
            const int seek_delta_new = 2 * (tok.id - vocab_.token_beg);
            if (d.has_ts && d.seek_delta > seek_delta_new && result_len < i) { d.failed = true; return; }  // went back in time
            d.seek_delta = seek_delta_new;
            result_len = i + 1;
            d.has_ts = true;
        }
        if (tok.id == vocab_.token_eot || (j.p.max_tokens > 0 && i >= j.p.max_tokens) ||
            (d.has_ts && j.seek + d.seek_delta + kDeltaMin >= j.seek_end)) {
            if (result_len == 0 && !j.p.no_timestamps) {
                if (j.seek + d.seek_delta + kDeltaMin >= j.seek_end) result_len = i + 1;
                else { d.failed = true; return; }
            }
            if (j.p.single_segment || j.p.no_timestamps) { result_len = i + 1; d.seek_delta = 100 * kChunk; }
            d.completed = true;
            return;
        }
        if (i == j.n_max - 1 && (result_len == 0 || d.seek_delta < 100 * kChunk / 2)) { d.failed = true; return; }  // repetition guard
    }

    // One iteration of the reference's token loop, from "sample" to "all decoders finished?".
    void advance(Job& j) {
        whisper_state* st = j.st;
        std::vector<KvCopy>& kv_pairs_a_ = lanes_[j.lane].kv_pairs_a;
        std::vector<KvCopy>& kv_pairs_b_ = lanes_[j.lane].kv_pairs_b;
        const int i = j.step;
        const bool beam = j.p.strategy == WHISPER_SAMPLING_BEAM_SEARCH;
        if (!beam) {
            for (int k = 0; k < j.n_cur; ++k) {
                Dec& d = j.dec[k];
                if (d.completed || d.failed) continue;
                d.seq.tokens.push_back(make_token(d.res, d.res.id, d.res.p, d.res.plog, 0, vocab_.token_beg));
                d.seq.sum_logprobs_all += d.res.plog;
            }
        } else {
            std::vector<Candidate> cands;
            for (int k = 0; k < j.n_cur; ++k) {
                Dec& d = j.dec[k];
                if (d.completed || d.failed) continue;
                for (int c = 0; c < d.res.n_topk; ++c) {
                    Candidate cd{k, d.seek_delta, d.has_ts, d.seq};
                    cd.seq.tokens.push_back(make_token(d.res, d.res.topk_id[c], d.res.topk_p[c], d.res.topk_plog[c], vocab_.token_beg, vocab_.token_beg));
                    cd.seq.sum_logprobs_all += d.res.topk_plog[c];
                    cands.push_back(std::move(cd));
                }
            }
            std::stable_sort(cands.begin(), cands.end(), [](const Candidate& a, const Candidate& b) {
                if (a.seq.sum_logprobs_all != b.seq.sum_logprobs_all) return a.seq.sum_logprobs_all > b.seq.sum_logprobs_all;
                if (a.decoder_idx != b.decoder_idx) return a.decoder_idx < b.decoder_idx;
                return a.seq.tokens.back().id < b.seq.tokens.back().id;
            });
            size_t cur_c = 0;
            const int n_pos = (int)j.prompt.size() + i;  // cache positions filled so far
            for (int k = 0; k < j.n_cur; ++k) {
                Dec& d = j.dec[k];
                if (d.completed || d.failed) continue;
                if (cands.empty()) { d.failed = true; continue; }
                if (cur_c >= cands.size()) cur_c = 0;
                const Candidate& cur = cands[cur_c++];
                while (cands.size() > cur_c && same_tokens(cands[cur_c].seq, cur.seq) && i > 0) ++cur_c;
                d.seek_delta = cur.seek_delta;
                d.has_ts = cur.has_ts;
                d.seq = cur.seq;
                // move the source decoder's KV history into this decoder's slot via a scratch slot.
                // At i == 0 every slot holds (a copy of) the same prompt history: nothing to move.
                if (i > 0 && cur.decoder_idx != k) {
                    kv_pairs_a_.push_back(KvCopy{st->kv_slots[cur.decoder_idx], st->kv_slots[j.n_decoders + k], n_pos});
                    kv_pairs_b_.push_back(KvCopy{st->kv_slots[j.n_decoders + k], st->kv_slots[k], n_pos});
                }
            }
        }
        // completion / failure / sliding window update
        for (int k = 0; k < j.n_cur; ++k) {
            Dec& d = j.dec[k];
            if (d.completed || d.failed) continue;
            settle(j, d, i);
        }
        bool all_done = true;
        for (int k = 0; k < j.n_cur; ++k) if (!(j.dec[k].completed || j.dec[k].failed)) all_done = false;
        if (all_done || i == j.n_max - 1) {
            finish_temperature(j);
        } else {
            j.phase = Phase::Step;
        }
    }

    void finish_temperature(Job& j) {
        whisper_state* st = j.st;
        double best_score = -INFINITY;
        for (int k = 0; k < j.n_cur; ++k) {
            Dec& d = j.dec[k];
            if (d.failed) continue;
            d.seq.tokens.resize(d.seq.result_len);
            sequence_score(j.p, d.seq);
            if (d.seq.result_len > 32 && d.seq.entropy < j.p.entropy_thold) { d.failed = true; continue; }
            if (best_score < d.seq.score) { best_score = d.seq.score; j.best = k; }
        }
        bool success = true;
        if (j.it != (int)j.temps.size() - 1) {
            const Dec& d = j.dec[j.best];
            if (d.failed || (d.seq.avg_logprobs < j.p.logprob_thold && st->no_speech_prob < j.p.no_speech_thold)) success = false;
        }
        if (!success) {
            st->stats.n_fallbacks++;
            if (j.it == 0 && j.sh.on) { shadow_adopt(j); return; }
            j.it += 1;
            j.phase = Phase::Prefill;
            return;
        }
        if (j.sh.on) {   // pass 0 accepted: as if the shadow had never drawn from the generator
            st->rng[0] = j.sh.rng0;
            j.sh = Shadow();
        }
        emit_segments(j);
        j.phase = Phase::Window;
    }

    void emit_segments(Job& j) {
        whisper_state* st = j.st;
        const Dec& best = j.dec[j.best];
        int seek_delta = best.seek_delta;
        const int result_len = best.seq.result_len;
        const auto& tc = best.seq.tokens;
        const int beg = vocab_.token_beg;
        const bool is_no_speech = st->no_speech_prob > j.p.no_speech_thold && best.seq.avg_logprobs < j.p.logprob_thold;
        auto& past = st->prompt_past;
        past.clear();
        if (!j.prompt.empty() && j.prompt.front() == vocab_.token_prev)
            past.insert(past.end(), j.prompt.begin() + 1, j.prompt.end() - j.prompt_init.size());
        for (int i = 0; i < result_len && !is_no_speech; ++i) past.push_back(tc[i].id);
        if (!tc.empty() && !is_no_speech) {
            int i0 = 0;
            int64_t t0 = j.seek + 2 * (tc.front().tid - beg);
            std::string text;
            for (int i = 0; i < (int)tc.size(); ++i) {
                if (j.p.print_special || tc[i].id < vocab_.token_eot) text += vocab_.id_to_token[tc[i].id];
                if (tc[i].id > beg && !j.p.single_segment) {
                    const int64_t t1 = j.seek + 2 * (tc[i].tid - beg);
                    if (!text.empty()) {
                        Segment s;
                        s.t0 = t0; s.t1 = t1; s.text = text; s.no_speech_prob = st->no_speech_prob;
                        for (int q = i0; q <= i; ++q) s.tokens.push_back(tc[q]);
                        st->result_all.push_back(std::move(s));
                    }
                    text.clear();
                    while (i < (int)tc.size() && tc[i].id > beg) ++i;
                    --i;
                    t0 = t1;
                    i0 = i + 1;
                }
            }
            if (!text.empty()) {
                Segment s;
                s.t0 = t0; s.t1 = j.seek + seek_delta; s.text = text; s.no_speech_prob = st->no_speech_prob;
                for (int q = i0; q < (int)tc.size(); ++q) s.tokens.push_back(tc[q]);
                st->result_all.push_back(std::move(s));
            }
        }
        const bool single_ts_ending = tc.size() > 1 && tc[tc.size() - 2].id < beg && tc[tc.size() - 1].id > beg;
        if (single_ts_ending) seek_delta = std::min(j.seek_end - j.seek, kChunk * 100);
        j.seek += seek_delta;
    }

    whisper_context* ctx_;
    Engine& eng_;
    const Vocab& vocab_;
    const HParams& hp_;
    std::vector<Job> jobs_;
    std::vector<LaneState> lanes_;
    int n_lanes_ = 1;
};

}  // namespace

bool ensure_state_slots(whisper_context* ctx, whisper_state* st, int n_kv) {
    Engine& e = *ctx->engine;
    if (st->audio_slot < 0) {
        st->audio_slot = e.acquire_audio_slot();
        st->encoded_seek = -1;
        if (st->audio_slot < 0) { set_last_error(e.last_error()); return false; }
    }
    while ((int)st->kv_slots.size() < n_kv) {
        const int s = e.acquire_kv_slot();
        if (s < 0) { set_last_error(e.last_error()); return false; }
        st->kv_slots.push_back(s);
    }
    return true;
}

int full_batch(whisper_context* ctx, whisper_state* const* states, int n, const whisper_full_params& params, const float* const* samples,
               const int* n_samples, int* rc, const char* const* initial_prompts) {
    if (!ctx || !ctx->engine || n <= 0) return -100;
    std::lock_guard<std::mutex> lock(ctx->engine->mu);
    Driver drv(ctx);
    return drv.run(states, n, params, samples, n_samples, rc, initial_prompts);
}

}  // namespace nobs
