/*
 * whisper_b200.h — C ABI of libnobswhisper_b200.so.
 *
 * Drop-in boundary for the path nobs-whisper reaches through whisper-rs (reference
 * src-tauri/src/whisper.rs:3, :36-52, :83-141).  whisper-rs 0.15.1 binds the whisper.h C API
 * through whisper-rs-sys 0.14.1 (reference src-tauri/Cargo.toml:27, src-tauri/Cargo.lock:5642-5660);
 * neither crate nor whisper.cpp is vendored in the reference tree, so the declarations below
 * restate the upstream whisper.h (v1.7.x) names, argument meaning and struct layouts that those
 * crates bind.  Every entry point states the reference call site it serves.
 *
 * Plain C, pointers and sizes only.  No torch / CUDA types appear in any signature.
 * There is no CPU fallback: every compute entry point fails (NULL / non-zero) when no
 * sm_100 device is usable.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WHISPER_SAMPLE_RATE 16000
#define WHISPER_N_FFT 400
#define WHISPER_HOP_LENGTH 160
#define WHISPER_CHUNK_SIZE 30
#define WHISPER_MAX_DECODERS 8

struct whisper_context;
struct whisper_state;
typedef int32_t whisper_token;
typedef int32_t whisper_pos;
typedef int32_t whisper_seq_id;

enum whisper_alignment_heads_preset { WHISPER_AHEADS_NONE = 0 };

typedef struct whisper_ahead {
    int n_text_layer;
    int n_head;
} whisper_ahead;

typedef struct whisper_aheads {
    size_t n_heads;
    const whisper_ahead* heads;
} whisper_aheads;

/* reference src-tauri/src/whisper.rs:39-40: WhisperContextParameters::default() + use_gpu(true).
 * use_gpu is accepted and ignored: this library has only the GPU path. */
struct whisper_context_params {
    bool use_gpu;
    bool flash_attn;
    int gpu_device; /* CUDA device ordinal */
    bool dtw_token_timestamps;
    enum whisper_alignment_heads_preset dtw_aheads_preset;
    int dtw_n_top;
    struct whisper_aheads dtw_aheads;
    size_t dtw_mem_size;
};

typedef struct whisper_token_data {
    whisper_token id;  /* token id */
    whisper_token tid; /* forced timestamp token id */
    float p;           /* probability of the token */
    float plog;        /* log probability of the token */
    float pt;          /* probability of the timestamp token */
    float ptsum;       /* sum of probabilities of all timestamp tokens */
    int64_t t0;        /* token-level timestamps (not produced: -1) */
    int64_t t1;
    int64_t t_dtw;
    float vlen;
} whisper_token_data;

enum whisper_sampling_strategy {
    WHISPER_SAMPLING_GREEDY = 0,      /* reference whisper.rs:88 SamplingStrategy::Greedy{best_of:1} */
    WHISPER_SAMPLING_BEAM_SEARCH = 1, /* BASELINE.json config 2: BeamSearch{beam_size:5, patience:-1} */
};

typedef void (*whisper_new_segment_callback)(struct whisper_context*, struct whisper_state*, int n_new, void* user_data);
typedef void (*whisper_progress_callback)(struct whisper_context*, struct whisper_state*, int progress, void* user_data);
typedef bool (*whisper_encoder_begin_callback)(struct whisper_context*, struct whisper_state*, void* user_data);
typedef bool (*whisper_abort_callback)(void* user_data);
typedef void (*whisper_logits_filter_callback)(struct whisper_context*, struct whisper_state*, const whisper_token_data* tokens,
                                               int n_tokens, float* logits, void* user_data);

typedef struct whisper_grammar_element {
    int type;
    uint32_t value;
} whisper_grammar_element;

typedef struct whisper_vad_params {
    float threshold;
    int min_speech_duration_ms;
    int min_silence_duration_ms;
    float max_speech_duration_s;
    int speech_pad_ms;
    float samples_overlap;
} whisper_vad_params;

/* The parameter block passed BY VALUE to whisper_full_with_state (upstream layout).
 * Fields the reference sets: whisper.rs:91-95 (language), :106-107 (initial_prompt), :112-118
 * (print_*, translate, no_context, single_segment), :121-124 (suppress_blank, no_speech_thold,
 * entropy_thold, logprob_thold).  Callbacks, grammar, DTW, VAD and tinydiarize are accepted and
 * ignored (the reference never sets them). */
struct whisper_full_params {
    enum whisper_sampling_strategy strategy;
    int n_threads;
    int n_max_text_ctx;
    int offset_ms;
    int duration_ms;
    bool translate;
    bool no_context;
    bool no_timestamps;
    bool single_segment;
    bool print_special;
    bool print_progress;
    bool print_realtime;
    bool print_timestamps;
    bool token_timestamps;
    float thold_pt;
    float thold_ptsum;
    int max_len;
    bool split_on_word;
    int max_tokens;
    bool debug_mode;
    int audio_ctx;
    bool tdrz_enable;
    const char* suppress_regex;
    const char* initial_prompt;
    const whisper_token* prompt_tokens;
    int prompt_n_tokens;
    const char* language;
    bool detect_language;
    bool suppress_blank;
    bool suppress_nst;
    float temperature;
    float max_initial_ts;
    float length_penalty;
    float temperature_inc;
    float entropy_thold;
    float logprob_thold;
    float no_speech_thold;
    struct {
        int best_of;
    } greedy;
    struct {
        int beam_size;
        float patience;
    } beam_search;
    whisper_new_segment_callback new_segment_callback;
    void* new_segment_callback_user_data;
    whisper_progress_callback progress_callback;
    void* progress_callback_user_data;
    whisper_encoder_begin_callback encoder_begin_callback;
    void* encoder_begin_callback_user_data;
    whisper_abort_callback abort_callback;
    void* abort_callback_user_data;
    whisper_logits_filter_callback logits_filter_callback;
    void* logits_filter_callback_user_data;
    const whisper_grammar_element** grammar_rules;
    size_t n_grammar_rules;
    size_t i_start_rule;
    float grammar_penalty;
    bool vad;
    const char* vad_model_path;
    whisper_vad_params vad_params;
};

/* ---------------------------------------------------------------- context (model) */

/* whisper.rs:39 WhisperContextParameters::default() */
struct whisper_context_params whisper_context_default_params(void);

/* whisper.rs:41-45 WhisperContext::new_with_params(path, params).  Parses `ggml-<id>.bin`
 * (reference lib.rs:29, config.rs:144, model.rs:50-188) and places the weights in HBM.
 * Returns NULL on failure (whisper-rs maps that to InitError -> WhisperError::LoadError). */
struct whisper_context* whisper_init_from_file_with_params_no_state(const char* path_model, struct whisper_context_params params);

/* Drop of the last Arc<WhisperEngine> holder (reference lib.rs:33, state.rs:177). */
void whisper_free(struct whisper_context* ctx);

/* ---------------------------------------------------------------- state */

/* whisper.rs:83-85 ctx.create_state().  States are cheap here: KV caches and activations
 * are leased from a per-context pool for the duration of a `full` call. */
struct whisper_state* whisper_init_state(struct whisper_context* ctx);
void whisper_free_state(struct whisper_state* state);

/* ---------------------------------------------------------------- full */

/* whisper.rs:88 FullParams::new(strategy) */
struct whisper_full_params whisper_full_default_params(enum whisper_sampling_strategy strategy);

/* whisper.rs:127-129 state.full(params, audio): mel -> [language detect] -> per window
 * encode / decode with temperature fallback -> segments.  0 = ok; non-zero = error
 * (whisper-rs maps -1 UnableToCalculateSpectrogram, 7 FailedToEncode, 8 FailedToDecode).
 * `samples` is borrowed for the call (host memory — pinned or pageable — or device memory: it is
 * read with cudaMemcpyDefault); language / initial_prompt are borrowed C strings. */
int whisper_full_with_state(struct whisper_context* ctx, struct whisper_state* state, struct whisper_full_params params,
                            const float* samples, int n_samples);

/* whisper.rs:132 state.full_n_segments() */
int whisper_full_n_segments_from_state(struct whisper_state* state);
/* whisper.rs:136-137 state.get_segment(i)?.to_str_lossy(): UTF-8 bytes owned by the state,
 * valid until the next whisper_full* on that state or whisper_free_state. */
const char* whisper_full_get_segment_text_from_state(struct whisper_state* state, int i_segment);
/* WhisperSegment accessors of whisper-rs 0.15 (not used by the reference, part of the API) */
int64_t whisper_full_get_segment_t0_from_state(struct whisper_state* state, int i_segment);
int64_t whisper_full_get_segment_t1_from_state(struct whisper_state* state, int i_segment);
bool whisper_full_get_segment_speaker_turn_next_from_state(struct whisper_state* state, int i_segment);
float whisper_full_get_segment_no_speech_prob_from_state(struct whisper_state* state, int i_segment);
int whisper_full_n_tokens_from_state(struct whisper_state* state, int i_segment);
const char* whisper_full_get_token_text_from_state(struct whisper_context* ctx, struct whisper_state* state, int i_segment, int i_token);
whisper_token whisper_full_get_token_id_from_state(struct whisper_state* state, int i_segment, int i_token);
whisper_token_data whisper_full_get_token_data_from_state(struct whisper_state* state, int i_segment, int i_token);
float whisper_full_get_token_p_from_state(struct whisper_state* state, int i_segment, int i_token);
int whisper_full_lang_id_from_state(struct whisper_state* state);

/* ---------------------------------------------------------------- model / vocab queries */
int whisper_n_vocab(struct whisper_context* ctx);
int whisper_n_text_ctx(struct whisper_context* ctx);
int whisper_n_audio_ctx(struct whisper_context* ctx);
int whisper_is_multilingual(struct whisper_context* ctx);
int whisper_model_n_vocab(struct whisper_context* ctx);
int whisper_model_n_audio_ctx(struct whisper_context* ctx);
int whisper_model_n_audio_state(struct whisper_context* ctx);
int whisper_model_n_audio_head(struct whisper_context* ctx);
int whisper_model_n_audio_layer(struct whisper_context* ctx);
int whisper_model_n_text_ctx(struct whisper_context* ctx);
int whisper_model_n_text_state(struct whisper_context* ctx);
int whisper_model_n_text_head(struct whisper_context* ctx);
int whisper_model_n_text_layer(struct whisper_context* ctx);
int whisper_model_n_mels(struct whisper_context* ctx);
int whisper_model_ftype(struct whisper_context* ctx);
int whisper_model_type(struct whisper_context* ctx);
const char* whisper_token_to_str(struct whisper_context* ctx, whisper_token token);
whisper_token whisper_token_eot(struct whisper_context* ctx);
whisper_token whisper_token_sot(struct whisper_context* ctx);
whisper_token whisper_token_solm(struct whisper_context* ctx);
whisper_token whisper_token_prev(struct whisper_context* ctx);
whisper_token whisper_token_nosp(struct whisper_context* ctx);
whisper_token whisper_token_not(struct whisper_context* ctx);
whisper_token whisper_token_beg(struct whisper_context* ctx);
whisper_token whisper_token_lang(struct whisper_context* ctx, int lang_id);
whisper_token whisper_token_translate(struct whisper_context* ctx);
whisper_token whisper_token_transcribe(struct whisper_context* ctx);
/* returns the number of tokens, or -(needed) when n_max_tokens is too small */
int whisper_tokenize(struct whisper_context* ctx, const char* text, whisper_token* tokens, int n_max_tokens);
int whisper_token_count(struct whisper_context* ctx, const char* text);
int whisper_lang_max_id(void);
int whisper_lang_id(const char* lang);
const char* whisper_lang_str(int id);
const char* whisper_lang_str_full(int id);
const char* whisper_print_system_info(void);
const char* whisper_version(void);

/* ---------------------------------------------------------------- stage-level entry points
 * (upstream whisper.h names; used by whisper-rs' lower-level API and by the parity tests) */
int whisper_pcm_to_mel_with_state(struct whisper_context* ctx, struct whisper_state* state, const float* samples, int n_samples, int n_threads);
int whisper_n_len_from_state(struct whisper_state* state); /* mel.n_len_org */
int whisper_encode_with_state(struct whisper_context* ctx, struct whisper_state* state, int offset, int n_threads);
int whisper_decode_with_state(struct whisper_context* ctx, struct whisper_state* state, const whisper_token* tokens, int n_tokens,
                              int n_past, int n_threads);
/* logits of the last token of the last decode call: n_vocab floats owned by the state */
float* whisper_get_logits_from_state(struct whisper_state* state);
int whisper_lang_auto_detect_with_state(struct whisper_context* ctx, struct whisper_state* state, int offset_ms, int n_threads,
                                        float* lang_probs);

/* ================================================================= B200 extensions
 * Not in upstream whisper.h.  The single-call entry points above keep upstream semantics; the
 * north star targets batches of independent windows, so these add batching, precision and
 * inspection.  */

enum whisper_b200_precision {
    WHISPER_B200_PRECISION_DEFAULT = 0, /* env NOBS_WHISPER_PRECISION (fp32|bf16), else bf16 */
    WHISPER_B200_PRECISION_FP32 = 1,    /* fp32 storage + fp32 CUDA-core math: the parity mode (1e-4) */
    WHISPER_B200_PRECISION_BF16 = 2,    /* bf16 weights/activations/KV, fp32 accumulate on tcgen05 */
};

/* Same as whisper_init_from_file_with_params_no_state with an explicit precision. */
struct whisper_context* whisper_b200_init_from_file(const char* path_model, struct whisper_context_params params, int precision);
int whisper_b200_precision(struct whisper_context* ctx);
/* Number of decode lanes of the context (independent stream + workspace pairs that batches are partitioned over;
 * env NOBS_WHISPER_LANES, default 2 in bf16 and 1 in fp32). */
int whisper_b200_decode_lanes(struct whisper_context* ctx);
/* Parses the file (hparams, vocabulary, mel filters) WITHOUT touching a GPU: a handle for the
 * vocabulary / tokenizer / model-query functions only.  Every compute entry point fails on it. */
struct whisper_context* whisper_b200_init_host_only(const char* path_model);

/* Batched `full`: n independent audios (windows / utterances), one state each, the same
 * params for all; windows are decoded together in lock step on ctx's GPU.  Equivalent to
 * calling whisper_full_with_state on each (state, audio) pair in turn; rc[i] receives the
 * per-audio return code.  Returns 0 if the batch ran (see rc[] for per-audio status). */
int whisper_b200_full_batch(struct whisper_context* ctx, struct whisper_state* const* states, int n, struct whisper_full_params params,
                            const float* const* samples, const int* n_samples, int* rc);

/* The same with one initial_prompt per audio (NULL entries: params.initial_prompt): what the reference's chunk chaining needs
 * when chunks are decoded together (whisper.rs:165-180 puts the previous chunk's text into the next chunk's prompt). */
int whisper_b200_full_batch_prompts(struct whisper_context* ctx, struct whisper_state* const* states, int n, struct whisper_full_params params,
                                    const char* const* initial_prompts, const float* const* samples, const int* n_samples, int* rc);

/* Batched whisper_decode_with_state: ONE decoder round over n states (each already encoded: whisper_encode_with_state).
 * State i contributes n_tokens[i] token rows (tokens are concatenated in `tokens`) at positions n_past[i].. ; the logits of its
 * last row are written to logits_out[i * n_vocab .. ] (fp32).  All rows run in the same launches (the decode-lane path `full`
 * uses, on lane `lane`), so this is the stage-level hook for teacher-forced parity checks of step batches.  0 on success. */
int whisper_b200_decode_batch(struct whisper_context* ctx, struct whisper_state* const* states, int n, const whisper_token* tokens,
                              const int* n_tokens, const int* n_past, int lane, float* logits_out);

/* Scripted-logits test hook (known-answer tests of the control flow, tests/test_control_flow_kat.py): while set, `whisper_full*`
 * offers the logits of every row it is about to sample from to `hook` BEFORE the decoder round is queued: seek of the window
 * (10 ms units), temperature index, step = index of the token about to be sampled (0: sampled from the prompt), decoder index,
 * number of prompt rows of this pass.  If hook returns non-zero, what it wrote to logits[n_vocab] replaces the model's logits of
 * that row on the device (the filter / log-softmax / sampling kernel then runs on them); 0 keeps the model's logits.  NULL removes
 * the hook.  Not meant for production use. */
typedef int (*whisper_b200_logits_hook)(void* user, int seek, int i_temp, int step, int decoder, int n_prompt, int n_vocab, float* logits);
void whisper_b200_set_logits_hook(struct whisper_context* ctx, whisper_b200_logits_hook hook, void* user);

/* Copy of the normalised log-mel of the last pcm_to_mel/full call, upstream layout
 * [n_mel][n_len]; returns n_len (or the needed float count negated if cap is too small). */
int whisper_b200_get_mel(struct whisper_state* state, float* out, size_t cap_floats);
/* Copy of the encoder output [n_audio_ctx][n_audio_state] (fp32) of the last encode. */
int whisper_b200_get_encoder_output(struct whisper_context* ctx, struct whisper_state* state, float* out, size_t cap_floats);
/* Copy of the cross-attention K / V of decoder layer `layer`, [n_audio_ctx][n_text_state] fp32 */
int whisper_b200_get_cross_kv(struct whisper_context* ctx, struct whisper_state* state, int layer, float* k_out, float* v_out, size_t cap_floats);

/* Filter + log-softmax + sample one row of logits on the GPU exactly as `full` does
 * (stage-level parity hook).  hist[n_hist] = tokens sampled so far in this window. */
typedef struct whisper_b200_sample_result {
    whisper_token id, tid;
    float p, plog, pt, ptsum, no_speech_prob;
    int n_topk;
    whisper_token topk_id[WHISPER_MAX_DECODERS];
    float topk_plog[WHISPER_MAX_DECODERS];
    float topk_p[WHISPER_MAX_DECODERS];
} whisper_b200_sample_result;
int whisper_b200_process_logits(struct whisper_context* ctx, struct whisper_full_params params, const float* logits,
                                const whisper_token* hist, int n_hist, int has_ts, int seek_delta, float temperature,
                                int mode /*0 argmax, 1 sample, 2 top-k*/, double u, int k, float* logprobs_out, float* probs_out,
                                whisper_b200_sample_result* result);

/* Counters of the last `full` on this state (for the benchmark's roofline arithmetic). */
typedef struct whisper_b200_stats {
    int64_t n_windows;        /* encoder passes */
    int64_t n_decode_rounds;  /* batched decoder forward calls this audio took part in */
    int64_t n_decode_rows;    /* token rows pushed through the decoder (prefill + steps) */
    int64_t n_sample_rows;    /* rows whose logits were computed and sampled */
    int64_t n_fallbacks;      /* temperature fallbacks */
    int64_t n_kernel_launches;/* kernels launched by the batch this audio was part of */
    double gpu_ms_mel, gpu_ms_encode, gpu_ms_decode; /* CUDA-event stage times of that batch */
    /* with profiling on: summed per-launch CUDA-event durations and launch counts of the encoder-side
     * GEMM kernel (conv stem, QKV/out/MLP, cross-KV projection) and of the encoder attention kernel */
    double gpu_ms_enc_gemm, gpu_ms_enc_attn;
    int64_t n_enc_gemm, n_enc_attn;
    /* decoder cross-attention kernel (the HBM-bound stream over the cross-KV panels): summed per-launch
     * durations, launches, and the algorithmic K/V bytes those launches had to read */
    double gpu_ms_dec_cross;
    int64_t n_dec_cross;
    double dec_cross_bytes;
} whisper_b200_stats;
/* Record a CUDA-event pair around every encoder GEMM / attention launch (costs ~1 us per launch). */
void whisper_b200_set_profiling(struct whisper_context* ctx, int on);
int whisper_b200_get_stats(struct whisper_state* state, whisper_b200_stats* out);

/* Kernel-level test hook: C[M][N] = epi(A * W^T) through the bf16 tcgen05 GEMM.  A is given as
 * a_elems fp32 values with row stride lda (lda < K gives the overlapping-row view the stem
 * convolutions use); inputs are rounded to bf16 on the device.  res: [res_mod ? res_mod : M][N]. */
int whisper_b200_debug_gemm_bf16(int M, int N, int K, int lda, const float* A, size_t a_elems, const float* W, const float* bias, int act,
                                 const float* res, int res_mod, int win_rows, int valid_rows, int out_f32, float* C_out);

/* Kernel-level test hook: encoder self-attention (non-causal, 1500 valid keys of 1536 rows per window,
 * head size 64).  qkv: [n_win*1536][3*64*n_head] fp32 (q|k|v), out: [n_win*1536][64*n_head] fp32. */
int whisper_b200_debug_enc_attention(int n_win, int n_head, const float* qkv, float* out, int use_simt);

/* Kernel-level test hook: decoder cross-attention of R single-token rows over head-major K/V panels
 * (k, v: [n_slots][n_head][1536][64] fp32, rounded to bf16 on the device; row r uses slot r % n_slots and
 * keys 0..n_keys-1).  streaming: 2 = tcgen05 streaming kernel, 1 = SIMT cp.async.bulk streaming kernel,
 * 0 = block-per-head SIMT kernel. */
int whisper_b200_debug_dec_cross_attention(int R, int n_head, int n_slots, int n_keys, const float* q, const float* k, const float* v, float* out,
                                           int streaming);
/* streaming 20 + g (g = 2..9): the tcgen05 kernel with ROW GROUPS — row r uses slot (r / g) % n_slots, i.e. runs of g consecutive rows per
 * audio; a run is cut into work items of up to 4 rows whose K / V panels are streamed once.
 * Host-only hook (no GPU needed): the grouping itself.  audio_slots[n_rows] -> groups[i] = first row | size << 24 for i < the number of
 * groups, which is left in groups[n_rows] (the array holds n_rows + 1 ints).  Returns the kernel's group width: 1 (no two neighbours share
 * a slot), 2 or 4. */
int whisper_b200_debug_cross_groups(const int* audio_slots, int n_rows, int* groups);

/* Micro-benchmark hook: average device microseconds per launch of the decoder-step kernels for R token
 * rows at model width d (see csrc/debug.cu for the index meaning of out_us[0..12]; out_us holds 16 floats). */
int whisper_b200_debug_time_decode_kernels(int R, int d, int iters, float* out_us);
/* Fused decoder projection (csrc/decode_proj_sm100.cu) on caller data: out [R][N] = in [R][K] * W[N][K]^T (+ bias, GELU); with resid the
 * fp32 sum resid + v is returned, with ln_g / ln_b also y_out = LayerNorm(out) (bf16 widened) and the per-tile (mean, M2) workspace
 * stats_out [N/128][128][2].  iters > 0 also times back-to-back launches. */
int whisper_b200_debug_dec_proj(int R, int N, int K, const float* in, const float* W, const float* bias, int act, const float* resid,
                                const float* ln_g, const float* ln_b, float* out, float* y_out, float* stats_out, int iters,
                                float* us_per_launch);
/* Micro-benchmark of the device-wide barrier of the fused projection chains: microseconds per barrier over `ctas` CTAs;
 * variant 0 fence + atomic + nanosleep polling, 1 without nanosleep, 2 acq_rel atomic / release increment / acquire polling;
 * store_floats fp32 stores per thread in front of every barrier. */
int whisper_b200_debug_grid_sync(int ctas, int iters, int variant, int store_floats, float* us_per_barrier);

/* ---- silence chunker (reference src-tauri/src/audio.rs) ----------------------------------
 * Cuts a long 16 kHz recording into the pieces the transcription path consumes (reference state.rs:757-780).
 * The per-window RMS runs on the current CUDA device (audio may be host or device memory); the results are
 * bit-identical to the reference's float32 arithmetic.  Return 0 on success, -1 bad arguments, -100 CUDA failure. */
/* audio.rs:364-370 for every full window [i*window, (i+1)*window): rms_out receives min(cap, *n_windows) values */
int whisper_b200_window_rms(const float* audio, size_t n_samples, uint32_t window, float* rms_out, size_t cap, size_t* n_windows);
/* audio.rs:400-463 find_silence_boundaries(audio, sample_rate): split points (sample indices, centre of each
 * silence gap >= 700 ms that leaves a chunk >= 1 s), adaptive threshold from the first 25 windows */
int nobs_find_silence_boundaries(const float* audio, size_t n_samples, uint32_t sample_rate, size_t* boundaries, size_t cap, size_t* n_found);
/* audio.rs:473-507 split_at_silences_with_overlap: chunk k = audio[ranges[2k], ranges[2k+1]) with 200 ms of the
 * previous chunk prepended; ranges must hold 2 * (n_boundaries + 1) entries */
int nobs_split_at_silences_with_overlap(size_t n_samples, const size_t* boundaries, size_t n_boundaries, uint32_t sample_rate, size_t* ranges,
                                        size_t* n_chunks);
/* audio.rs:467-469 split_at_silences (16 kHz) */
int nobs_split_at_silences(size_t n_samples, const size_t* boundaries, size_t n_boundaries, size_t* ranges, size_t* n_chunks);

/* ---- resampler (reference audio.rs:509-563 / :329-334, rubato FftFixedIn restated) ----------
 * The block FFT resampler as one fp32 GEMM per recording on the current CUDA device; audio / out may be host or
 * device memory.  *n_out = min(floor(n * to / from), complete blocks * block output); out == NULL only queries it. */
int nobs_resample_audio(const float* audio, size_t n, uint32_t from_rate, uint32_t to_rate, float* out, size_t cap, size_t* n_out);
/* audio.rs:329-334 resample_chunk: to 16 kHz, a plain copy when the input already is 16 kHz */
int nobs_resample_chunk(const float* audio, size_t n, uint32_t input_sample_rate, float* out, size_t cap, size_t* n_out);
/* state.rs:590-594: interleaved frames -> mono (sum of the channels / channels, float32); host */
int nobs_mix_to_mono(const float* interleaved, size_t n_frames, uint32_t channels, float* out);

/* Streaming capture buffer (reference audio.rs:29-244 `AudioBuffer`; used while recording, state.rs:586-605): host
 * logic, float32 arithmetic in the reference's order.  Chunk pointers stay valid until the next take_* / free. */
struct nobs_audio_buffer;
float nobs_calculate_rms(const float* samples, size_t n);                                   /* audio.rs:364-370 */
struct nobs_audio_buffer* nobs_audio_buffer_new(uint32_t sample_rate);                      /* with_sample_rate */
void nobs_audio_buffer_free(struct nobs_audio_buffer* b);
void nobs_audio_buffer_push_samples(struct nobs_audio_buffer* b, const float* samples, size_t n);
int nobs_audio_buffer_has_silence_boundary(const struct nobs_audio_buffer* b);
const float* nobs_audio_buffer_take_chunk_at_silence(struct nobs_audio_buffer* b, size_t* n);   /* NULL == None */
const float* nobs_audio_buffer_take_forced_chunk(struct nobs_audio_buffer* b, size_t* n);       /* NULL == None */
const float* nobs_audio_buffer_take(struct nobs_audio_buffer* b, size_t* n);
size_t nobs_audio_buffer_len(const struct nobs_audio_buffer* b);
size_t nobs_audio_buffer_overlap_len(const struct nobs_audio_buffer* b);
float nobs_audio_buffer_noise_floor(const struct nobs_audio_buffer* b);

/* CUDA events on the library's own stream (slots 0..7): device-side timing of whole calls */
int whisper_b200_event_record(struct whisper_context* ctx, int slot);
double whisper_b200_event_elapsed_ms(struct whisper_context* ctx, int slot_a, int slot_b);

int whisper_b200_device_count(void);
const char* whisper_b200_last_error(void);

/* ---------------------------------------------------------------- host-side engine mirror
 * C++ class nobs::WhisperEngine (csrc/host/whisper_engine.h) restates the reference's
 * WhisperEngine (whisper.rs:16-197) above this ABI; these shims expose it to C / ctypes. */
struct nobs_engine;
struct nobs_engine* nobs_engine_new(void);                                   /* whisper.rs:22-27 */
void nobs_engine_free(struct nobs_engine* e);
int nobs_engine_load_model(struct nobs_engine* e, const char* path);         /* whisper.rs:36-52; 0 ok, 1 LoadError */
void nobs_engine_unload_model(struct nobs_engine* e);                        /* whisper.rs:55-59 */
int nobs_engine_is_loaded(struct nobs_engine* e);                            /* whisper.rs:62-64 */
/* whisper.rs:66-148.  language/vocabulary/context may be NULL (= None).  Returns 0 ok,
 * 2 TranscriptionError, 3 NoModel; *out is a NUL-terminated UTF-8 string owned by the engine
 * (valid until the next call on this engine from the same thread). */
int nobs_engine_transcribe(struct nobs_engine* e, const float* audio, int n, const char* language, const char* vocabulary,
                           const char* context, const char** out);
/* whisper.rs:152-197 */
int nobs_engine_transcribe_chunked(struct nobs_engine* e, const float* const* chunks, const int* n, int n_chunks, const char* language,
                                   const char* vocabulary, const char** out);
/* state.rs:757-792: the audio left when a recording stops (16 kHz): longer than 30 s -> cut at silences
 * (audio.rs find_silence_boundaries / split_at_silences), pieces transcribed in order with the previous text as
 * context (parallel != 0: decoded together, data-parallel, no chaining); results joined with " " and trimmed. */
int nobs_engine_transcribe_recording(struct nobs_engine* e, const float* audio, size_t n, const char* language, const char* vocabulary, int parallel,
                                     const char** out);
/* Data-parallel variant of transcribe for independent windows: no context
 * chaining; texts[i] receives pointers owned by the engine. */
/* whisper.rs:152-197 with the chunks decoded data-parallel: same result as nobs_engine_transcribe_chunked (chunk k's prompt carries
 * the text of the last non-empty chunk before it), obtained by speculative batches — decode a window of chunks together with the
 * context known so far, accept the longest prefix whose context turned out right, re-decode the rest.  abort_on_error = 1: stop at
 * the first failing chunk (whisper.rs:182-185); 0: skip it (state.rs:773-775).  n_decodes / n_rounds (optional): chunk decodes and
 * batched rounds spent (n_chunks decodes = no speculation wasted). */
int nobs_engine_transcribe_chunked_parallel(struct nobs_engine* e, const float* const* chunks, const int* n, int n_chunks, const char* language,
                                            const char* vocabulary, int abort_on_error, const char** out, int* n_decodes, int* n_rounds);
int nobs_engine_transcribe_batch(struct nobs_engine* e, const float* const* audios, const int* n, int n_audios, const char* language,
                                 const char* vocabulary, int beam_size, const char** texts);
const char* nobs_engine_last_error(struct nobs_engine* e);
/* counters of the engine's last transcribe / transcribe_batch call, summed over its audios */
int nobs_engine_last_stats(struct nobs_engine* e, whisper_b200_stats* out);
/* the loaded context (NULL if none), e.g. for whisper_b200_set_profiling */
struct whisper_context* nobs_engine_context(struct nobs_engine* e);
/* whisper.rs:233-260; returns a pointer to a thread-local buffer */
const char* nobs_filter_hallucinations(const char* text);

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H */
