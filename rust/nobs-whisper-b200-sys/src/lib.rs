//! Raw bindings to `include/whisper_b200.h`.  Struct layouts mirror the C header field for field
//! (`#[repr(C)]`); `whisper_full_params` is 296 bytes and `whisper_context_params` 48 bytes on LP64,
//! pinned by `tests/test_host_logic.py::test_full_params_defaults_and_layout`.
#![allow(non_camel_case_types)]

use libc::{c_char, c_float, c_int, c_void, size_t};

#[repr(C)]
pub struct whisper_context {
    _private: [u8; 0],
}
#[repr(C)]
pub struct whisper_state {
    _private: [u8; 0],
}
pub type whisper_token = i32;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct whisper_ahead {
    pub n_text_layer: c_int,
    pub n_head: c_int,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct whisper_aheads {
    pub n_heads: size_t,
    pub heads: *const whisper_ahead,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct whisper_context_params {
    pub use_gpu: bool,
    pub flash_attn: bool,
    pub gpu_device: c_int,
    pub dtw_token_timestamps: bool,
    pub dtw_aheads_preset: c_int,
    pub dtw_n_top: c_int,
    pub dtw_aheads: whisper_aheads,
    pub dtw_mem_size: size_t,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct whisper_token_data {
    pub id: whisper_token,
    pub tid: whisper_token,
    pub p: c_float,
    pub plog: c_float,
    pub pt: c_float,
    pub ptsum: c_float,
    pub t0: i64,
    pub t1: i64,
    pub t_dtw: i64,
    pub vlen: c_float,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct whisper_greedy_params {
    pub best_of: c_int,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct whisper_beam_search_params {
    pub beam_size: c_int,
    pub patience: c_float,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct whisper_vad_params {
    pub threshold: c_float,
    pub min_speech_duration_ms: c_int,
    pub min_silence_duration_ms: c_int,
    pub max_speech_duration_s: c_float,
    pub speech_pad_ms: c_int,
    pub samples_overlap: c_float,
}
pub const WHISPER_SAMPLING_GREEDY: c_int = 0;
pub const WHISPER_SAMPLING_BEAM_SEARCH: c_int = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct whisper_full_params {
    pub strategy: c_int,
    pub n_threads: c_int,
    pub n_max_text_ctx: c_int,
    pub offset_ms: c_int,
    pub duration_ms: c_int,
    pub translate: bool,
    pub no_context: bool,
    pub no_timestamps: bool,
    pub single_segment: bool,
    pub print_special: bool,
    pub print_progress: bool,
    pub print_realtime: bool,
    pub print_timestamps: bool,
    pub token_timestamps: bool,
    pub thold_pt: c_float,
    pub thold_ptsum: c_float,
    pub max_len: c_int,
    pub split_on_word: bool,
    pub max_tokens: c_int,
    pub debug_mode: bool,
    pub audio_ctx: c_int,
    pub tdrz_enable: bool,
    pub suppress_regex: *const c_char,
    pub initial_prompt: *const c_char,
    pub prompt_tokens: *const whisper_token,
    pub prompt_n_tokens: c_int,
    pub language: *const c_char,
    pub detect_language: bool,
    pub suppress_blank: bool,
    pub suppress_nst: bool,
    pub temperature: c_float,
    pub max_initial_ts: c_float,
    pub length_penalty: c_float,
    pub temperature_inc: c_float,
    pub entropy_thold: c_float,
    pub logprob_thold: c_float,
    pub no_speech_thold: c_float,
    pub greedy: whisper_greedy_params,
    pub beam_search: whisper_beam_search_params,
    pub new_segment_callback: *mut c_void,
    pub new_segment_callback_user_data: *mut c_void,
    pub progress_callback: *mut c_void,
    pub progress_callback_user_data: *mut c_void,
    pub encoder_begin_callback: *mut c_void,
    pub encoder_begin_callback_user_data: *mut c_void,
    pub abort_callback: *mut c_void,
    pub abort_callback_user_data: *mut c_void,
    pub logits_filter_callback: *mut c_void,
    pub logits_filter_callback_user_data: *mut c_void,
    pub grammar_rules: *const *const c_void,
    pub n_grammar_rules: size_t,
    pub i_start_rule: size_t,
    pub grammar_penalty: c_float,
    pub vad: bool,
    pub vad_model_path: *const c_char,
    pub vad_params: whisper_vad_params,
}

extern "C" {
    pub fn whisper_context_default_params() -> whisper_context_params;
    pub fn whisper_init_from_file_with_params_no_state(path_model: *const c_char, params: whisper_context_params) -> *mut whisper_context;
    pub fn whisper_free(ctx: *mut whisper_context);
    pub fn whisper_init_state(ctx: *mut whisper_context) -> *mut whisper_state;
    pub fn whisper_free_state(state: *mut whisper_state);
    pub fn whisper_full_default_params(strategy: c_int) -> whisper_full_params;
    pub fn whisper_full_with_state(ctx: *mut whisper_context, state: *mut whisper_state, params: whisper_full_params, samples: *const c_float,
                                   n_samples: c_int) -> c_int;
    pub fn whisper_full_n_segments_from_state(state: *mut whisper_state) -> c_int;
    pub fn whisper_full_get_segment_text_from_state(state: *mut whisper_state, i_segment: c_int) -> *const c_char;
    pub fn whisper_full_get_segment_t0_from_state(state: *mut whisper_state, i_segment: c_int) -> i64;
    pub fn whisper_full_get_segment_t1_from_state(state: *mut whisper_state, i_segment: c_int) -> i64;
    pub fn whisper_full_get_segment_no_speech_prob_from_state(state: *mut whisper_state, i_segment: c_int) -> c_float;
    pub fn whisper_full_n_tokens_from_state(state: *mut whisper_state, i_segment: c_int) -> c_int;
    pub fn whisper_full_get_token_id_from_state(state: *mut whisper_state, i_segment: c_int, i_token: c_int) -> whisper_token;
    pub fn whisper_full_get_token_data_from_state(state: *mut whisper_state, i_segment: c_int, i_token: c_int) -> whisper_token_data;
    pub fn whisper_full_lang_id_from_state(state: *mut whisper_state) -> c_int;
    pub fn whisper_tokenize(ctx: *mut whisper_context, text: *const c_char, tokens: *mut whisper_token, n_max_tokens: c_int) -> c_int;
    pub fn whisper_token_to_str(ctx: *mut whisper_context, token: whisper_token) -> *const c_char;
    pub fn whisper_lang_id(lang: *const c_char) -> c_int;
    // B200 extension: independent audios decoded in lock step on the context's GPU
    pub fn whisper_b200_full_batch(ctx: *mut whisper_context, states: *const *mut whisper_state, n: c_int, params: whisper_full_params,
                                   samples: *const *const c_float, n_samples: *const c_int, rc: *mut c_int) -> c_int;
    pub fn whisper_b200_last_error() -> *const c_char;

    // ---- audio.rs entry points (include/whisper_b200.h: silence chunker, resampler, capture buffer)
    pub fn nobs_find_silence_boundaries(audio: *const c_float, n_samples: size_t, sample_rate: u32, boundaries: *mut size_t, cap: size_t,
                                        n_found: *mut size_t) -> c_int;
    pub fn nobs_split_at_silences_with_overlap(n_samples: size_t, boundaries: *const size_t, n_boundaries: size_t, sample_rate: u32,
                                               ranges: *mut size_t, n_chunks: *mut size_t) -> c_int;
    pub fn nobs_split_at_silences(n_samples: size_t, boundaries: *const size_t, n_boundaries: size_t, ranges: *mut size_t, n_chunks: *mut size_t) -> c_int;
    pub fn nobs_resample_audio(audio: *const c_float, n: size_t, from_rate: u32, to_rate: u32, out: *mut c_float, cap: size_t, n_out: *mut size_t) -> c_int;
    pub fn nobs_resample_chunk(audio: *const c_float, n: size_t, input_sample_rate: u32, out: *mut c_float, cap: size_t, n_out: *mut size_t) -> c_int;
    pub fn nobs_mix_to_mono(interleaved: *const c_float, n_frames: size_t, channels: u32, out: *mut c_float) -> c_int;
    pub fn nobs_calculate_rms(samples: *const c_float, n: size_t) -> c_float;
    pub fn nobs_audio_buffer_new(sample_rate: u32) -> *mut nobs_audio_buffer;
    pub fn nobs_audio_buffer_free(b: *mut nobs_audio_buffer);
    pub fn nobs_audio_buffer_push_samples(b: *mut nobs_audio_buffer, samples: *const c_float, n: size_t);
    pub fn nobs_audio_buffer_has_silence_boundary(b: *const nobs_audio_buffer) -> c_int;
    pub fn nobs_audio_buffer_take_chunk_at_silence(b: *mut nobs_audio_buffer, n: *mut size_t) -> *const c_float;
    pub fn nobs_audio_buffer_take_forced_chunk(b: *mut nobs_audio_buffer, n: *mut size_t) -> *const c_float;
    pub fn nobs_audio_buffer_take(b: *mut nobs_audio_buffer, n: *mut size_t) -> *const c_float;
    pub fn nobs_audio_buffer_len(b: *const nobs_audio_buffer) -> size_t;
    pub fn nobs_audio_buffer_noise_floor(b: *const nobs_audio_buffer) -> c_float;

    // ---- host-side engine mirror (include/whisper_b200.h: the reference's WhisperEngine, whisper.rs:16-197, above the C ABI)
    pub fn nobs_engine_new() -> *mut nobs_engine;
    pub fn nobs_engine_free(e: *mut nobs_engine);
    pub fn nobs_engine_load_model(e: *mut nobs_engine, path: *const c_char) -> c_int;
    pub fn nobs_engine_unload_model(e: *mut nobs_engine);
    pub fn nobs_engine_is_loaded(e: *mut nobs_engine) -> c_int;
    pub fn nobs_engine_transcribe(e: *mut nobs_engine, audio: *const c_float, n: c_int, language: *const c_char, vocabulary: *const c_char,
                                  context: *const c_char, out: *mut *const c_char) -> c_int;
    pub fn nobs_engine_transcribe_chunked(e: *mut nobs_engine, chunks: *const *const c_float, n: *const c_int, n_chunks: c_int, language: *const c_char,
                                          vocabulary: *const c_char, out: *mut *const c_char) -> c_int;
    pub fn nobs_engine_transcribe_chunked_parallel(e: *mut nobs_engine, chunks: *const *const c_float, n: *const c_int, n_chunks: c_int,
                                                   language: *const c_char, vocabulary: *const c_char, abort_on_error: c_int, out: *mut *const c_char,
                                                   n_decodes: *mut c_int, n_rounds: *mut c_int) -> c_int;
    pub fn nobs_engine_transcribe_recording(e: *mut nobs_engine, audio: *const c_float, n: size_t, language: *const c_char, vocabulary: *const c_char,
                                            parallel: c_int, out: *mut *const c_char) -> c_int;
    pub fn nobs_engine_transcribe_batch(e: *mut nobs_engine, audios: *const *const c_float, n: *const c_int, n_audios: c_int, language: *const c_char,
                                        vocabulary: *const c_char, beam_size: c_int, texts: *mut *const c_char) -> c_int;
    pub fn nobs_engine_last_error(e: *mut nobs_engine) -> *const c_char;
    pub fn nobs_filter_hallucinations(text: *const c_char) -> *const c_char;
}

#[repr(C)]
pub struct nobs_engine {
    _private: [u8; 0],
}

#[repr(C)]
pub struct nobs_audio_buffer {
    _private: [u8; 0],
}
