"""Loader of libnobswhisper_b200.so and ctypes declarations of include/whisper_b200.h.

The library is the product; there is no Python or CPU fallback.  If the shared object is
missing it is built in-tree with nvcc (csrc/Makefile); if that fails the import raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnobswhisper_b200.so")
CSRC = os.path.join(_HERE, "csrc")

WHISPER_MAX_DECODERS = 8


def source_hash() -> str:
    """sha256 over every file the shared library is compiled from (sources, headers, Makefile)."""
    import hashlib
    h = hashlib.sha256()
    files = []
    for root, _, names in os.walk(CSRC):
        if os.path.basename(root) == "build" or os.sep + "build" + os.sep in root + os.sep:
            continue
        files += [os.path.join(root, n) for n in names if n.endswith((".cu", ".cuh", ".cpp", ".h")) or n == "Makefile"]
    files.append(os.path.join(os.path.dirname(_HERE), "include", "whisper_b200.h"))
    for f in sorted(files):
        h.update(os.path.relpath(f, _HERE).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    """Compile every CUDA / C++ source for sm_100a into the in-tree shared library (csrc/Makefile:
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo).  force: clean rebuild.  A record of what was
    done (mode, seconds, source hash) is left in csrc/build/build_record.json."""
    import json
    import time
    t0 = time.time()
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    before = os.path.getmtime(LIB_PATH) if os.path.exists(LIB_PATH) else None
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=subprocess.DEVNULL)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("build did not produce " + LIB_PATH)
    compiled = before is None or os.path.getmtime(LIB_PATH) != before
    rec = {"build_mode": "clean" if force else "incremental", "build_exercised": bool(compiled), "seconds": round(time.time() - t0, 1),
           "source_sha256": source_hash(), "library": os.path.relpath(LIB_PATH, os.path.dirname(_HERE)), "library_bytes": os.path.getsize(LIB_PATH),
           "flags": "-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"}
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    with open(os.path.join(CSRC, "build", "build_record.json"), "w") as f:
        json.dump(rec, f, indent=1)
    return LIB_PATH


class WhisperAhead(C.Structure):
    _fields_ = [("n_text_layer", C.c_int), ("n_head", C.c_int)]


class WhisperAheads(C.Structure):
    _fields_ = [("n_heads", C.c_size_t), ("heads", C.POINTER(WhisperAhead))]


class WhisperContextParams(C.Structure):
    _fields_ = [
        ("use_gpu", C.c_bool), ("flash_attn", C.c_bool), ("gpu_device", C.c_int),
        ("dtw_token_timestamps", C.c_bool), ("dtw_aheads_preset", C.c_int), ("dtw_n_top", C.c_int),
        ("dtw_aheads", WhisperAheads), ("dtw_mem_size", C.c_size_t),
    ]


class WhisperTokenData(C.Structure):
    _fields_ = [
        ("id", C.c_int32), ("tid", C.c_int32), ("p", C.c_float), ("plog", C.c_float), ("pt", C.c_float),
        ("ptsum", C.c_float), ("t0", C.c_int64), ("t1", C.c_int64), ("t_dtw", C.c_int64), ("vlen", C.c_float),
    ]


class _Greedy(C.Structure):
    _fields_ = [("best_of", C.c_int)]


class _BeamSearch(C.Structure):
    _fields_ = [("beam_size", C.c_int), ("patience", C.c_float)]


class WhisperVadParams(C.Structure):
    _fields_ = [
        ("threshold", C.c_float), ("min_speech_duration_ms", C.c_int), ("min_silence_duration_ms", C.c_int),
        ("max_speech_duration_s", C.c_float), ("speech_pad_ms", C.c_int), ("samples_overlap", C.c_float),
    ]


class WhisperFullParams(C.Structure):
    _fields_ = [
        ("strategy", C.c_int), ("n_threads", C.c_int), ("n_max_text_ctx", C.c_int), ("offset_ms", C.c_int),
        ("duration_ms", C.c_int),
        ("translate", C.c_bool), ("no_context", C.c_bool), ("no_timestamps", C.c_bool), ("single_segment", C.c_bool),
        ("print_special", C.c_bool), ("print_progress", C.c_bool), ("print_realtime", C.c_bool),
        ("print_timestamps", C.c_bool),
        ("token_timestamps", C.c_bool), ("thold_pt", C.c_float), ("thold_ptsum", C.c_float), ("max_len", C.c_int),
        ("split_on_word", C.c_bool), ("max_tokens", C.c_int),
        ("debug_mode", C.c_bool), ("audio_ctx", C.c_int), ("tdrz_enable", C.c_bool),
        ("suppress_regex", C.c_char_p), ("initial_prompt", C.c_char_p), ("prompt_tokens", C.POINTER(C.c_int32)),
        ("prompt_n_tokens", C.c_int),
        ("language", C.c_char_p), ("detect_language", C.c_bool),
        ("suppress_blank", C.c_bool), ("suppress_nst", C.c_bool),
        ("temperature", C.c_float), ("max_initial_ts", C.c_float), ("length_penalty", C.c_float),
        ("temperature_inc", C.c_float), ("entropy_thold", C.c_float), ("logprob_thold", C.c_float),
        ("no_speech_thold", C.c_float),
        ("greedy", _Greedy), ("beam_search", _BeamSearch),
        ("new_segment_callback", C.c_void_p), ("new_segment_callback_user_data", C.c_void_p),
        ("progress_callback", C.c_void_p), ("progress_callback_user_data", C.c_void_p),
        ("encoder_begin_callback", C.c_void_p), ("encoder_begin_callback_user_data", C.c_void_p),
        ("abort_callback", C.c_void_p), ("abort_callback_user_data", C.c_void_p),
        ("logits_filter_callback", C.c_void_p), ("logits_filter_callback_user_data", C.c_void_p),
        ("grammar_rules", C.c_void_p), ("n_grammar_rules", C.c_size_t), ("i_start_rule", C.c_size_t),
        ("grammar_penalty", C.c_float),
        ("vad", C.c_bool), ("vad_model_path", C.c_char_p), ("vad_params", WhisperVadParams),
    ]


class B200SampleResult(C.Structure):
    _fields_ = [
        ("id", C.c_int32), ("tid", C.c_int32), ("p", C.c_float), ("plog", C.c_float), ("pt", C.c_float),
        ("ptsum", C.c_float), ("no_speech_prob", C.c_float), ("n_topk", C.c_int),
        ("topk_id", C.c_int32 * WHISPER_MAX_DECODERS), ("topk_plog", C.c_float * WHISPER_MAX_DECODERS),
        ("topk_p", C.c_float * WHISPER_MAX_DECODERS),
    ]


class B200Stats(C.Structure):
    _fields_ = [
        ("n_windows", C.c_int64), ("n_decode_rounds", C.c_int64), ("n_decode_rows", C.c_int64),
        ("n_sample_rows", C.c_int64), ("n_fallbacks", C.c_int64), ("n_kernel_launches", C.c_int64),
        ("gpu_ms_mel", C.c_double), ("gpu_ms_encode", C.c_double), ("gpu_ms_decode", C.c_double),
        ("gpu_ms_enc_gemm", C.c_double), ("gpu_ms_enc_attn", C.c_double), ("n_enc_gemm", C.c_int64), ("n_enc_attn", C.c_int64),
        ("gpu_ms_dec_cross", C.c_double), ("n_dec_cross", C.c_int64), ("dec_cross_bytes", C.c_double),
    ]


vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
# int hook(user, seek, i_temp, step, decoder, n_prompt, n_vocab, logits*): whisper_b200_logits_hook
LOGITS_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float))

# name -> (restype, argtypes); every symbol include/whisper_b200.h declares
SIGNATURES = {
    "whisper_context_default_params": (WhisperContextParams, []),
    "whisper_init_from_file_with_params_no_state": (vp, [C.c_char_p, WhisperContextParams]),
    "whisper_free": (None, [vp]),
    "whisper_init_state": (vp, [vp]),
    "whisper_free_state": (None, [vp]),
    "whisper_full_default_params": (WhisperFullParams, [C.c_int]),
    "whisper_full_with_state": (C.c_int, [vp, vp, WhisperFullParams, fp, C.c_int]),
    "whisper_full_n_segments_from_state": (C.c_int, [vp]),
    "whisper_full_get_segment_text_from_state": (vp, [vp, C.c_int]),
    "whisper_full_get_segment_t0_from_state": (C.c_int64, [vp, C.c_int]),
    "whisper_full_get_segment_t1_from_state": (C.c_int64, [vp, C.c_int]),
    "whisper_full_get_segment_speaker_turn_next_from_state": (C.c_bool, [vp, C.c_int]),
    "whisper_full_get_segment_no_speech_prob_from_state": (C.c_float, [vp, C.c_int]),
    "whisper_full_n_tokens_from_state": (C.c_int, [vp, C.c_int]),
    "whisper_full_get_token_text_from_state": (vp, [vp, vp, C.c_int, C.c_int]),
    "whisper_full_get_token_id_from_state": (C.c_int32, [vp, C.c_int, C.c_int]),
    "whisper_full_get_token_data_from_state": (WhisperTokenData, [vp, C.c_int, C.c_int]),
    "whisper_full_get_token_p_from_state": (C.c_float, [vp, C.c_int, C.c_int]),
    "whisper_full_lang_id_from_state": (C.c_int, [vp]),
    "whisper_n_vocab": (C.c_int, [vp]),
    "whisper_n_text_ctx": (C.c_int, [vp]),
    "whisper_n_audio_ctx": (C.c_int, [vp]),
    "whisper_is_multilingual": (C.c_int, [vp]),
    "whisper_model_n_vocab": (C.c_int, [vp]),
    "whisper_model_n_audio_ctx": (C.c_int, [vp]),
    "whisper_model_n_audio_state": (C.c_int, [vp]),
    "whisper_model_n_audio_head": (C.c_int, [vp]),
    "whisper_model_n_audio_layer": (C.c_int, [vp]),
    "whisper_model_n_text_ctx": (C.c_int, [vp]),
    "whisper_model_n_text_state": (C.c_int, [vp]),
    "whisper_model_n_text_head": (C.c_int, [vp]),
    "whisper_model_n_text_layer": (C.c_int, [vp]),
    "whisper_model_n_mels": (C.c_int, [vp]),
    "whisper_model_ftype": (C.c_int, [vp]),
    "whisper_model_type": (C.c_int, [vp]),
    "whisper_token_to_str": (vp, [vp, C.c_int32]),
    "whisper_token_eot": (C.c_int32, [vp]),
    "whisper_token_sot": (C.c_int32, [vp]),
    "whisper_token_solm": (C.c_int32, [vp]),
    "whisper_token_prev": (C.c_int32, [vp]),
    "whisper_token_nosp": (C.c_int32, [vp]),
    "whisper_token_not": (C.c_int32, [vp]),
    "whisper_token_beg": (C.c_int32, [vp]),
    "whisper_token_lang": (C.c_int32, [vp, C.c_int]),
    "whisper_token_translate": (C.c_int32, [vp]),
    "whisper_token_transcribe": (C.c_int32, [vp]),
    "whisper_tokenize": (C.c_int, [vp, C.c_char_p, C.POINTER(C.c_int32), C.c_int]),
    "whisper_token_count": (C.c_int, [vp, C.c_char_p]),
    "whisper_lang_max_id": (C.c_int, []),
    "whisper_lang_id": (C.c_int, [C.c_char_p]),
    "whisper_lang_str": (C.c_char_p, [C.c_int]),
    "whisper_lang_str_full": (C.c_char_p, [C.c_int]),
    "whisper_print_system_info": (C.c_char_p, []),
    "whisper_version": (C.c_char_p, []),
    "whisper_pcm_to_mel_with_state": (C.c_int, [vp, vp, fp, C.c_int, C.c_int]),
    "whisper_n_len_from_state": (C.c_int, [vp]),
    "whisper_encode_with_state": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "whisper_decode_with_state": (C.c_int, [vp, vp, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int]),
    "whisper_get_logits_from_state": (fp, [vp]),
    "whisper_lang_auto_detect_with_state": (C.c_int, [vp, vp, C.c_int, C.c_int, fp]),
    "whisper_b200_init_from_file": (vp, [C.c_char_p, WhisperContextParams, C.c_int]),
    "whisper_b200_precision": (C.c_int, [vp]),
    "whisper_b200_decode_lanes": (C.c_int, [vp]),
    "whisper_b200_init_host_only": (vp, [C.c_char_p]),
    "whisper_b200_set_logits_hook": (None, [vp, LOGITS_HOOK, vp]),
    "whisper_b200_decode_batch": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.POINTER(C.c_int32), ip, ip, C.c_int, fp]),
    "whisper_b200_full_batch_prompts": (C.c_int, [vp, C.POINTER(vp), C.c_int, WhisperFullParams, C.POINTER(C.c_char_p), C.POINTER(fp), ip, ip]),
    "whisper_b200_full_batch": (C.c_int, [vp, C.POINTER(vp), C.c_int, WhisperFullParams, C.POINTER(fp), ip, ip]),
    "whisper_b200_get_mel": (C.c_int, [vp, fp, C.c_size_t]),
    "whisper_b200_get_encoder_output": (C.c_int, [vp, vp, fp, C.c_size_t]),
    "whisper_b200_get_cross_kv": (C.c_int, [vp, vp, C.c_int, fp, fp, C.c_size_t]),
    "whisper_b200_process_logits": (C.c_int, [vp, WhisperFullParams, fp, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_float,
                                              C.c_int, C.c_double, C.c_int, fp, fp, C.POINTER(B200SampleResult)]),
    "whisper_b200_get_stats": (C.c_int, [vp, C.POINTER(B200Stats)]),
    "whisper_b200_debug_gemm_bf16": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_size_t, fp, fp, C.c_int, fp, C.c_int, C.c_int, C.c_int,
                                               C.c_int, fp]),
    "whisper_b200_set_profiling": (None, [vp, C.c_int]),
    "whisper_b200_debug_enc_attention": (C.c_int, [C.c_int, C.c_int, fp, fp, C.c_int]),
    "whisper_b200_debug_dec_cross_attention": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, C.c_int]),
    "whisper_b200_debug_cross_groups": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    "whisper_b200_debug_time_decode_kernels": (C.c_int, [C.c_int, C.c_int, C.c_int, fp]),
    "whisper_b200_debug_dec_proj": (C.c_int, [C.c_int, C.c_int, C.c_int, fp, fp, fp, C.c_int, fp, fp, fp, fp, fp, fp, C.c_int, fp]),
    "whisper_b200_debug_grid_sync": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, fp]),
    "whisper_b200_window_rms": (C.c_int, [fp, C.c_size_t, C.c_uint32, fp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "nobs_find_silence_boundaries": (C.c_int, [fp, C.c_size_t, C.c_uint32, C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(C.c_size_t)]),
    "nobs_split_at_silences_with_overlap": (C.c_int, [C.c_size_t, C.POINTER(C.c_size_t), C.c_size_t, C.c_uint32, C.POINTER(C.c_size_t),
                                                      C.POINTER(C.c_size_t)]),
    "nobs_split_at_silences": (C.c_int, [C.c_size_t, C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "nobs_resample_audio": (C.c_int, [fp, C.c_size_t, C.c_uint32, C.c_uint32, fp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "nobs_resample_chunk": (C.c_int, [fp, C.c_size_t, C.c_uint32, fp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "nobs_mix_to_mono": (C.c_int, [fp, C.c_size_t, C.c_uint32, fp]),
    "nobs_calculate_rms": (C.c_float, [fp, C.c_size_t]),
    "nobs_audio_buffer_new": (vp, [C.c_uint32]),
    "nobs_audio_buffer_free": (None, [vp]),
    "nobs_audio_buffer_push_samples": (None, [vp, fp, C.c_size_t]),
    "nobs_audio_buffer_has_silence_boundary": (C.c_int, [vp]),
    "nobs_audio_buffer_take_chunk_at_silence": (fp, [vp, C.POINTER(C.c_size_t)]),
    "nobs_audio_buffer_take_forced_chunk": (fp, [vp, C.POINTER(C.c_size_t)]),
    "nobs_audio_buffer_take": (fp, [vp, C.POINTER(C.c_size_t)]),
    "nobs_audio_buffer_len": (C.c_size_t, [vp]),
    "nobs_audio_buffer_overlap_len": (C.c_size_t, [vp]),
    "nobs_audio_buffer_noise_floor": (C.c_float, [vp]),
    "whisper_b200_event_record": (C.c_int, [vp, C.c_int]),
    "whisper_b200_event_elapsed_ms": (C.c_double, [vp, C.c_int, C.c_int]),
    "whisper_b200_device_count": (C.c_int, []),
    "whisper_b200_last_error": (C.c_char_p, []),
    "nobs_engine_new": (vp, []),
    "nobs_engine_free": (None, [vp]),
    "nobs_engine_load_model": (C.c_int, [vp, C.c_char_p]),
    "nobs_engine_unload_model": (None, [vp]),
    "nobs_engine_is_loaded": (C.c_int, [vp]),
    "nobs_engine_transcribe": (C.c_int, [vp, fp, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p)]),
    "nobs_engine_transcribe_chunked": (C.c_int, [vp, C.POINTER(fp), ip, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p)]),
    "nobs_engine_transcribe_recording": (C.c_int, [vp, fp, C.c_size_t, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p)]),
    "nobs_engine_transcribe_chunked_parallel": (C.c_int, [vp, C.POINTER(fp), ip, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), ip, ip]),
    "nobs_engine_transcribe_batch": (C.c_int, [vp, C.POINTER(fp), ip, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p)]),
    "nobs_engine_last_error": (C.c_char_p, [vp]),
    "nobs_engine_last_stats": (C.c_int, [vp, C.POINTER(B200Stats)]),
    "nobs_engine_context": (vp, [vp]),
    "nobs_filter_hallucinations": (C.c_char_p, [C.c_char_p]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded product library.  Raises if it cannot be built or loaded — never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
