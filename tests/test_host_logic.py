"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, and the host-side
logic above the ABI (loader, vocabulary, tokenizer, parameter defaults, the WhisperEngine mirror's
error behaviour and hallucination filter) behaves like the reference — no compute call is made."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nw():
    import nobs_whisper_b200
    return nobs_whisper_b200


def test_library_exports_every_symbol_the_header_declares(nw):
    from nobs_whisper_b200 import _lib
    L = _lib.lib()  # resolves every entry of SIGNATURES or raises
    header = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    declared = set(re.findall(r"\b((?:whisper|nobs)_[a-z0-9_]+)\s*\(", header))
    declared -= {"whisper_new_segment_callback", "whisper_progress_callback", "whisper_encoder_begin_callback", "whisper_abort_callback",
                 "whisper_logits_filter_callback"}
    raw = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(raw, s)]
    assert not missing, missing
    assert declared <= set(_lib.SIGNATURES), sorted(declared - set(_lib.SIGNATURES))
    assert b"CPU_FALLBACK = 0" in L.whisper_print_system_info()


def test_full_params_defaults_and_layout(nw):
    from nobs_whisper_b200 import _lib
    L = _lib.lib()
    g = L.whisper_full_default_params(0)
    b = L.whisper_full_default_params(1)
    assert (g.strategy, g.greedy.best_of, g.beam_search.beam_size) == (0, 5, -1)
    assert (b.strategy, b.greedy.best_of, b.beam_search.beam_size) == (1, -1, 5)
    assert g.n_max_text_ctx == 16384 and g.no_context is True and g.suppress_blank is True and g.language == b"en"
    assert abs(g.temperature_inc - 0.2) < 1e-7 and abs(g.entropy_thold - 2.4) < 1e-6 and g.logprob_thold == -1.0 and abs(g.no_speech_thold - 0.6) < 1e-6
    assert g.max_initial_ts == 1.0 and g.length_penalty == -1.0 and g.n_threads == min(4, os.cpu_count())
    # the struct crosses the ABI by value: pin its size / key offsets (LP64)
    P = _lib.WhisperFullParams
    assert C.sizeof(P) == 296 and P.language.offset == 96 and P.greedy.offset == 136 and P.new_segment_callback.offset == 152 and P.vad_params.offset == 272
    cp = L.whisper_context_default_params()
    assert cp.use_gpu is True and cp.gpu_device == 0 and C.sizeof(_lib.WhisperContextParams) == 48 and C.sizeof(_lib.WhisperTokenData) == 56
    # the reference's parameter block (whisper.rs:88-124) through the whisper-rs mirror
    p = nw.FullParams.new(nw.SamplingStrategy.Greedy(best_of=1))
    p.set_language(None); p.set_initial_prompt("abc"); p.set_no_context(False)
    assert p._p.greedy.best_of == 1 and p._p.language is None and p._p.initial_prompt == b"abc" and p._p.no_context is False


def test_host_only_context_vocab_and_tokenizer(nw, model_dir):
    from nobs_whisper_b200 import ggml_synth
    from oracle import oracle
    path = ggml_synth.ensure_model(model_dir, "micro")
    ctx = nw.WhisperContext(path, host_only=True)
    o = oracle.Oracle(path)
    assert (ctx.n_vocab(), ctx.n_audio_ctx(), ctx.n_audio_state(), ctx.n_mels()) == (51865, 1500, 128, 80)
    assert (ctx.token_eot(), ctx.token_sot(), ctx.token_beg(), ctx.token_transcribe(), ctx.token_nosp()) == (50257, 50258, 50364, 50359, 50362)
    assert ctx.token_to_bytes(32) == b" " and ctx.token_to_bytes(50364) == b"[_BEG_]"
    vocab = ("Claude Code, Anthropic, Supabase, Vercel, shadcn, tRPC, Drizzle, Zod, pnpm, Bun, Deno, Turso, Neon, PlanetScale, Turborepo, "
             "Tauri, SvelteKit, Nuxt, Astro, Vite, Fly.io, Cloudflare Workers, v0")  # reference config.rs:41 (excerpt)
    cases = [vocab, "", " ", "it's  a\ttest\n\n 42nd!? ", "ünïcödé 한국어 テスト", "a" * 300, "'ll've 'd't", "  leading", "trailing   "]
    rnd = np.random.default_rng(0)
    alphabet = list("abcXYZ 019'\t\n,.!?-é한")
    cases += ["".join(rnd.choice(alphabet, size=int(rnd.integers(0, 60)))) for _ in range(300)]
    for text in cases:
        assert ctx.tokenize(text) == o.tokenize(text), text
    assert b"".join(ctx.token_to_bytes(t) for t in ctx.tokenize(vocab)) == vocab.encode()
    # no compute on a host-only handle, and no CPU fallback
    with pytest.raises(nw.WhisperError):
        ctx.create_state()
    L = __import__("nobs_whisper_b200._lib", fromlist=["lib"]).lib()
    assert L.whisper_lang_id(b"en") == 0 and L.whisper_lang_id(b"korean") == 5 and L.whisper_lang_id(b"xx") == -1
    assert L.whisper_lang_str(99) == b"yue" and L.whisper_lang_max_id() == 99
    ctx.close(); o.close()


def test_v3_vocabulary_layout(nw, model_dir):
    from nobs_whisper_b200 import ggml_synth
    ctx = nw.WhisperContext(ggml_synth.ensure_model(model_dir, "micro128"), host_only=True)
    assert ctx.n_vocab() == 51866 and ctx.n_mels() == 128
    assert (ctx.token_transcribe(), ctx.token_nosp(), ctx.token_beg()) == (50360, 50363, 50365)
    ctx.close()


def test_loader_rejects_bad_files(nw, model_dir, tmp_path):
    from nobs_whisper_b200 import ggml_synth
    good = open(ggml_synth.ensure_model(model_dir, "micro"), "rb").read()
    cases = {"empty.bin": b"", "magic.bin": b"\x00\x01\x02\x03" + good[4:2000], "truncated_header.bin": good[:30],
             "truncated_vocab.bin": good[:80000], "truncated_tensor.bin": good[: len(good) - 1000],
             "bad_dims.bin": good[:4] + struct.pack("<11i", 51865, 1500, 100, 3, 2, 448, 100, 3, 3, 80, 0) + good[48:]}
    for name, blob in cases.items():
        p = tmp_path / name
        p.write_bytes(blob)
        with pytest.raises(nw.WhisperError):
            nw.WhisperContext(str(p), host_only=True)
    with pytest.raises(nw.WhisperError):
        nw.WhisperContext(str(tmp_path / "does-not-exist.bin"), host_only=True)


def test_engine_wrapper_error_behaviour_without_a_model(nw, tmp_path):
    """Reference tests whisper.rs:272-283: new() is not loaded; transcribe without a model -> NoModel."""
    e = nw.WhisperEngine()
    assert not e.is_loaded()
    with pytest.raises(nw.NoModel):
        e.transcribe(np.zeros(16000, np.float32), None, None, None)
    with pytest.raises(nw.NoModel):
        e.transcribe_batch([np.zeros(16000, np.float32)])
    with pytest.raises(nw.LoadError) as ei:
        e.load_model(str(tmp_path / "missing.bin"))
    assert str(ei.value).startswith("Failed to load model: ")
    assert e.transcribe_chunked([]) == ""
    e.close()


def test_filter_hallucinations(nw):
    """Reference test whisper.rs:286-305, assertion for assertion, plus the documented edge cases."""
    f = nw.filter_hallucinations
    assert f("Thank you for watching!") == ""
    assert f("thanks for watching.") == ""
    assert f("Thank you for watching") == ""
    assert f("Subscribe to my channel") == ""
    assert f("you") == ""
    assert f("...") == ""
    assert f("시청해 주셔서 감사합니다") == ""
    assert f("Hello, this is a real sentence.") == "Hello, this is a real sentence."
    assert f("Thank you for watching the demo, now let me explain") == "Thank you for watching the demo, now let me explain"
    assert f("") == "" and f("   ") == "" and f("♪") == "" and f("…") == "" and f(" ♫♬ ") == "" and f("♫ ♬") == "♫ ♬"
    assert f("  padded text \n") == "padded text"
    assert f("YOU!!!") == "" and f("ご視聴ありがとうございました。") == "ご視聴ありがとうございました。" and f("ご視聴ありがとうございました") == ""
    assert f("谢谢观看…") == "" and f("you too") == "you too"


def test_synthetic_model_writer_roundtrip(model_dir):
    from nobs_whisper_b200 import ggml_synth
    p32 = ggml_synth.ensure_model(model_dir, "micro", seed=3, ftype=0)
    p16 = ggml_synth.ensure_model(model_dir, "micro", seed=3, ftype=1)
    assert os.path.getsize(p16) < os.path.getsize(p32)
    from oracle import oracle
    a, b = oracle.Oracle(p32), oracle.Oracle(p16)
    x = np.zeros(16000 * 2, np.float32)
    x[::50] = 0.2
    a.mel(x); b.mel(x)
    ea, eb = a.encode(0), b.encode(0)
    assert 0 < np.abs(ea - eb).max() < 5e-2 * np.abs(ea).max()   # f16 weight rounding only
    fb = ggml_synth.slaney_filterbank(80)
    assert fb.shape == (80, 201) and np.all(fb >= 0) and np.all(fb.sum(axis=1) > 0)
    assert len(set(ggml_synth.synthetic_vocab())) == 50257
    a.close(); b.close()


def test_cross_attention_row_groups(nw):
    """Host side of the decoder cross-attention's row groups (cross_attention_sm100.cu cross_attention_groups): consecutive rows of one
    audio slot are one work item of at most 4 rows; the kernel build (group width 1 / 2 / 4) follows the longest run."""
    import ctypes as C
    from nobs_whisper_b200 import _lib
    L = _lib.lib()

    def groups(slots):
        a = (C.c_int * len(slots))(*slots)
        g = (C.c_int * (len(slots) + 1))()
        width = L.whisper_b200_debug_cross_groups(a, len(slots), g)
        n = g[len(slots)]
        return width, [(g[i] & 0xFFFFFF, g[i] >> 24) for i in range(n)]

    # a plain step round: every row its own audio -> no groups needed
    assert groups([5, 6, 7, 9]) == (1, [(0, 1), (1, 1), (2, 1), (3, 1)])
    # a speculative round: (pass 0, shadow) pairs, one audio without a shadow
    assert groups([3, 3, 4, 4, 8, 9, 9]) == (2, [(0, 2), (2, 2), (4, 1), (5, 2)])
    # beam search: five beams per audio -> 4 + 1; the same slot coming back later is a new run
    assert groups([1] * 5 + [2] * 5 + [1]) == (4, [(0, 4), (4, 1), (5, 4), (9, 1), (10, 1)])
    # a three-row run needs the 4-wide build; a 9-row prompt is cut 4 + 4 + 1
    assert groups([7, 7, 7, 2]) == (4, [(0, 3), (3, 1)])
    assert groups([0] * 9) == (4, [(0, 4), (4, 4), (8, 1)])
    # every row is covered exactly once, in order
    rng = np.random.default_rng(0)
    for _ in range(50):
        slots = np.repeat(rng.integers(0, 6, 20), rng.integers(1, 7, 20)).tolist()
        width, gs = groups(slots)
        covered = [r for first, size in gs for r in range(first, first + size)]
        assert covered == list(range(len(slots)))
        assert all(1 <= size <= 4 and len(set(slots[first:first + size])) == 1 for first, size in gs)
        assert width == (1 if max(s for _, s in gs) == 1 else 2 if max(s for _, s in gs) == 2 else 4)
    assert L.whisper_b200_debug_cross_groups(None, 0, None) == -1


def test_rust_sys_crate_matches_the_header(nw):
    """rustc is not in this image, so the FFI stub a maintainer links (rust/nobs-whisper-b200-sys) is held to the C ABI textually:
    every function its extern block declares is exported by the library, declared in include/whisper_b200.h, and takes the same
    number of arguments there."""
    from nobs_whisper_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rust = open(os.path.join(root, "rust", "nobs-whisper-b200-sys", "src", "lib.rs")).read()
    header = open(os.path.join(root, "include", "whisper_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)

    def n_args(arglist):
        arglist = arglist.strip()
        if arglist in ("", "void"):
            return 0
        depth, n = 0, 1
        for ch in arglist:
            depth += ch in "(<["
            depth -= ch in ")>]"
            n += ch == "," and depth == 0
        return n

    c_decl = {m.group(1): n_args(m.group(2)) for m in re.finditer(r"\b((?:whisper|nobs)_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)}
    externs = "".join(re.findall(r'extern "C" \{(.*?)\n\}', rust, flags=re.S))
    rust_decl = {m.group(1): n_args(m.group(2)) for m in re.finditer(r"pub fn (\w+)\(([^;]*?)\)\s*(?:->[^;]*)?;", externs, flags=re.S)}
    assert len(rust_decl) >= 45
    raw = _lib.lib()
    for name, n in sorted(rust_decl.items()):
        assert hasattr(raw, name), name
        assert name in c_decl, name
        assert c_decl[name] == n, (name, c_decl[name], n)
    # the wrapper crate only calls what the sys crate declares
    wrapper = open(os.path.join(root, "rust", "nobs-whisper-b200", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::((?:whisper|nobs)_[a-z0-9_]+)\s*\(", wrapper))
    assert used and used <= set(rust_decl), sorted(used - set(rust_decl))
