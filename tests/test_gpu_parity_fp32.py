"""GPU parity tests, fp32 mode: the CUDA path (through the C ABI) against the CPU oracle on
identical ggml weights and synthetic 16 kHz audio.  Tolerances from BASELINE.json north_star:
mel 1e-4 relative, encoder output / logits 1e-4 relative in fp32 mode, token-exact greedy."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

VOCAB_PROMPT = ("Claude Code, Anthropic, Supabase, Vercel, shadcn, tRPC, Drizzle, Zod, pnpm, Bun, Deno, Turso, Neon, "
                "PlanetScale, Turborepo, Tauri, SvelteKit, Nuxt, Astro, Vite")


@pytest.fixture(scope="module")
def env(model_dir):
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    from oracle import oracle

    class Env:
        pass

    e = Env()
    e.nw, e.oracle, e.synth = nw, oracle, synth_audio
    e.ctxs, e.oracles = {}, {}

    def get(arch, init="survey"):
        key = (arch, init)
        if key not in e.ctxs:
            path = ggml_synth.ensure_model(model_dir, arch, init=init)
            e.ctxs[key] = nw.WhisperContext.new_with_params(path, nw.WhisperContextParameters.default().use_gpu(True), precision="fp32")
            e.oracles[key] = oracle.Oracle(path)
        return e.ctxs[key], e.oracles[key]

    e.get = get
    yield e
    for c in e.ctxs.values():
        c.close()


def ref_params(nw, language="en", prompt=None, beam=0, temperature_inc=None):
    """The exact parameter block of the reference (src-tauri/src/whisper.rs:88-124)."""
    p = nw.FullParams.new(nw.SamplingStrategy.BeamSearch(beam_size=beam) if beam else nw.SamplingStrategy.Greedy(best_of=1))
    p.set_language(language)
    if prompt:
        p.set_initial_prompt(prompt)
    p.set_print_special(False); p.set_print_progress(False); p.set_print_realtime(False); p.set_print_timestamps(False)
    p.set_translate(False); p.set_no_context(False); p.set_single_segment(False)
    p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(-1.0)
    if temperature_inc is not None:
        p.set_temperature_inc(temperature_inc)
    return p


def _compare_segments(got, want):
    assert len(got) == len(want), (got, want)
    for g, w in zip(got, want):
        assert g["tokens"] == w["tokens"]      # token-exact
        assert g["text"] == w["text"]
        assert (g["t0"], g["t1"]) == (w["t0"], w["t1"])


@pytest.mark.parametrize("seconds", [30.0, 5.0, 0.37, 47.3])
def test_mel_parity(env, seconds):
    ctx, orc = env.get("micro")
    pcm = env.synth.synth_clip(3, seconds)
    st = ctx.create_state()
    st.pcm_to_mel(pcm)
    got = st.get_mel()
    want, n_len_org = orc.mel(pcm)
    assert got.shape == want.shape
    assert st.n_len() == n_len_org
    assert rel_err(got, want) < 1e-4  # north star: mel within 1e-4 relative
    st.close()


def test_v3_layout_128_mel_bins(env):
    """large-v3 family specifics on a small model: 128-bin filterbank, 51866-token vocabulary
    (special tokens shifted by one)."""
    ctx, orc = env.get("micro128", "fanin")
    assert ctx.n_mels() == 128 and ctx.n_vocab() == 51866 and ctx.token_beg() == 50365 == orc.token_beg
    pcm = env.synth.synth_clip(31, 14.0)
    st = ctx.create_state()
    st.pcm_to_mel(pcm)
    want, _ = orc.mel(pcm)
    assert rel_err(st.get_mel(), want) < 1e-4
    st.encode(0)
    assert rel_err(st.encoder_output(), orc.encode(0)) < 1e-4
    st.full(ref_params(env.nw), pcm)
    _compare_segments(st.segments(), orc.full(env.oracle.reference_params("en"), pcm))
    st.close()


def test_mel_edge_signals(env):
    ctx, orc = env.get("micro")
    for pcm in (np.zeros(16000 * 3, np.float32), np.ones(16000 * 2, np.float32), -np.ones(7777, np.float32)):
        st = ctx.create_state()
        st.pcm_to_mel(pcm)
        want, _ = orc.mel(pcm)
        assert rel_err(st.get_mel(), want) < 1e-4
        st.close()


@pytest.mark.parametrize("arch,init", [("micro", "survey"), ("micro", "fanin"), ("tiny", "survey")])
def test_encoder_and_logits_parity(env, arch, init):
    ctx, orc = env.get(arch, init)
    pcm = env.synth.synth_clip(0, 30.0)
    st = ctx.create_state()
    st.pcm_to_mel(pcm)
    st.encode(0)
    orc.mel(pcm)
    want_enc = orc.encode(0)
    got_enc = st.encoder_output()
    assert rel_err(got_enc, want_enc) < 1e-4  # north star: encoder output 1e-4 relative in fp32 mode
    for layer in (0, ctx.n_text_layer() - 1):
        gk, gv = st.cross_kv(layer)
        wk, wv = orc.cross_kv(layer)
        assert rel_err(gk, wk) < 1e-4 and rel_err(gv, wv) < 1e-4
    prompt = [ctx.token_sot(), ctx.token_lang(0), ctx.token_transcribe(), 11, 22, 33]
    got = st.decode(prompt, 0)
    want = orc.decode(prompt, 0, 0)
    assert rel_err(got, want) < 1e-4  # north star: logits 1e-4 relative in fp32 mode
    # incremental step on top of the cached prompt
    got2 = st.decode([44], len(prompt))
    want2 = orc.decode([44], len(prompt), 0)
    assert rel_err(got2, want2) < 1e-4
    assert int(np.argmax(got2)) == int(np.argmax(want2))
    st.close()


def test_encoder_second_window_offset(env):
    ctx, orc = env.get("micro")
    pcm = env.synth.synth_clip(5, 50.0)
    st = ctx.create_state()
    st.pcm_to_mel(pcm)
    orc.mel(pcm)
    for off in (1234, 3000):
        st.encode(off)
        assert rel_err(st.encoder_output(), orc.encode(off)) < 1e-4
    st.close()


@pytest.mark.parametrize("temperature,mode", [(0.0, 0), (0.4, 1), (1.0, 1), (0.0, 2)])
def test_logit_filter_and_sampling_parity(env, temperature, mode):
    """K6 against the oracle's process_logits on the same logits, for the histories that
    exercise every rule: initial step, after text, after one / two timestamps, has_ts."""
    ctx, orc = env.get("micro", "fanin")
    nw = env.nw
    rng = np.random.default_rng(7)
    beg, eot = ctx.token_beg(), ctx.token_eot()
    p = ref_params(nw)
    op = env.oracle.reference_params("en")
    us = env.oracle.canonical_stream(8)
    cases = [([], False, 3000), ([100], False, 3000), ([beg + 10], True, 20), ([beg + 10, beg + 10], True, 20),
             ([beg + 5, 200, beg + 40], True, 80), ([300, 400], False, 3000)]
    for ci, (hist, has_ts, seek_delta) in enumerate(cases):
        logits = (rng.standard_normal(ctx.n_vocab()) * (0.3 if ci % 2 else 2.0)).astype(np.float32)
        want_lp, want_pr = orc.process_logits(op, logits, hist, has_ts, seek_delta, temperature)
        lp, pr, res = ctx.process_logits(p, logits, hist, has_ts, seek_delta, temperature, mode=mode, u=float(us[ci]), k=5)
        fin = np.isfinite(want_lp)
        assert np.array_equal(np.isfinite(lp), fin)  # identical suppression mask
        assert np.abs(lp[fin] - want_lp[fin]).max() < 1e-4
        assert rel_err(pr, want_pr) < 1e-4  # the oracle's sequential fp32 logsumexp vs the GPU's tree reduction
        if mode == 0:
            assert res.id == int(np.argmax(want_pr))
        elif mode == 1:
            cp = np.cumsum(want_pr.astype(np.float64) / want_pr.astype(np.float64).sum())
            assert res.id == int(np.searchsorted(cp, us[ci], side="left"))
        else:
            order = np.lexsort((np.arange(len(want_lp)), -want_lp))[:5]
            assert [res.topk_id[i] for i in range(res.n_topk)] == [int(i) for i in order if np.isfinite(want_lp[i])]
        raw = logits.astype(np.float64)
        nosp = np.exp(raw[ctx.token_nosp()] - raw.max()) / np.exp(raw - raw.max()).sum()
        assert abs(res.no_speech_prob - nosp) < 1e-6


@pytest.mark.parametrize("arch,init,clip,seconds", [
    ("micro", "survey", 0, 30.0), ("micro", "fanin", 1, 30.0), ("tiny", "survey", 0, 30.0),
    ("micro", "survey", 2, 5.0), ("micro", "survey", 4, 47.3), ("micro", "fanin", 6, 11.0)])
def test_full_greedy_token_exact(env, arch, init, clip, seconds):
    """BASELINE config 1 (and ragged lengths): the reference's own call sequence, greedy with the
    default temperature fallback, token-exact against the oracle on the fp32 path."""
    ctx, orc = env.get(arch, init)
    pcm = env.synth.synth_clip(clip, seconds)
    st = ctx.create_state()
    st.full(ref_params(env.nw), pcm)
    want = orc.full(env.oracle.reference_params("en"), pcm)
    _compare_segments(st.segments(), want)
    s = st.stats()
    assert s.n_windows >= 1 and s.n_kernel_launches > 0
    st.close()


def test_full_with_initial_prompt_and_no_fallback(env):
    ctx, orc = env.get("micro", "fanin")
    pcm = env.synth.synth_clip(8, 20.0)
    st = ctx.create_state()
    st.full(ref_params(env.nw, prompt=VOCAB_PROMPT, temperature_inc=0.0), pcm)
    want = orc.full(env.oracle.reference_params("en", initial_prompt=VOCAB_PROMPT, temperature_inc=0.0), pcm)
    _compare_segments(st.segments(), want)
    st.close()


def test_full_auto_language(env):
    ctx, orc = env.get("micro", "fanin")
    pcm = env.synth.synth_clip(9, 12.0)
    st = ctx.create_state()
    st.full(ref_params(env.nw, language=None), pcm)
    want = orc.full(env.oracle.reference_params(None), pcm)
    _compare_segments(st.segments(), want)
    assert st.full_lang_id() == orc.L.wo_lang_id(orc.h)
    st2 = ctx.create_state()
    st2.pcm_to_mel(pcm)
    lid, probs = st2.lang_auto_detect()
    olid, oprobs = orc.lang_detect()
    assert lid == olid and np.abs(probs - oprobs).max() < 1e-5
    st.close(); st2.close()


def test_full_beam_search_with_prompt(env):
    """BASELINE config 2 shape (beam_size 5 + initial_prompt), on the micro architecture."""
    ctx, orc = env.get("micro", "fanin")
    pcm = env.synth.synth_clip(10, 30.0)
    st = ctx.create_state()
    st.full(ref_params(env.nw, prompt=VOCAB_PROMPT, beam=5), pcm)
    want = orc.full(env.oracle.reference_params("en", initial_prompt=VOCAB_PROMPT, beam_size=5), pcm)
    _compare_segments(st.segments(), want)
    st.close()


def test_short_and_empty_inputs(env):
    ctx, orc = env.get("micro")
    nw = env.nw
    st = ctx.create_state()
    with pytest.raises(nw.WhisperError):
        st.full(ref_params(nw), np.zeros(0, np.float32))       # whisper-rs rejects an empty slice
    st.full(ref_params(nw), np.zeros(1200, np.float32))         # < 100 ms (delta_min): returns 0 segments
    assert st.full_n_segments() == 0
    assert orc.full(env.oracle.reference_params("en"), np.zeros(1200, np.float32)) == []
    # 0.1 s .. 1 s: the reference app sends every clip longer than 1600 samples (state.rs:749) and expects text back
    for seed, secs in ((31, 0.5), (32, 0.12), (33, 0.99)):
        short = env.synth.synth_clip(seed, secs)
        fresh = ctx.create_state()                  # the reference creates a state per call (whisper.rs:83-85): no carried prompt_past / RNG position
        fresh.full(ref_params(nw), short)
        want = orc.full(env.oracle.reference_params("en"), short)
        assert (len(want) > 0) == (secs > 0.2)      # 0.5 s and 0.99 s come back as one segment each; the 0.12-s clip decodes to nothing
        _compare_segments(fresh.segments(), want)
        fresh.close()
    pcm = np.zeros(16000 * 4, np.float32)                        # silence
    st.full(ref_params(nw), pcm)
    _compare_segments(st.segments(), orc.full(env.oracle.reference_params("en"), pcm))
    st.close()


def test_batch_equals_sequential(env):
    """Data-parallel batches give exactly what one `full` per audio gives."""
    ctx, orc = env.get("micro", "survey")
    nw = env.nw
    clips = [env.synth.synth_clip(20 + i, s) for i, s in enumerate([30.0, 30.0, 7.5, 30.0, 0.5, 18.0, 30.0])]
    states = [ctx.create_state() for _ in clips]
    rcs = nw.full_batch(ctx, states, ref_params(nw), clips)
    assert rcs == [0] * len(clips)
    for st, pcm in zip(states, clips):
        single = ctx.create_state()
        single.full(ref_params(nw), pcm)
        assert st.segments() == single.segments()
        _compare_segments(st.segments(), orc.full(env.oracle.reference_params("en"), pcm))
        single.close()
    for st in states:
        st.close()


def test_engine_wrapper_matches_reference_call_sequence(env, model_dir):
    """WhisperEngine mirror (whisper.rs:16-197): transcribe == trim + filter over the segments."""
    import os
    os.environ["NOBS_WHISPER_PRECISION"] = "fp32"
    nw = env.nw
    from nobs_whisper_b200 import ggml_synth
    path = ggml_synth.ensure_model(model_dir, "micro", init="fanin")
    ctx, orc = env.get("micro", "fanin")
    eng = nw.WhisperEngine.from_file(path)
    assert eng.is_loaded()
    pcm = env.synth.synth_clip(12, 9.0)
    text = eng.transcribe(pcm, language="en", vocabulary=VOCAB_PROMPT, context="previous words")
    want = orc.full(env.oracle.reference_params("en", initial_prompt=VOCAB_PROMPT + " previous words"), pcm)
    want_text = b"".join(s["text"] for s in want).decode("utf-8", errors="replace").strip()
    assert text == nw.filter_hallucinations(want_text)
    chunks = [env.synth.synth_clip(13, 4.0), env.synth.synth_clip(14, 6.0)]
    joined = eng.transcribe_chunked(chunks, language="en", vocabulary=None)
    t0 = eng.transcribe(chunks[0], "en", None, None)
    t1 = eng.transcribe(chunks[1], "en", None, t0 if t0 else None)
    assert joined == " ".join(t for t in (t0, t1) if t)
    batch = eng.transcribe_batch(chunks, language="en")
    assert batch == [eng.transcribe(c, "en", None, None) for c in chunks]
    eng.unload_model()
    assert not eng.is_loaded()
    with pytest.raises(nw.NoModel):
        eng.transcribe(pcm)
    eng.close()
