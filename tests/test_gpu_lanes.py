"""Decode lanes: partitioning a batch over independent streams must not change any result.
In the fp32 parity mode that is a bit-for-bit statement, checked against 1 lane, single calls and the oracle, for
greedy with fallback, beam search (KV copies per lane) and language auto-detect (per-lane logits)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def params(nw, beam=0, language="en", prompt=None):
    p = nw.FullParams.new(nw.SamplingStrategy.BeamSearch(beam_size=beam) if beam else nw.SamplingStrategy.Greedy(best_of=1))
    p.set_language(language)
    if prompt:
        p.set_initial_prompt(prompt)
    p.set_no_context(False); p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(-1.0)
    return p


def run_batch(nw, path, precision, lanes, clips, **kw):
    os.environ["NOBS_WHISPER_LANES"] = str(lanes)
    try:
        ctx = nw.WhisperContext.new_with_params(path, nw.WhisperContextParameters.default(), precision=precision)
    finally:
        del os.environ["NOBS_WHISPER_LANES"]
    assert ctx.decode_lanes() == lanes
    states = [ctx.create_state() for _ in clips]
    assert nw.full_batch(ctx, states, params(nw, **kw), clips) == [0] * len(clips)
    out = [(st.segments(), st.full_lang_id()) for st in states]
    for st in states:
        st.close()
    ctx.close()
    return out


@pytest.mark.parametrize("kw", [dict(), dict(beam=3, prompt="Claude Code, Anthropic"), dict(language=None)])
def test_fp32_results_do_not_depend_on_the_lane_count(model_dir, kw):
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    from oracle import oracle
    path = ggml_synth.ensure_model(model_dir, "micro", init="fanin")
    # 30-s windows, short clips, one under 1 s, one under 100 ms (no output) and a 75-s recording that needs three sequential windows
    # (its later windows are encoded while the other lanes keep decoding)
    clips = [synth_audio.synth_clip(70 + i, s) for i, s in enumerate([30.0, 12.0, 75.0, 0.4, 21.0, 30.0, 5.0, 0.06])]
    one = run_batch(nw, path, "fp32", 1, clips, **kw)
    for lanes in (2, 3):
        assert run_batch(nw, path, "fp32", lanes, clips, **kw) == one
    if not kw:   # and the oracle agrees token for token (greedy + fallback)
        orc = oracle.Oracle(path)
        for (segs, _), pcm in zip(one, clips):
            want = orc.full(oracle.reference_params("en"), pcm)
            assert [s["tokens"] for s in segs] == [s["tokens"] for s in want]
        orc.close()


def test_bf16_lanes_are_deterministic_and_well_formed(model_dir):
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    path = ggml_synth.ensure_model(model_dir, "tiny", ftype=1)
    clips = [synth_audio.synth_clip(90 + i, 30.0) for i in range(9)]
    a = run_batch(nw, path, "bf16", 2, clips, language=None)
    b = run_batch(nw, path, "bf16", 2, clips, language=None)
    assert a == b
    c = run_batch(nw, path, "bf16", 3, clips, beam=4)
    assert c == run_batch(nw, path, "bf16", 3, clips, beam=4)
    for segs, lang in a + c:
        assert 0 <= lang < 100
        for s in segs:
            assert s["t0"] <= s["t1"]


def test_sample_budget_cut_inside_a_shared_prompt_row(model_dir):
    """A lane round whose samples exceed the engine's per-chunk sample budget is cut into chunks; with best_of / beam_size > 1 the
    prompt's last row is sampled n times, and the cut must never separate such a row from the samples that refer to it
    (engine.cu decode_submit).  NOBS_WHISPER_DEC_SAMPLES=8 forces the cut into every round of a 6-audio batch with 5 decoders each;
    in the fp32 parity mode the result must be bit-identical to the unconstrained engine."""
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    path = ggml_synth.ensure_model(model_dir, "micro", init="fanin")
    clips = [synth_audio.synth_clip(300 + i, s) for i, s in enumerate([30.0, 9.0, 30.0, 17.0, 4.0, 30.0])]

    def run(budget, **kw):
        if budget:
            os.environ["NOBS_WHISPER_DEC_SAMPLES"] = str(budget)
        try:
            return run_batch(nw, path, "fp32", 1, clips, **kw)
        finally:
            os.environ.pop("NOBS_WHISPER_DEC_SAMPLES", None)

    for kw in (dict(beam=5), dict(beam=5, prompt="Claude Code, Anthropic")):
        assert run(8, **kw) == run(0, **kw)

    # greedy with best_of 5 at a sampling temperature: five decoders draw their first token from the one prompt row
    def run_best_of(budget):
        if budget:
            os.environ["NOBS_WHISPER_DEC_SAMPLES"] = str(budget)
        try:
            ctx = nw.WhisperContext.new_with_params(path, nw.WhisperContextParameters.default(), precision="fp32")
        finally:
            os.environ.pop("NOBS_WHISPER_DEC_SAMPLES", None)
        p = nw.FullParams.new(nw.SamplingStrategy.Greedy(best_of=5))
        p.set_language("en")
        p.set_temperature(0.4)
        p.set_no_context(False); p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(-1.0)
        states = [ctx.create_state() for _ in clips]
        assert nw.full_batch(ctx, states, p, clips) == [0] * len(clips)
        out = [st.segments() for st in states]
        for st in states:
            st.close()
        ctx.close()
        return out

    assert run_best_of(8) == run_best_of(0)
