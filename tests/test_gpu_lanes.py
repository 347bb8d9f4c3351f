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
