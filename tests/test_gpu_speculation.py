"""Speculative first fallback (csrc/full.cpp, Shadow): temperature pass 1 is decoded in the same rounds as pass 0 from its own KV
slot.  Whether pass 0 is accepted (shadow dropped, generator restored) or rejected (shadow adopted, possibly already finished),
segments, tokens, timestamps and the fallback count must be exactly those of the sequential ladder (reference path:
whisper_full_with_state's temperature loop) — and the oracle agrees token for token."""
import os

import pytest

pytestmark = pytest.mark.gpu


def params(nw, thold):
    p = nw.FullParams.new(nw.SamplingStrategy.Greedy(best_of=1))
    p.set_language("en")
    p.set_no_context(False); p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(thold)
    return p


def run(nw, ctx, clips, thold, speculate):
    os.environ["NOBS_WHISPER_SPECULATE"] = "1" if speculate else "0"
    try:
        states = [ctx.create_state() for _ in clips]
        assert nw.full_batch(ctx, states, params(nw, thold), clips) == [0] * len(clips)
        out = [(st.segments(), int(st.stats().n_fallbacks), int(st.stats().n_windows)) for st in states]
        rows = sum(int(st.stats().n_decode_rows) for st in states)
        for st in states:
            st.close()
    finally:
        del os.environ["NOBS_WHISPER_SPECULATE"]
    return out, rows


@pytest.mark.parametrize("precision,arch", [("fp32", "micro"), ("bf16", "tiny")])
def test_speculation_changes_nothing(model_dir, precision, arch):
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    path = ggml_synth.ensure_model(model_dir, arch, init="fanin") if arch == "micro" else ggml_synth.ensure_model(model_dir, arch, ftype=1)
    ctx = nw.WhisperContext.new_with_params(path, nw.WhisperContextParameters.default(), precision=precision)
    # 30-s windows, short clips and a 75-s recording (three sequential windows: the generator state crosses windows)
    clips = [synth_audio.synth_clip(40 + i, s) for i, s in enumerate([30.0, 12.0, 75.0, 0.4, 21.0, 30.0, 5.0, 47.3])]
    mixed = False
    # thresholds from "pass 0 always accepted" to "always rejected"; around ln(1 / n_vocab) some windows pass and some do not
    for thold in (-1e9, -12.0, -11.0, -10.5, -10.0, -9.0, -1.0):
        seq, rows0 = run(nw, ctx, clips, thold, speculate=False)
        spec, rows1 = run(nw, ctx, clips, thold, speculate=True)
        assert spec == seq, thold
        fb = [f for _, f, _ in seq]
        wins = [w for _, _, w in seq]
        if any(f == 0 for f, w in zip(fb, wins) if w) and any(f > 0 for f in fb):
            mixed = True
        if thold == -1e9:
            assert rows1 > rows0                           # shadow rows were decoded next to accepted first passes and thrown away
        if thold == -1.0:
            assert all(f > 0 for f, w in zip(fb, wins) if w)
    assert mixed, "no threshold produced both accepted and rejected first passes"
    ctx.close()


def test_oracle_agrees_with_the_speculative_ladder(model_dir):
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    from oracle import oracle
    path = ggml_synth.ensure_model(model_dir, "micro", init="fanin")
    ctx = nw.WhisperContext.new_with_params(path, nw.WhisperContextParameters.default(), precision="fp32")
    clips = [synth_audio.synth_clip(60 + i, s) for i, s in enumerate([30.0, 47.3, 8.0])]
    got, _ = run(nw, ctx, clips, -1.0, speculate=True)
    orc = oracle.Oracle(path)
    for (segs, _, _), pcm in zip(got, clips):
        want = orc.full(oracle.reference_params("en"), pcm)
        assert [s["tokens"] for s in segs] == [s["tokens"] for s in want]
        assert [(s["t0"], s["t1"]) for s in segs] == [(s["t0"], s["t1"]) for s in want]
    orc.close()
    ctx.close()


def test_graph_replayed_rounds_change_nothing(model_dir):
    """Small step batches (single utterances: 1-8 rows) are captured once per shape as a CUDA graph and replayed
    (csrc/engine.cu decode_chunk).  Same kernels, same arguments: the bf16 result must be bit-identical to direct launches,
    for one clip at a time (the latency path) and for a small batch, greedy and beam."""
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    path = ggml_synth.ensure_model(model_dir, "tiny", ftype=1)
    clips = [synth_audio.synth_clip(20 + i, s) for i, s in enumerate([5.0, 30.0, 11.0])]

    def run(graph_rows):
        os.environ["NOBS_WHISPER_GRAPH_ROWS"] = str(graph_rows)
        try:
            ctx = nw.WhisperContext.new_with_params(path, nw.WhisperContextParameters.default(), precision="bf16")
        finally:
            del os.environ["NOBS_WHISPER_GRAPH_ROWS"]
        out = []
        for beam in (0, 3):
            p = nw.FullParams.new(nw.SamplingStrategy.BeamSearch(beam_size=beam) if beam else nw.SamplingStrategy.Greedy(best_of=1))
            p.set_language("en")
            p.set_no_context(False); p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(-1.0)
            for pcm in clips:                      # one utterance per call
                st = ctx.create_state()
                assert st.full(p, pcm) == 0
                out.append(st.segments())
                st.close()
            states = [ctx.create_state() for _ in clips]
            assert nw.full_batch(ctx, states, p, clips) == [0] * len(clips)
            out.append([st.segments() for st in states])
            for st in states:
                st.close()
        ctx.close()
        return out

    assert run(8) == run(0)
