"""Fused decoder projection chains (csrc/decode_chain_sm100.cu).

The chain kernel keeps the tiling, split order and epilogue arithmetic of the multi-launch path
(gemm_skinny_sm100_kernel + skinny_reduce_kernel), so switching it on must not change a single bit:
logits of prefill / step decodes and whole transcripts are compared for equality between
NOBS_WHISPER_CHAIN=0 and =1, over the three row-tile widths (<= 32, <= 64, <= 128 rows), 1-3 decode
lanes, greedy with fallback and beam search.  The chain is an OPTION (NOBS_WHISPER_CHAIN=1): measured on B200 it is
not faster than the multi-launch path (a device-wide barrier costs what a PDL launch boundary costs,
profiles/README.md), so it is off by default; these tests keep it correct."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_ctx(nw, path, chain, lanes=None):
    os.environ["NOBS_WHISPER_CHAIN"] = "1" if chain else "0"
    os.environ["NOBS_WHISPER_FC1_FUSED"] = "0"      # the reference side of this comparison is the split-K GEMM + epilogue pair for every projection
    if lanes is not None:
        os.environ["NOBS_WHISPER_LANES"] = str(lanes)
    try:
        return nw.WhisperContext.new_with_params(path, nw.WhisperContextParameters.default(), precision="bf16")
    finally:
        del os.environ["NOBS_WHISPER_CHAIN"]
        del os.environ["NOBS_WHISPER_FC1_FUSED"]
        os.environ.pop("NOBS_WHISPER_LANES", None)


def params(nw, beam=0, prompt=None):
    p = nw.FullParams.new(nw.SamplingStrategy.BeamSearch(beam_size=beam) if beam else nw.SamplingStrategy.Greedy(best_of=1))
    p.set_language("en")
    if prompt:
        p.set_initial_prompt(prompt)
    p.set_no_context(False); p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(-1.0)
    return p


@pytest.mark.parametrize("arch,ftype", [("tiny", 1), ("base", 1), ("large-v3-turbo", 1)])
def test_stage_logits_are_bit_identical(model_dir, arch, ftype):
    """whisper_decode_with_state: a 6-token prefill (rows share a KV slot: the chain reduces QKV itself), a 40-token
    prefill and single-token steps (the self-attention kernel finishes the chain's QKV partial sums)."""
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    path = ggml_synth.ensure_model(model_dir, arch, ftype=ftype, init="fanin" if arch == "base" else "survey")
    pcm = synth_audio.synth_clip(3, 30.0)
    outs = []
    for chain in (0, 1):
        ctx = make_ctx(nw, path, chain)
        st = ctx.create_state()
        st.pcm_to_mel(pcm)
        st.encode(0)
        prompt = [ctx.token_sot(), ctx.token_lang(0), ctx.token_transcribe(), 11, 22, 33]
        got = [np.array(st.decode(prompt, 0))]
        for i, t in enumerate([44, 55, 66, 77]):
            got.append(np.array(st.decode([t], len(prompt) + i)))
        got.append(np.array(st.decode(list(range(100, 140)), len(prompt) + 4)))
        got.append(np.array(st.decode([88], len(prompt) + 44)))
        outs.append(got)
        st.close()
        ctx.close()
    for a, b in zip(*outs):
        assert np.isfinite(a).all()
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n_clips,lanes,kw", [(9, 1, {}), (40, 2, {}), (100, 1, {}), (26, 3, {}), (7, 2, dict(beam=3, prompt="Claude Code, Anthropic"))])
def test_transcripts_are_bit_identical(model_dir, n_clips, lanes, kw):
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth, synth_audio
    path = ggml_synth.ensure_model(model_dir, "tiny", ftype=1)
    clips = [synth_audio.synth_clip(300 + i, 30.0 if i % 5 else 11.0) for i in range(n_clips)]
    outs = []
    for chain in (0, 1):
        ctx = make_ctx(nw, path, chain, lanes)
        states = [ctx.create_state() for _ in clips]
        assert nw.full_batch(ctx, states, params(nw, **kw), clips) == [0] * n_clips
        outs.append([st.segments() for st in states])
        launches = states[0].stats().n_kernel_launches
        for st in states:
            st.close()
        ctx.close()
        outs.append(launches)
    assert outs[0] == outs[2]
    assert sum(len(s) for s in outs[0]) > 0
    assert outs[3] < outs[1]   # the fused path really ran: fewer launches for the same work
