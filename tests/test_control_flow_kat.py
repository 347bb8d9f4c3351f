"""Known-answer tests of `whisper_full`'s control flow with SCRIPTED logits.

The oracle and the product were written from the same recollection of whisper.cpp, so their agreeing with each
other proves little about the rules themselves.  Here the model's logits are replaced by a script (logits hook:
`wo_set_logits_hook` in the oracle, `whisper_b200_set_logits_hook` in the product — the product runs its real GPU
filter / log-softmax / sampling kernel on the injected logits) and the EXPECTED tokens, segment boundaries, seek
positions, fallback counts and prompt lengths are literals derived by hand from the rules as SURVEY.md states them:

  * timestamp pairing, max_initial_ts (1.0 s = beg + 50), suppress_blank on the first step, "timestamps never go back"
  * seek_delta = 2 * (ts - beg); a window completes when seek + seek_delta + 10 >= seek_end (delta_min = 100 ms)
  * segments split at timestamp boundaries, t0 / t1 = seek + 2 * (tid - beg); a single trailing timestamp consumes the
    whole window; a double one moves seek by the timestamp
  * step-219 repetition guard and the entropy threshold (2.4 over the last 32 ids) trigger the temperature fallback
  * the prompt ([prev] + past + [sot, lang, transcribe]) is dropped at temperature >= 0.5
  * a trailing window shorter than 5 s clears prompt_past
  * no_speech_prob > 0.6 with avg_logprob < -1 emits nothing and does not fall back

The CPU tests hold the ORACLE to the literals; the GPU tests hold the PRODUCT to the same literals."""
import numpy as np
import pytest

BEG, EOT, NOSP, BLANK = 50364, 50257, 50362, 32      # micro / multilingual v1-v2 vocabulary (51865 entries)
A, B, C_, D, E = 1000, 1001, 1002, 1003, 1004
DECOYS = list(range(2000, 2020))


class Script:
    """{(seek, i_temp): [entry per step]}; an entry is a token id (made the clear argmax) or {token: logit}."""

    def __init__(self, table):
        self.table = table
        self.calls = []

    def __call__(self, seek, it, step, dec, n_prompt, logits):
        self.calls.append((seek, it, step, n_prompt))
        seq = self.table.get((seek, it))
        logits[:] = -20.0
        entry = seq[step] if seq is not None and step < len(seq) else EOT     # an exhausted script ends the sequence
        if isinstance(entry, dict):
            for k, v in entry.items():
                logits[k] = v
        else:
            logits[entry] = 10.0
        return True


def flat(tok):       # the token wins by 0.1 over 20 decoys: log-probability -2.9495 (< logprob_thold -1)
    d = {k: 1.9 for k in DECOYS}
    d[tok] = 2.0
    return d


SCENARIOS = {
    # two windows; double timestamps move seek by the timestamp, a single trailing one consumes the window
    "pairs_and_seek": dict(
        script={(0, 0): [BEG, A, B, BEG + 100, BEG + 100, C_, BEG + 250, BEG + 250, EOT], (500, 0): [BEG, D, BEG + 1250]},
        segments=[(0, 200, [BEG, A, B, BEG + 100]), (200, 500, [C_, BEG + 250]), (500, 3000, [BEG, D, BEG + 1250])],
        windows=2, fallbacks=0,
        calls=[(0, 0, s, 3) for s in range(9)] + [(500, 0, s, 12) for s in range(3)]),
    # has_ts with seek_delta 400: timestamps below beg+200 are masked, so the preferred beg+150 loses to E
    "timestamps_never_go_back": dict(
        script={(0, 0): [BEG, A, BEG + 200, BEG + 200, B, {BEG + 150: 10.0, E: 5.0}, BEG + 1495]},
        segments=[(0, 400, [BEG, A, BEG + 200]), (400, 2990, [B, E, BEG + 1495])],
        windows=1, fallbacks=0, calls=[(0, 0, s, 3) for s in range(7)]),
    # 220 tokens without a timestamp: failed at step 219, next temperature
    "repetition_guard_step_219": dict(
        script={(0, 0): [BEG] + [2000 + (i % 50) for i in range(219)], (0, 1): [BEG, A, BEG + 1495]},
        segments=[(0, 2990, [BEG, A, BEG + 1495])],
        windows=1, fallbacks=1, calls=[(0, 0, s, 3) for s in range(220)] + [(0, 1, s, 3) for s in range(3)]),
    # 31 x A + one timestamp in the last 32 ids: entropy 0.139 < 2.4 -> fallback; at t = 0.6 the prompt is dropped
    "entropy_fallback_and_prompt_drop": dict(
        prompt="hello world",
        script={(0, it): [BEG] + [A] * 40 + [BEG + 1495] for it in range(3)} | {(0, 3): [BEG, B, BEG + 1495]},
        segments=[(0, 2990, [BEG, B, BEG + 1495])],
        windows=1, fallbacks=3,
        calls=[(0, it, s, "full") for it in range(3) for s in range(42)] + [(0, 3, s, 3) for s in range(3)]),
    # no_speech_prob 0.999 and avg_logprob -2.95: nothing is emitted and nothing falls back
    "no_speech_skip": dict(
        script={(0, 0): [flat(BEG) | {NOSP: 12.0}, flat(A), flat(BEG + 1495)]},
        segments=[], windows=1, fallbacks=0, calls=[(0, 0, s, 3) for s in range(3)]),
    # the window that starts at 26 s has under 5 s left: prompt_past is cleared (3 prompt rows, not 1 + 4 + 3)
    "trailing_window_clears_prompt_past": dict(
        script={(0, 0): [BEG, A, BEG + 1300, BEG + 1300, EOT], (2600, 0): [BEG, B, BEG + 195]},
        segments=[(0, 2600, [BEG, A, BEG + 1300]), (2600, 2990, [BEG, B, BEG + 195])],
        windows=2, fallbacks=0, calls=[(0, 0, s, 3) for s in range(5)] + [(2600, 0, s, 3) for s in range(3)]),
    # first step: timestamps above 1.0 s (beg + 50) are masked.  The opening timestamp beg+50 is > beg, so it is a segment
    # boundary with no text in front of it: nothing is emitted for it and the segment that follows starts at 2 * 50 = 100
    "max_initial_timestamp": dict(
        script={(0, 0): [{BEG + 51: 10.0, BEG + 50: 5.0}, A, BEG + 1495]},
        segments=[(100, 2990, [A, BEG + 1495])],
        windows=1, fallbacks=0, calls=[(0, 0, s, 3) for s in range(3)]),
    # first step: " " and EOT are masked (suppress_blank, whisper.rs:121)
    "suppress_blank_first_step": dict(
        script={(0, 0): [{BLANK: 10.0, EOT: 9.0, A: 5.0}, BEG + 1495]},
        segments=[(0, 2990, [A, BEG + 1495])],
        windows=1, fallbacks=0, calls=[(0, 0, s, 3) for s in range(2)]),
}


def pcm30():
    from nobs_whisper_b200 import synth_audio
    return synth_audio.synth_clip(5, 30.0)          # 480 000 samples: seek_end = n_len_org = 2999


def expected_calls(sc, n_prompt_tokens):
    full = 1 + n_prompt_tokens + 3
    return [(a, b, c, full if d == "full" else d) for a, b, c, d in sc["calls"]]


@pytest.fixture(scope="module")
def micro_path(model_dir):
    from nobs_whisper_b200 import ggml_synth
    return ggml_synth.ensure_model(model_dir, "micro")


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_follows_the_rules(micro_path, name):
    from oracle import oracle
    sc = SCENARIOS[name]
    orc = oracle.Oracle(micro_path)
    assert (orc.token_beg, orc.token_eot, orc.token_nosp, orc.n_text_ctx) == (BEG, EOT, NOSP, 448) and orc.token_bytes(BLANK) == b" "
    script = Script(sc["script"])
    orc.set_logits_hook(script)
    s0 = orc.stats()
    segs = orc.full(oracle.reference_params("en", initial_prompt=sc.get("prompt")), pcm30())
    s1 = orc.stats()
    assert [(s["t0"], s["t1"], s["tokens"]) for s in segs] == sc["segments"]
    for s in segs:
        assert s["text"] == b"".join(orc.token_bytes(t) for t in s["tokens"] if t < EOT)
    assert s1["n_encode"] - s0["n_encode"] == sc["windows"]
    assert s1["n_fail_p"] - s0["n_fail_p"] == sc["fallbacks"]
    n_p = len(orc.tokenize(sc["prompt"])) if sc.get("prompt") else 0
    assert script.calls == expected_calls(sc, n_p)
    orc.set_logits_hook(None)
    orc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_product_follows_the_rules(micro_path, name, precision):
    import nobs_whisper_b200 as nw
    sc = SCENARIOS[name]
    ctx = nw.WhisperContext.new_with_params(micro_path, nw.WhisperContextParameters.default(), precision=precision)
    assert (ctx.token_beg(), ctx.token_eot(), ctx.token_nosp()) == (BEG, EOT, NOSP) and ctx.token_to_bytes(BLANK) == b" "
    script = Script(sc["script"])
    ctx.set_logits_hook(script)
    p = nw.FullParams.new(nw.SamplingStrategy.Greedy(best_of=1))
    p.set_language("en")
    if sc.get("prompt"):
        p.set_initial_prompt(sc["prompt"])
    p.set_no_context(False); p.set_suppress_blank(True); p.set_no_speech_thold(0.6); p.set_entropy_thold(2.4); p.set_logprob_thold(-1.0)
    st = ctx.create_state()
    st.full(p, pcm30())
    assert [(s["t0"], s["t1"], s["tokens"]) for s in st.segments()] == sc["segments"]
    for s in st.segments():
        assert s["text"] == b"".join(ctx.token_to_bytes(t) for t in s["tokens"] if t < EOT)
    stats = st.stats()
    assert stats.n_windows == sc["windows"] and stats.n_fallbacks == sc["fallbacks"]
    n_p = len(ctx.tokenize(sc["prompt"])) if sc.get("prompt") else 0
    assert script.calls == expected_calls(sc, n_p)
    ctx.set_logits_hook(None)
    st.close()
    ctx.close()
