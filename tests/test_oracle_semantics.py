"""CPU tests (no GPU): control-flow semantics of the oracle's `full`."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def orc(model_dir):
    from nobs_whisper_b200 import ggml_synth
    from oracle import oracle
    o = oracle.Oracle(ggml_synth.ensure_model(model_dir, "micro", init="fanin"))
    yield o
    o.close()


def test_special_tokens_and_vocab(orc):
    assert (orc.token_eot, orc.token_sot, orc.token_translate, orc.token_transcribe) == (50257, 50258, 50358, 50359)
    assert (orc.token_solm, orc.token_prev, orc.token_nosp, orc.token_not, orc.token_beg) == (50360, 50361, 50362, 50363, 50364)
    assert orc.token_bytes(32) == b" " and orc.token_bytes(orc.token_beg) == b"[_BEG_]" and orc.token_bytes(orc.token_sot + 1) == b"[_LANG_en]"


def test_under_one_second_returns_no_segments(orc):
    from oracle import oracle
    assert orc.full(oracle.reference_params("en"), np.zeros(15999 // 2, np.float32)) == []


def test_logit_rules(orc):
    from oracle import oracle
    p = oracle.reference_params("en")
    rng = np.random.default_rng(0)
    logits = rng.standard_normal(orc.n_vocab).astype(np.float32)
    beg, eot = orc.token_beg, orc.token_eot
    # initial step: blank / eot / specials suppressed, timestamps limited to <= 1 s (beg+50)
    lp, pr = orc.process_logits(p, logits, [], False, 3000, 0.0)
    assert lp[eot] == -np.inf and lp[32] == -np.inf and lp[orc.token_sot] == -np.inf and lp[orc.token_nosp] == -np.inf
    assert np.all(np.isinf(lp[orc.token_sot + 1: orc.token_sot + 101]))
    assert np.all(np.isinf(lp[beg + 51:])) and np.isfinite(lp[beg + 50])
    # after a single timestamp (penultimate missing counts as timestamp): no more timestamps
    lp, _ = orc.process_logits(p, logits, [beg + 10], True, 20, 0.0)
    assert np.all(np.isinf(lp[beg:])) and np.isfinite(lp[100])
    # after text + timestamp: text is forbidden, timestamps may not go back in time
    lp, _ = orc.process_logits(p, logits, [100, beg + 40], True, 80, 0.0)
    assert np.all(np.isinf(lp[:eot])) and np.all(np.isinf(lp[beg: beg + 40])) and np.isfinite(lp[beg + 40])
    # timestamp mass beats the best text token -> text suppressed
    boosted = logits.copy()
    boosted[beg: beg + 51] += 6.0
    lp, pr = orc.process_logits(p, boosted, [], False, 3000, 0.0)
    assert np.all(np.isinf(lp[:beg])) and pr[:beg].sum() == 0.0
    # temperature only rescales
    lp_t, _ = orc.process_logits(p, logits, [300], False, 3000, 0.5)
    lp_1, _ = orc.process_logits(p, logits * 2.0, [300], False, 3000, 0.0)
    fin = np.isfinite(lp_1)
    assert np.allclose(lp_t[fin], lp_1[fin], atol=1e-5)


def test_full_is_deterministic_and_well_formed(orc):
    from nobs_whisper_b200 import synth_audio
    from oracle import oracle
    pcm = synth_audio.synth_clip(3, 8.0)
    a = orc.full(oracle.reference_params("en"), pcm)
    b = orc.full(oracle.reference_params("en"), pcm)
    assert a == b and len(a) > 0
    for s in a:
        assert s["t0"] <= s["t1"] and all(0 <= t < orc.n_vocab for t in s["tokens"])
        assert s["text"] == b"".join(orc.token_bytes(t) for t in s["tokens"] if t < orc.token_eot)
    st = orc.stats()
    assert st["n_encode"] >= 2 and st["n_decode_tokens"] > 0


def test_beam_and_prompt_paths_run(orc):
    from nobs_whisper_b200 import synth_audio
    from oracle import oracle
    pcm = synth_audio.synth_clip(4, 6.0)
    out = orc.full(oracle.reference_params("en", initial_prompt="Claude Code, Anthropic", beam_size=3, temperature_inc=0.0), pcm)
    assert isinstance(out, list)
    lid, probs = orc.lang_detect()
    assert 0 <= lid < 100 and abs(float(probs.sum()) - 1.0) < 1e-4
