"""the reference's chunk-to-chunk context chaining (whisper.rs:152-197, state.rs:757-792) under data
parallelism.  `transcribe_chunked_parallel` decodes windows of chunks together with speculated contexts and re-decodes what
turned out wrong; the result must be EXACTLY what the reference's sequential loop gives (fp32 parity mode: token-identical)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

VOCAB = "Claude Code, Anthropic, Supabase"


def recording(seed, parts):
    """speech-like pieces separated by 1-s silences (the audio.rs test idiom), as separate chunks"""
    from nobs_whisper_b200 import synth_audio
    return [synth_audio.synth_clip(seed + i, s) for i, s in enumerate(parts)]


@pytest.fixture(scope="module")
def engine(model_dir):
    import os
    import nobs_whisper_b200 as nw
    from nobs_whisper_b200 import ggml_synth
    os.environ["NOBS_WHISPER_PRECISION"] = "fp32"
    try:
        e = nw.WhisperEngine.from_file(ggml_synth.ensure_model(model_dir, "micro", init="fanin"))
    finally:
        del os.environ["NOBS_WHISPER_PRECISION"]
    yield e
    e.close()


@pytest.mark.parametrize("seed,parts,vocab", [(500, [12.0, 7.0, 30.0, 4.0, 18.0], None), (520, [30.0] * 9, VOCAB), (540, [5.0, 0.05, 9.0, 30.0, 0.3, 11.0, 6.0], VOCAB)])
def test_chained_parallel_equals_the_sequential_loop(engine, seed, parts, vocab):
    chunks = recording(seed, parts)
    want = engine.transcribe_chunked(chunks, "en", vocab)
    got, n_decodes, n_rounds = engine.transcribe_chunked_parallel(chunks, "en", vocab)
    assert got == want
    assert len(want) > 0
    # every chunk is decoded at least once, and speculation is bounded (window halving): at most ~3 decodes per chunk
    assert len(chunks) <= n_decodes <= 3 * len(chunks) + 2
    assert 1 <= n_rounds <= len(chunks)
    print(f"{len(chunks)} chunks: {n_decodes} chunk decodes in {n_rounds} batched rounds")


def test_silent_chunks_confirm_whole_windows(engine):
    """Chunks that come back empty do not change the context of their successors: one round confirms them all."""
    chunks = [np.zeros(16000 * 3, np.float32) for _ in range(6)]
    want = engine.transcribe_chunked(chunks, "en", None)
    got, n_decodes, n_rounds = engine.transcribe_chunked_parallel(chunks, "en", None)
    assert got == want
    if want == "":
        assert (n_decodes, n_rounds) == (6, 1)


def test_recording_chained_mode(engine):
    """state.rs:757-792 on a 100-s recording with pauses: parallel='chained' gives the text of the reference's loop."""
    from nobs_whisper_b200 import synth_audio
    parts = []
    for i, s in enumerate([21.0, 17.0, 26.0, 14.0, 19.0]):
        parts.append(synth_audio.synth_clip(600 + i, s))
        parts.append(np.zeros(16000, np.float32))
    audio = np.concatenate(parts)
    want = engine.transcribe_recording(audio, "en", VOCAB, parallel=False)
    assert engine.transcribe_recording(audio, "en", VOCAB, parallel="chained") == want
    assert len(want) > 0
