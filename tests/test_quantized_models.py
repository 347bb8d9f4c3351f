"""Quantised ggml files (the reference catalogue lists q5_0 / q5_1 / q8_0 models,
src-tauri/src/model.rs:155-186).  A quantised file and its float32 dequantisation (written by numpy from the
published block layouts) are the same model: every consumer must give identical results on both."""
import numpy as np
import pytest

from nobs_whisper_b200 import ggml_synth, synth_audio

FTYPES = {"q4_0": 2, "q4_1": 3, "q8_0": 7, "q5_0": 8, "q5_1": 9}


def test_block_layouts_round_trip():
    """Known answers for the block formats: hand-built blocks dequantise to the values the layout defines."""
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((8, 64)) * 0.05).astype(np.float32)
    for name, ft in FTYPES.items():
        tt = ggml_synth.QUANT_TTYPE[ft]
        raw = ggml_synth.quantize_blocks(w, tt)
        assert len(raw) == w.size // 32 * {2: 18, 3: 20, 6: 22, 7: 24, 8: 34}[tt], name
        y = ggml_synth.dequantize_blocks(raw, tt, w.size)
        assert np.abs(y - w.ravel()).max() < {2: 0.09, 3: 0.06, 6: 0.045, 7: 0.03, 8: 0.005}[tt] * np.abs(w).max(), name
    # q5_0, one block by hand: d = 0.5, qh bit i = fifth bit of weight i, low nibbles first
    q = np.arange(32, dtype=np.uint8)                    # 5-bit codes 0..31
    qs = (q[:16] & 15) | ((q[16:] & 15) << 4)
    qh = np.uint32(sum(int((q[i] >> 4) & 1) << i for i in range(32)))
    blk = np.float16(0.5).tobytes() + qh.tobytes() + qs.tobytes()
    assert np.array_equal(ggml_synth.dequantize_blocks(blk, 6, 32), (np.arange(32) - 16).astype(np.float32) * 0.5)
    # q8_0
    blk = np.float16(0.25).tobytes() + np.arange(-16, 16, dtype=np.int8).tobytes()
    assert np.array_equal(ggml_synth.dequantize_blocks(blk, 8, 32), np.arange(-16, 16).astype(np.float32) * 0.25)


@pytest.mark.parametrize("name", ["q5_0", "q5_1", "q8_0", "q4_0", "q4_1"])
def test_oracle_reads_quantised_files(model_dir, name):
    """The CPU oracle on the quantised file == the oracle on the float32 dequantisation, bit for bit."""
    from oracle import oracle
    ft = FTYPES[name]
    pq = ggml_synth.ensure_model(model_dir, "micro", ftype=ft, init="fanin")
    pd = ggml_synth.ensure_model(model_dir, "micro", ftype=ft, init="fanin", dequantized=True)
    pcm = synth_audio.synth_clip(3, 4.0)
    outs = []
    for p in (pq, pd):
        o = oracle.Oracle(p)
        o.mel(pcm)
        enc = o.encode(0)
        logits = o.decode([50258, 50259, 50359], 0, 0)
        outs.append((enc.copy(), logits.copy()))
        o.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["q5_0", "q5_1", "q8_0", "q4_0", "q4_1"])
def test_engine_loads_quantised_files(model_dir, name):
    """fp32 parity mode: quantised file == its dequantisation (bit-exact encoder output and transcript), and the
    encoder output agrees with the oracle within the fp32 tolerance (1e-4 relative)."""
    import nobs_whisper_b200 as nw
    from conftest import rel_err
    from oracle import oracle
    ft = FTYPES[name]
    pq = ggml_synth.ensure_model(model_dir, "micro", ftype=ft, init="fanin")
    pd = ggml_synth.ensure_model(model_dir, "micro", ftype=ft, init="fanin", dequantized=True)
    pcm = synth_audio.synth_clip(3, 6.0)
    res = []
    for p in (pq, pd):
        ctx = nw.WhisperContext.new_with_params(p, nw.WhisperContextParameters.default(), precision="fp32")
        st = ctx.create_state()
        st.pcm_to_mel(pcm)
        st.encode(0)
        enc = st.encoder_output().copy()
        prm = nw.FullParams.new(nw.SamplingStrategy.Greedy(best_of=1))
        prm.set_language("en")
        st.full(prm, pcm)
        res.append((enc, st.segments()))
        st.close(); ctx.close()
    assert np.array_equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1]
    o = oracle.Oracle(pq)
    o.mel(pcm)
    assert rel_err(res[0][0], o.encode(0)) < 1e-4
    o.close()


@pytest.mark.gpu
def test_bf16_engine_on_a_q5_0_model(model_dir):
    import nobs_whisper_b200 as nw
    from conftest import rel_err
    from oracle import oracle
    p = ggml_synth.ensure_model(model_dir, "micro", ftype=8, init="fanin")
    pcm = synth_audio.synth_clip(5, 8.0)
    ctx = nw.WhisperContext.new_with_params(p, nw.WhisperContextParameters.default(), precision="bf16")
    st = ctx.create_state()
    st.pcm_to_mel(pcm)
    st.encode(0)
    o = oracle.Oracle(p)
    o.mel(pcm)
    assert rel_err(st.encoder_output(), o.encode(0)) < 2e-2
    st.close(); ctx.close(); o.close()


def test_loader_accepts_quantised_files_and_rejects_damaged_ones(model_dir, tmp_path):
    """The product's ggml reader without a GPU (host-only handle): every block format parses; truncated block data,
    an unknown tensor type and an unknown file type are load errors (-> InitError, whisper.rs:41-45)."""
    import struct

    import nobs_whisper_b200 as nw
    for name, ft in FTYPES.items():
        p = ggml_synth.ensure_model(model_dir, "micro", ftype=ft, init="fanin")
        ctx = nw.WhisperContext(p, host_only=True)
        assert ctx.n_vocab() == 51865
        ctx.close()
    good = open(ggml_synth.ensure_model(model_dir, "micro", ftype=8, init="fanin"), "rb").read()
    (tmp_path / "cut.bin").write_bytes(good[: len(good) - 777])
    with pytest.raises(nw.WhisperError):
        nw.WhisperContext(str(tmp_path / "cut.bin"), host_only=True)
    bad_ftype = good[:4] + struct.pack("<11i", *struct.unpack("<11i", good[4:48])[:10], 5) + good[48:]
    (tmp_path / "ftype.bin").write_bytes(bad_ftype)
    with pytest.raises(nw.WhisperError):
        nw.WhisperContext(str(tmp_path / "ftype.bin"), host_only=True)
