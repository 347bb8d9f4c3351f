"""Generate tests/golden/hf_micro.npz: outputs of an INDEPENDENT implementation of the same
published model (HF transformers Whisper, CPU fp32) on the seeded synthetic `micro` weights
and synthetic clip 0.  Used to validate the oracle's arithmetic (mel, encoder, cross-KV-fed
decoder logits).  HF differs from whisper.cpp in two known ways that the test accounts for:
exact-erf GELU (the oracle has an erf switch used only for this cross-check) and reflect
padding on the right edge of the STFT (last 3 mel frames excluded).  SURVEY.md §8c.

Run in the build container (transformers is importable here; not needed on the GPU box):
    python tools/make_golden_hf.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nobs_whisper_b200 import ggml_synth, synth_audio  # noqa: E402

NAME_MAP = [
    ("blocks", "layers"), ("attn.query", "self_attn.q_proj"), ("attn.key", "self_attn.k_proj"),
    ("attn.value", "self_attn.v_proj"), ("attn.out", "self_attn.out_proj"), ("attn_ln", "self_attn_layer_norm"),
    ("cross_self_attn.q_proj", "encoder_attn.q_proj"), ("cross_self_attn.k_proj", "encoder_attn.k_proj"),
    ("cross_self_attn.v_proj", "encoder_attn.v_proj"), ("cross_self_attn.out_proj", "encoder_attn.out_proj"),
    ("cross_self_attn_layer_norm", "encoder_attn_layer_norm"), ("mlp.0", "fc1"), ("mlp.2", "fc2"),
    ("mlp_ln", "final_layer_norm"), ("encoder.ln_post", "encoder.layer_norm"), ("decoder.ln", "decoder.layer_norm"),
    ("token_embedding", "embed_tokens"),
]


def to_hf_name(n: str) -> str:
    if n == "encoder.positional_embedding":
        return "model.encoder.embed_positions.weight"
    if n == "decoder.positional_embedding":
        return "model.decoder.embed_positions.weight"
    for a, b in NAME_MAP:
        n = n.replace(a, b)
    return "model." + n


def whisper_cpp_style_mel(pcm: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """Independent float64 numpy statement of the log-mel the reference path computes
    (SURVEY.md §8a row a4): 200-sample reflect pad on the left, 30 s + 200 samples of zeros on
    the right, periodic Hann-400, hop 160, |rfft|^2, filterbank, log10 clamp 1e-10, global
    max-8 clamp, (x+4)/4.  Returns [n_mel, n_len]."""
    n = len(pcm)
    x = np.zeros(n + 480000 + 400, np.float64)
    x[200:200 + n] = pcm
    x[:200] = pcm[1:201][::-1]
    n_len = (len(x) - 400) // 160
    hann = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(400) / 400.0))
    idx = np.arange(n_len)[:, None] * 160 + np.arange(400)[None, :]
    spec = np.abs(np.fft.rfft(x[idx] * hann[None, :], axis=1)) ** 2  # [n_len, 201]
    mel = np.log10(np.maximum(filters.astype(np.float64) @ spec.T, 1e-10))
    mel = np.maximum(mel, mel.max() - 8.0)
    return ((mel + 4.0) / 4.0).astype(np.float32)


def main():
    from transformers import WhisperConfig, WhisperFeatureExtractor, WhisperForConditionalGeneration

    arch = ggml_synth.ARCHS["micro"]
    out = {}
    for init in ("survey", "fanin"):
        cfg = WhisperConfig(
            vocab_size=arch.n_vocab, num_mel_bins=arch.n_mels, encoder_layers=arch.n_audio_layer,
            encoder_attention_heads=arch.n_audio_head, decoder_layers=arch.n_text_layer,
            decoder_attention_heads=arch.n_text_head, d_model=arch.n_audio_state,
            encoder_ffn_dim=4 * arch.n_audio_state, decoder_ffn_dim=4 * arch.n_text_state,
            max_source_positions=arch.n_audio_ctx, max_target_positions=arch.n_text_ctx,
            activation_function="gelu", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
        )
        model = WhisperForConditionalGeneration(cfg).eval()
        sd = model.state_dict()
        used = set()
        for name, w in ggml_synth.generate_weights("micro", seed=0, init=init):
            hn = to_hf_name(name)
            t = torch.from_numpy(np.ascontiguousarray(w))
            if name.endswith("conv1.bias") or name.endswith("conv2.bias"):
                t = t.reshape(-1)
            assert hn in sd, (name, hn)
            assert sd[hn].shape == t.shape, (hn, sd[hn].shape, t.shape)
            sd[hn].copy_(t)
            used.add(hn)
        sd["proj_out.weight"].copy_(sd["model.decoder.embed_tokens.weight"])
        missing = [k for k in sd if k not in used and k != "proj_out.weight" and "k_proj.bias" not in k]
        assert not missing, missing
        for k in sd:
            if "k_proj.bias" in k:
                sd[k].zero_()

        pcm = synth_audio.synth_clip(0, 30.0)
        fe = WhisperFeatureExtractor(feature_size=arch.n_mels)
        fe.mel_filters = ggml_synth.slaney_filterbank(arch.n_mels).T.astype(np.float64)
        feats = fe(pcm, sampling_rate=16000, return_tensors="pt").input_features  # [1, 80, 3000]
        mel_w = whisper_cpp_style_mel(pcm, ggml_synth.slaney_filterbank(arch.n_mels))
        feats_w = torch.from_numpy(mel_w[None, :, :3000].copy())
        with torch.no_grad():
            enc_w = model.model.encoder(feats_w).last_hidden_state[0]
            logits_w = model(input_features=feats_w, decoder_input_ids=torch.tensor([[50258, 50259, 50359, 11, 22, 33]])).logits[0]
        out[f"{init}_enc_wmel"] = enc_w.numpy().astype(np.float32)
        out[f"{init}_logits_last_wmel"] = logits_w[-1].numpy().astype(np.float32)
        out["prompt"] = np.array([50258, 50259, 50359, 11, 22, 33], np.int32)
        if init == "survey":
            # HF's own extractor differs from the reference path only on the right edge
            # (reflect vs zero padding): keep its last frames to document that.
            out["mel_hf_first8"] = feats[0, :, :8].numpy().astype(np.float32)
            out["mel_hf_last8"] = feats[0, :, -8:].numpy().astype(np.float32)
            out["mel_wcpp"] = mel_w[:, :3000].copy()
            out["mel_wcpp_tail_minmax"] = np.array([mel_w[:, 3000:].min(), mel_w[:, 3000:].max()], np.float32)
    path = os.path.join(ROOT, "tests", "golden", "hf_micro.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
