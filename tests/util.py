"""Shared helpers for the tests: synthetic models (oracle weights -> ONNX bytes) and inputs."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))

from oracle import vits as ov  # noqa: E402


def synth_assets(hp: ov.HParams, seed: int = 0, **kw):
    """-> (oracle model, model.onnx bytes)"""
    from sbv2_b200 import assets
    model = ov.build_model(hp, seed=seed, **kw)
    onnx = assets.synth_onnx(ov.state_dict_numpy(model), hp.upsample_rates, hp.resblock_dilation_sizes,
                             anonymize_weight_norm=kw.get("anonymize", False))
    return model, onnx


def make_utterance(hp: ov.HParams, t_x: int, seed: int, sdp_ratio: float = 0.0, length_scale: float = 1.0, sid: int = 0,
                   max_frames: int = 0):
    x, tone, lang, bert, style = ov.synthetic_inputs(hp, t_x, seed)
    g = torch.Generator().manual_seed(seed + 1000)
    noise_sdp = torch.randn(1, 2, t_x, generator=g)
    frames = max_frames or int(t_x * 8 * length_scale + 64)
    noise_zp = torch.randn(1, hp.inter_channels, frames, generator=g)
    return dict(x=x, tone=tone, lang=lang, bert=bert, style=style, noise_sdp=noise_sdp, noise_zp=noise_zp, t_x=t_x,
                sdp_ratio=sdp_ratio, length_scale=length_scale, sid=sid, noise_scale=0.677, noise_scale_w=0.8)


def oracle_run(model, u, return_intermediates=True, dtype=torch.float32):
    m = model if dtype == torch.float32 else model.double()
    cast = lambda t: t.to(dtype) if t.is_floating_point() else t
    out = m.infer(u["x"], torch.tensor([u["t_x"]]), torch.tensor([u["sid"]]), u["tone"], u["lang"], cast(u["bert"]),
                  cast(u["style"]), noise_sdp=cast(u["noise_sdp"]), noise_zp=cast(u["noise_zp"]),
                  noise_scale=u["noise_scale"], length_scale=u["length_scale"], noise_scale_w=u["noise_scale_w"],
                  sdp_ratio=u["sdp_ratio"], return_intermediates=return_intermediates)
    if dtype != torch.float32:
        model.float()
    return out


def to_api(u):
    """dict for sbv2_b200.Model.synthesize_batch"""
    return dict(bert=u["bert"][0].numpy(), x_tst=u["x"][0].numpy(), tones=u["tone"][0].numpy(), lang_ids=u["lang"][0].numpy(),
                style_vec=u["style"][0].numpy(), sid=u["sid"], sdp_ratio=u["sdp_ratio"], length_scale=u["length_scale"],
                noise_scale=u["noise_scale"], noise_scale_w=u["noise_scale_w"], noise_sdp=u["noise_sdp"][0].numpy(),
                noise_zp=u["noise_zp"][0].numpy())


@torch.no_grad()
def oracle_durations(model, u, dtype=torch.float32):
    """The text half of ``infer`` only (enc_p -> SDP/DP -> ceil): what decides durations.  -> dict(w, w_ceil, logw).
    Same statements as oracle/vits.py SynthesizerTrn.infer up to ``w_ceil``."""
    m = model if dtype == torch.float32 else model.double()
    cast = lambda t: t.to(dtype) if t.is_floating_point() else t
    g = m.emb_g(torch.tensor([u["sid"]])).unsqueeze(-1)
    x, m_p, logs_p, x_mask = m.enc_p(u["x"], torch.tensor([u["t_x"]]), u["tone"], u["lang"], cast(u["bert"]), cast(u["style"]), g=g)
    logw_sdp = m.sdp.forward_reverse(x, x_mask, g, cast(u["noise_sdp"]) * u["noise_scale_w"])
    logw_dp = m.dp(x, x_mask, g=g)
    logw = logw_sdp * u["sdp_ratio"] + logw_dp * (1 - u["sdp_ratio"])
    w = torch.exp(logw) * x_mask * u["length_scale"]
    out = dict(w=w[0, 0].double().numpy(), w_ceil=torch.ceil(w)[0, 0].numpy().astype(np.int32), logw=logw[0, 0].double().numpy())
    if dtype != torch.float32:
        model.float()
    return out


def word2ph_for(t_x: int, seed: int):
    """A synthetic tokenisation of a t_x-phoneme utterance as the frontend produces it (tts_util.rs:120-160): CLS and SEP
    own one phoneme each, every other token 1..6 phonemes (sum == t_x).  -> int32 [t_tok]"""
    rng = np.random.default_rng(seed)
    inner = t_x - 2
    assert inner >= 1
    w = []
    left = inner
    while left > 0:
        k = int(min(left, rng.integers(1, 7)))
        w.append(k)
        left -= k
    return np.asarray([1] + w + [1], dtype=np.int32)


def expand_features(feat: np.ndarray, word2ph: np.ndarray) -> np.ndarray:
    """[t_tok, H] -> [H, t_x] (token row i repeated word2ph[i] times, transposed): tts_util.rs:129-154."""
    return np.ascontiguousarray(np.repeat(feat, word2ph, axis=0).T)
