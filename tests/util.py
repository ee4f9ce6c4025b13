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
