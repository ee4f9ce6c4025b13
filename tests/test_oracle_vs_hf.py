"""Pins the oracle's VITS blocks against the independently written HuggingFace implementation
(transformers/models/vits/modeling_vits.py) by weight copy: HiFi-GAN generator, deterministic and
stochastic duration predictors, WN residual coupling layer, one relative-attention encoder layer.
The reference itself has no golden vectors; this is the strongest offline anchor.
"""
import math

import numpy as np
import pytest
import torch
from transformers import VitsConfig
from transformers.models.vits import modeling_vits as mv

from util import ov

HP = ov.HParams()
CFG = VitsConfig(speaker_embedding_size=HP.gin_channels, upsample_rates=list(HP.upsample_rates),
                 upsample_kernel_sizes=list(HP.upsample_kernel_sizes), hidden_dropout=0.0, attention_dropout=0.0,
                 activation_dropout=0.0, duration_predictor_dropout=0.0)


def rand_init(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / math.sqrt(max(1, p[0].numel()))))
            elif n.endswith("gamma"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return m.eval()


def copy_ln(dst, src):
    dst.weight.data.copy_(src.gamma.data)
    dst.bias.data.copy_(src.beta.data)


def copy_conv(dst, src):
    dst.weight.data.copy_(src.weight.data.reshape(dst.weight.shape))
    if src.bias is not None:
        dst.bias.data.copy_(src.bias.data)


def copy_dds(dst, src):
    for i in range(len(src.convs_sep)):
        copy_conv(dst.convs_dilated[i], src.convs_sep[i])
        copy_conv(dst.convs_pointwise[i], src.convs_1x1[i])
        copy_ln(dst.norms_1[i], src.norms_1[i])
        copy_ln(dst.norms_2[i], src.norms_2[i])


def copy_weight_normed(dst, src):
    """dst has torch.nn.utils.parametrizations.weight_norm: weight = g * v / ||v||."""
    w = src.weight.data
    dst.parametrizations.weight.original1.data.copy_(w)
    dst.parametrizations.weight.original0.data.copy_(w.norm(dim=tuple(range(1, w.dim())), keepdim=True))
    dst.bias.data.copy_(src.bias.data)


@torch.no_grad()
def test_hifigan_generator_matches_hf():
    mine = rand_init(ov.Generator(HP), 1)
    hf = mv.VitsHifiGan(CFG).eval()
    copy_conv(hf.conv_pre, mine.conv_pre)
    copy_conv(hf.cond, mine.cond)
    hf.conv_post.weight.data.copy_(mine.conv_post.weight.data)
    for i in range(len(mine.ups)):
        copy_conv(hf.upsampler[i], mine.ups[i])
    for i in range(len(mine.resblocks)):
        for l in range(3):
            copy_conv(hf.resblocks[i].convs1[l], mine.resblocks[i].convs1[l])
            copy_conv(hf.resblocks[i].convs2[l], mine.resblocks[i].convs2[l])
    g = torch.Generator().manual_seed(3)
    z = torch.randn(2, HP.inter_channels, 24, generator=g)
    spk = torch.randn(2, HP.gin_channels, 1, generator=g) * 0.1
    a, b = mine(z, spk), hf(z, spk)
    assert a.shape == b.shape == (2, 1, 24 * 512)
    assert float((a - b).abs().max()) < 1e-5


@torch.no_grad()
def test_duration_predictor_matches_hf():
    mine = rand_init(ov.DurationPredictor(HP.hidden_channels, HP.dp_filter_channels, 3, HP.gin_channels), 2)
    hf = mv.VitsDurationPredictor(CFG).eval()
    for n in ("conv_1", "conv_2", "proj", "cond"):
        copy_conv(getattr(hf, n), getattr(mine, n))
    copy_ln(hf.norm_1, mine.norm_1)
    copy_ln(hf.norm_2, mine.norm_2)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, HP.hidden_channels, 31, generator=g)
    mask = torch.ones(2, 1, 31)
    mask[1, :, 20:] = 0
    spk = torch.randn(2, HP.gin_channels, 1, generator=g)
    assert float((mine(x, mask, g=spk) - hf(x, mask, spk)).abs().max()) < 1e-5


@torch.no_grad()
def test_stochastic_duration_predictor_matches_hf():
    mine = rand_init(ov.StochasticDurationPredictor(HP.hidden_channels, 3, HP.sdp_n_flows, HP.gin_channels), 5)
    hf = mv.VitsStochasticDurationPredictor(CFG).eval()
    copy_conv(hf.conv_pre, mine.pre)
    copy_conv(hf.conv_proj, mine.proj)
    copy_conv(hf.cond, mine.cond)
    copy_dds(hf.conv_dds, mine.convs)
    hf.flows[0].translate.data.copy_(mine.flows[0].m.data)
    hf.flows[0].log_scale.data.copy_(mine.flows[0].logs.data)
    for j in range(1, HP.sdp_n_flows + 1):
        src = mine.flows[2 * j - 1]
        copy_conv(hf.flows[j].conv_pre, src.pre)
        copy_conv(hf.flows[j].conv_proj, src.proj)
        copy_dds(hf.flows[j].conv_dds, src.convs)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, HP.hidden_channels, 29, generator=g)
    mask = torch.ones(2, 1, 29)
    spk = torch.randn(2, HP.gin_channels, 1, generator=g)
    torch.manual_seed(77)
    noise = torch.randn(2, 2, 29)
    torch.manual_seed(77)
    b = hf(x, mask, spk, reverse=True, noise_scale=0.8)
    a = mine.forward_reverse(x, mask, spk, noise * 0.8)
    assert a.shape == b.shape == (2, 1, 29)
    assert float((a - b).abs().max()) < 2e-4, float((a - b).abs().max())


@torch.no_grad()
def test_wn_coupling_layer_matches_hf():
    hp = ov.HParams(use_transformer_flow=False)
    mine = rand_init(ov.ResidualCouplingLayer(hp), 7)
    hf = mv.VitsResidualCouplingLayer(CFG).eval()
    copy_conv(hf.conv_pre, mine.pre)
    copy_conv(hf.conv_post, mine.post)
    copy_weight_normed(hf.wavenet.cond_layer, mine.enc.cond_layer)
    for i in range(hp.wn_layers):
        copy_weight_normed(hf.wavenet.in_layers[i], mine.enc.in_layers[i])
        copy_weight_normed(hf.wavenet.res_skip_layers[i], mine.enc.res_skip_layers[i])
    g = torch.Generator().manual_seed(8)
    z = torch.randn(2, hp.inter_channels, 40, generator=g)
    mask = torch.ones(2, 1, 40)
    mask[0, :, 33:] = 0
    spk = torch.randn(2, hp.gin_channels, 1, generator=g)
    a = mine.forward_reverse(z * mask, mask, spk)
    b, _ = hf(z * mask, mask, spk, reverse=True)
    assert float((a - b).abs().max()) < 1e-5


@torch.no_grad()
def test_encoder_layer_matches_hf():
    enc = rand_init(ov.Encoder(HP.hidden_channels, HP.filter_channels, HP.n_heads, 1, HP.kernel_size, HP.window_size), 9)
    hf = mv.VitsEncoderLayer(CFG).eval()
    at = enc.attn_layers[0]
    for dst, src in ((hf.attention.q_proj, at.conv_q), (hf.attention.k_proj, at.conv_k), (hf.attention.v_proj, at.conv_v),
                     (hf.attention.out_proj, at.conv_o)):
        copy_conv(dst, src)
    hf.attention.emb_rel_k.data.copy_(at.emb_rel_k.data)
    hf.attention.emb_rel_v.data.copy_(at.emb_rel_v.data)
    copy_ln(hf.layer_norm, enc.norm_layers_1[0])
    copy_ln(hf.final_layer_norm, enc.norm_layers_2[0])
    copy_conv(hf.feed_forward.conv_1, enc.ffn_layers[0].conv_1)
    copy_conv(hf.feed_forward.conv_2, enc.ffn_layers[0].conv_2)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(2, HP.hidden_channels, 37, generator=g)
    mask = torch.ones(2, 1, 37)
    a = enc(x, mask)
    b = hf(x.transpose(1, 2), mask.transpose(1, 2))[0].transpose(1, 2)
    assert float((a - b).abs().max()) < 2e-5, float((a - b).abs().max())


@torch.no_grad()
def test_deberta_oracle_is_hf_layer_minus_3():
    from oracle import deberta as od
    cfg = od.tiny_config()
    m = od.build_model(cfg, seed=1)
    ids = torch.randint(3, cfg.vocab_size, (2, 9), generator=torch.Generator().manual_seed(0))
    out = od.predict(m, ids, torch.ones_like(ids))
    full = m(input_ids=ids, attention_mask=torch.ones_like(ids), output_hidden_states=True).hidden_states
    assert len(full) == cfg.num_hidden_layers + 1
    assert torch.equal(out, full[cfg.num_hidden_layers - 2])  # output of encoder layer L-2 (22 of 24)
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "deberta_tiny_s7.npz"))
    o2 = od.predict(m, torch.from_numpy(g["ids"]), torch.ones(1, 7, dtype=torch.long))
    np.testing.assert_allclose(o2[0].numpy(), g["out"], atol=2e-5)
