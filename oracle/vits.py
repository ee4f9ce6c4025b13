"""CPU oracle for the synthesizer graph (Style-Bert-VITS2 JP-Extra ``SynthesizerTrn.infer``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the product
(``sbv2-api_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker.

PARITY UNPINNED: the reference (tuna2134/sbv2-api) has no tests, golden vectors or published
outputs for this path and neither ONNX Runtime nor the third-party Python
package ``style-bert-vits2`` (unpinned in /root/reference/scripts/convert/requirements.txt:1) that
defines the graph is available offline.  This file therefore restates the published algorithm of
that package's ``models_jp_extra.SynthesizerTrn.infer`` — the function that
/root/reference/scripts/convert/convert_model.py:89-155 traces into ``model.onnx`` and that
/root/reference/crates/sbv2_core/src/model.rs:53-111 runs through ``ort::Session::run``.  Blocks
shared with original VITS are cross-checked against the independently written HuggingFace
implementation (transformers/models/vits/modeling_vits.py) in tests/test_oracle_vs_hf.py.

Differences from the traced graph, all deliberate:
  * the two RNG sites (ONNX RandomNormalLike) are explicit inputs ``noise_sdp [B,2,T_x]`` and
    ``noise_zp [B,192,T_y]`` so that runs are reproducible across runtimes;
  * every intermediate is returned in a dict so kernels can be checked stage by stage.

Parameter names follow the upstream ``state_dict`` with weight-norm folded, which is
what an exported ``model.onnx`` carries as initializers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

LRELU_SLOPE = 0.1


@dataclass
class HParams:
    """Hyper-parameters of the "tsukuyomi-shaped" JP-Extra model."""

    n_vocab: int = 112
    num_tones: int = 12
    num_languages: int = 3
    inter_channels: int = 192
    hidden_channels: int = 192
    filter_channels: int = 768
    n_heads: int = 2
    n_layers: int = 6
    kernel_size: int = 3
    window_size: int = 4
    gin_channels: int = 512
    n_speakers: int = 1
    bert_dim: int = 1024
    style_dim: int = 256
    # flow
    use_transformer_flow: bool = True
    n_flow_layer: int = 4
    n_layers_trans_flow: int = 6
    flow_kernel_size: int = 5
    wn_layers: int = 4
    # duration predictors
    dp_filter_channels: int = 256
    sdp_n_flows: int = 4
    sdp_num_bins: int = 10
    sdp_tail_bound: float = 5.0
    # decoder
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    upsample_rates: Tuple[int, ...] = (8, 8, 2, 2, 2)
    upsample_initial_channel: int = 512
    upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 8, 2, 2)

    @property
    def hop(self) -> int:
        return int(np.prod(self.upsample_rates))


def tiny_hparams(**kw) -> HParams:
    """A reduced configuration used by fast CPU tests (same structure, fewer layers)."""
    base = dict(n_layers=3, n_layers_trans_flow=3, n_speakers=2)
    base.update(kw)
    return HParams(**base)


# ----------------------------------------------------------------------------------------------
# commons
# ----------------------------------------------------------------------------------------------

def sequence_mask(length: torch.Tensor, max_length: Optional[int] = None) -> torch.Tensor:
    if max_length is None:
        max_length = int(length.max())
    x = torch.arange(max_length, dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def generate_path(duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """duration [b,1,t_x], mask [b,1,t_y,t_x] -> one-hot path [b,1,t_y,t_x] (upstream commons)."""
    b, _, t_y, t_x = mask.shape
    cum_duration = torch.cumsum(duration, -1)
    cum_duration_flat = cum_duration.view(b * t_x)
    path = sequence_mask(cum_duration_flat, t_y).to(mask.dtype)
    path = path.view(b, t_x, t_y)
    path = path - F.pad(path, (0, 0, 1, 0, 0, 0))[:, :-1]
    path = path.unsqueeze(1).transpose(2, 3) * mask
    return path


def length_regulate_int(w: np.ndarray) -> Tuple[np.ndarray, int, np.ndarray]:
    """Integer restatement of ceil / cumsum / generate_path for one utterance.

    w: float32 [T_x] (already ``exp(logw)*mask*length_scale``).  Returns (durations int32 [T_x],
    T_y, frame2ph int32 [T_y]) with frame j owned by phoneme i iff cum[i-1] <= j < cum[i].
    When every duration is zero the graph clamps y_length to 1 and the single frame has an
    all-zero path row; frame2ph is -1 there.
    """
    d = np.ceil(w.astype(np.float32)).astype(np.int64)
    cum = np.cumsum(d)
    total = int(cum[-1]) if len(cum) else 0
    t_y = max(total, 1)
    j = np.arange(t_y)
    f2p = np.searchsorted(cum, j, side="right").astype(np.int32)
    f2p[j >= total] = -1
    return d.astype(np.int32), t_y, f2p


class LayerNorm(nn.Module):
    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        self.channels, self.eps = channels, eps
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))

    def forward(self, x):
        x = x.transpose(1, -1)
        x = F.layer_norm(x, (self.channels,), self.gamma, self.beta, self.eps)
        return x.transpose(1, -1)


# ----------------------------------------------------------------------------------------------
# attentions.Encoder (window-4 relative attention, conv FFN, speaker conditioning at layer 2)
# ----------------------------------------------------------------------------------------------

class MultiHeadAttention(nn.Module):
    def __init__(self, channels, out_channels, n_heads, window_size=4):
        super().__init__()
        assert channels % n_heads == 0
        self.channels, self.n_heads, self.window_size = channels, n_heads, window_size
        self.k_channels = channels // n_heads
        self.conv_q = nn.Conv1d(channels, channels, 1)
        self.conv_k = nn.Conv1d(channels, channels, 1)
        self.conv_v = nn.Conv1d(channels, channels, 1)
        self.conv_o = nn.Conv1d(channels, out_channels, 1)
        rel_stddev = self.k_channels ** -0.5
        self.emb_rel_k = nn.Parameter(torch.randn(1, window_size * 2 + 1, self.k_channels) * rel_stddev)
        self.emb_rel_v = nn.Parameter(torch.randn(1, window_size * 2 + 1, self.k_channels) * rel_stddev)

    def forward(self, x, c, attn_mask=None):
        q, k, v = self.conv_q(x), self.conv_k(c), self.conv_v(c)
        x = self.attention(q, k, v, attn_mask)
        return self.conv_o(x)

    def attention(self, query, key, value, mask=None):
        b, d, t_s, t_t = (*key.size(), query.size(2))
        query = query.view(b, self.n_heads, self.k_channels, t_t).transpose(2, 3)
        key = key.view(b, self.n_heads, self.k_channels, t_s).transpose(2, 3)
        value = value.view(b, self.n_heads, self.k_channels, t_s).transpose(2, 3)
        scores = torch.matmul(query / math.sqrt(self.k_channels), key.transpose(-2, -1))
        assert t_s == t_t
        key_rel = self._get_relative_embeddings(self.emb_rel_k, t_s)
        rel_logits = torch.matmul(query / math.sqrt(self.k_channels), key_rel.unsqueeze(0).transpose(-2, -1))
        scores = scores + self._relative_position_to_absolute_position(rel_logits)
        if mask is not None:
            scores = scores.masked_fill(mask == 0, -1e4)
        p_attn = F.softmax(scores, dim=-1)
        output = torch.matmul(p_attn, value)
        relative_weights = self._absolute_position_to_relative_position(p_attn)
        value_rel = self._get_relative_embeddings(self.emb_rel_v, t_s)
        output = output + torch.matmul(relative_weights, value_rel.unsqueeze(0))
        return output.transpose(2, 3).contiguous().view(b, d, t_t)

    def _get_relative_embeddings(self, relative_embeddings, length):
        pad_length = max(length - (self.window_size + 1), 0)
        slice_start = max((self.window_size + 1) - length, 0)
        slice_end = slice_start + 2 * length - 1
        if pad_length > 0:
            padded = F.pad(relative_embeddings, (0, 0, pad_length, pad_length, 0, 0))
        else:
            padded = relative_embeddings
        return padded[:, slice_start:slice_end]

    @staticmethod
    def _relative_position_to_absolute_position(x):
        batch, heads, length, _ = x.size()
        x = F.pad(x, (0, 1, 0, 0, 0, 0, 0, 0))
        x_flat = x.view([batch, heads, length * 2 * length])
        x_flat = F.pad(x_flat, (0, length - 1, 0, 0, 0, 0))
        return x_flat.view([batch, heads, length + 1, 2 * length - 1])[:, :, :length, length - 1:]

    @staticmethod
    def _absolute_position_to_relative_position(x):
        batch, heads, length, _ = x.size()
        x = F.pad(x, (0, length - 1, 0, 0, 0, 0, 0, 0))
        x_flat = x.view([batch, heads, length ** 2 + length * (length - 1)])
        x_flat = F.pad(x_flat, (length, 0, 0, 0, 0, 0))
        return x_flat.view([batch, heads, length, 2 * length])[:, :, :, 1:]


class FFN(nn.Module):
    def __init__(self, in_channels, out_channels, filter_channels, kernel_size):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv_1 = nn.Conv1d(in_channels, filter_channels, kernel_size)
        self.conv_2 = nn.Conv1d(filter_channels, out_channels, kernel_size)

    def _pad(self, x):
        if self.kernel_size == 1:
            return x
        return F.pad(x, ((self.kernel_size - 1) // 2, self.kernel_size // 2))

    def forward(self, x, x_mask):
        x = self.conv_1(self._pad(x * x_mask))
        x = torch.relu(x)
        x = self.conv_2(self._pad(x * x_mask))
        return x * x_mask


class Encoder(nn.Module):
    def __init__(self, hidden_channels, filter_channels, n_heads, n_layers, kernel_size,
                 window_size=4, gin_channels=0, cond_layer_idx=2):
        super().__init__()
        self.n_layers = n_layers
        self.cond_layer_idx = n_layers
        if gin_channels != 0:
            self.spk_emb_linear = nn.Linear(gin_channels, hidden_channels)
            self.cond_layer_idx = cond_layer_idx
            assert self.cond_layer_idx < n_layers
        self.attn_layers = nn.ModuleList()
        self.norm_layers_1 = nn.ModuleList()
        self.ffn_layers = nn.ModuleList()
        self.norm_layers_2 = nn.ModuleList()
        for _ in range(n_layers):
            self.attn_layers.append(MultiHeadAttention(hidden_channels, hidden_channels, n_heads, window_size))
            self.norm_layers_1.append(LayerNorm(hidden_channels))
            self.ffn_layers.append(FFN(hidden_channels, hidden_channels, filter_channels, kernel_size))
            self.norm_layers_2.append(LayerNorm(hidden_channels))

    def forward(self, x, x_mask, g=None):
        attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
        x = x * x_mask
        for i in range(self.n_layers):
            if i == self.cond_layer_idx and g is not None:
                gg = self.spk_emb_linear(g.transpose(1, 2)).transpose(1, 2)
                x = (x + gg) * x_mask
            y = self.attn_layers[i](x, x, attn_mask)
            x = self.norm_layers_1[i](x + y)
            y = self.ffn_layers[i](x, x_mask)
            x = self.norm_layers_2[i](x + y)
        return x * x_mask


class TextEncoder(nn.Module):
    def __init__(self, hp: HParams):
        super().__init__()
        H = hp.hidden_channels
        self.hidden_channels, self.out_channels = H, hp.inter_channels
        self.emb = nn.Embedding(hp.n_vocab, H)
        nn.init.normal_(self.emb.weight, 0.0, H ** -0.5)
        self.tone_emb = nn.Embedding(hp.num_tones, H)
        nn.init.normal_(self.tone_emb.weight, 0.0, H ** -0.5)
        self.language_emb = nn.Embedding(hp.num_languages, H)
        nn.init.normal_(self.language_emb.weight, 0.0, H ** -0.5)
        self.bert_proj = nn.Conv1d(hp.bert_dim, H, 1)
        self.style_proj = nn.Linear(hp.style_dim, H)
        self.encoder = Encoder(H, hp.filter_channels, hp.n_heads, hp.n_layers, hp.kernel_size,
                               hp.window_size, gin_channels=hp.gin_channels)
        self.proj = nn.Conv1d(H, hp.inter_channels * 2, 1)

    def forward(self, x, x_lengths, tone, language, bert, style_vec, g=None):
        bert_emb = self.bert_proj(bert).transpose(1, 2)
        style_emb = self.style_proj(style_vec.unsqueeze(1))
        x = (self.emb(x) + self.tone_emb(tone) + self.language_emb(language) + bert_emb + style_emb) \
            * math.sqrt(self.hidden_channels)
        x = x.transpose(1, -1)
        x_mask = sequence_mask(x_lengths, x.size(2)).unsqueeze(1).to(x.dtype)
        x = self.encoder(x * x_mask, x_mask, g=g)
        stats = self.proj(x) * x_mask
        m, logs = torch.split(stats, self.out_channels, dim=1)
        return x, m, logs, x_mask


# ----------------------------------------------------------------------------------------------
# duration predictors
# ----------------------------------------------------------------------------------------------

class DurationPredictor(nn.Module):
    def __init__(self, in_channels, filter_channels, kernel_size, gin_channels):
        super().__init__()
        self.conv_1 = nn.Conv1d(in_channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.norm_1 = LayerNorm(filter_channels)
        self.conv_2 = nn.Conv1d(filter_channels, filter_channels, kernel_size, padding=kernel_size // 2)
        self.norm_2 = LayerNorm(filter_channels)
        self.proj = nn.Conv1d(filter_channels, 1, 1)
        self.cond = nn.Conv1d(gin_channels, in_channels, 1)

    def forward(self, x, x_mask, g=None):
        if g is not None:
            x = x + self.cond(g)
        x = self.norm_1(torch.relu(self.conv_1(x * x_mask)))
        x = self.norm_2(torch.relu(self.conv_2(x * x_mask)))
        x = self.proj(x * x_mask)
        return x * x_mask


class DDSConv(nn.Module):
    def __init__(self, channels, kernel_size, n_layers):
        super().__init__()
        self.n_layers = n_layers
        self.convs_sep = nn.ModuleList()
        self.convs_1x1 = nn.ModuleList()
        self.norms_1 = nn.ModuleList()
        self.norms_2 = nn.ModuleList()
        for i in range(n_layers):
            dilation = kernel_size ** i
            padding = (kernel_size * dilation - dilation) // 2
            self.convs_sep.append(nn.Conv1d(channels, channels, kernel_size, groups=channels,
                                            dilation=dilation, padding=padding))
            self.convs_1x1.append(nn.Conv1d(channels, channels, 1))
            self.norms_1.append(LayerNorm(channels))
            self.norms_2.append(LayerNorm(channels))

    def forward(self, x, x_mask, g=None):
        if g is not None:
            x = x + g
        for i in range(self.n_layers):
            y = self.convs_sep[i](x * x_mask)
            y = F.gelu(self.norms_1[i](y))
            y = self.convs_1x1[i](y)
            y = F.gelu(self.norms_2[i](y))
            x = x + y
        return x * x_mask


DEFAULT_MIN_BIN_WIDTH = 1e-3
DEFAULT_MIN_BIN_HEIGHT = 1e-3
DEFAULT_MIN_DERIVATIVE = 1e-3


def _searchsorted(bin_locations, inputs, eps=1e-6):
    bin_locations = bin_locations.clone()
    bin_locations[..., -1] += eps
    return torch.sum(inputs[..., None] >= bin_locations, dim=-1) - 1


def rational_quadratic_spline_inverse(inputs, uw, uh, ud, left, right, bottom, top):
    """Inverse branch of upstream transforms.rational_quadratic_spline (inside-interval part)."""
    num_bins = uw.shape[-1]
    widths = F.softmax(uw, dim=-1)
    widths = DEFAULT_MIN_BIN_WIDTH + (1 - DEFAULT_MIN_BIN_WIDTH * num_bins) * widths
    cumwidths = F.pad(torch.cumsum(widths, dim=-1), pad=(1, 0), mode="constant", value=0.0)
    cumwidths = (right - left) * cumwidths + left
    cumwidths[..., 0] = left
    cumwidths[..., -1] = right
    widths = cumwidths[..., 1:] - cumwidths[..., :-1]

    derivatives = DEFAULT_MIN_DERIVATIVE + F.softplus(ud)

    heights = F.softmax(uh, dim=-1)
    heights = DEFAULT_MIN_BIN_HEIGHT + (1 - DEFAULT_MIN_BIN_HEIGHT * num_bins) * heights
    cumheights = F.pad(torch.cumsum(heights, dim=-1), pad=(1, 0), mode="constant", value=0.0)
    cumheights = (top - bottom) * cumheights + bottom
    cumheights[..., 0] = bottom
    cumheights[..., -1] = top
    heights = cumheights[..., 1:] - cumheights[..., :-1]

    bin_idx = _searchsorted(cumheights, inputs)[..., None]

    input_cumwidths = cumwidths.gather(-1, bin_idx)[..., 0]
    input_bin_widths = widths.gather(-1, bin_idx)[..., 0]
    input_cumheights = cumheights.gather(-1, bin_idx)[..., 0]
    delta = heights / widths
    input_delta = delta.gather(-1, bin_idx)[..., 0]
    input_derivatives = derivatives.gather(-1, bin_idx)[..., 0]
    input_derivatives_plus_one = derivatives[..., 1:].gather(-1, bin_idx)[..., 0]
    input_heights = heights.gather(-1, bin_idx)[..., 0]

    a = (inputs - input_cumheights) * (input_derivatives + input_derivatives_plus_one - 2 * input_delta) \
        + input_heights * (input_delta - input_derivatives)
    b = input_heights * input_derivatives - (inputs - input_cumheights) \
        * (input_derivatives + input_derivatives_plus_one - 2 * input_delta)
    c = -input_delta * (inputs - input_cumheights)
    discriminant = b.pow(2) - 4 * a * c
    assert (discriminant >= 0).all()
    root = (2 * c) / (-b - torch.sqrt(discriminant))
    return root * input_bin_widths + input_cumwidths


def unconstrained_rqs_inverse(inputs, uw, uh, ud, tail_bound):
    """upstream transforms.unconstrained_rational_quadratic_spline(inverse=True, tails='linear')."""
    inside = (inputs >= -tail_bound) & (inputs <= tail_bound)
    outputs = torch.zeros_like(inputs)
    ud = F.pad(ud, pad=(1, 1))
    constant = math.log(math.exp(1 - DEFAULT_MIN_DERIVATIVE) - 1)
    ud[..., 0] = constant
    ud[..., -1] = constant
    outputs[~inside] = inputs[~inside]
    if inside.any():
        outputs[inside] = rational_quadratic_spline_inverse(
            inputs[inside], uw[inside, :], uh[inside, :], ud[inside, :],
            left=-tail_bound, right=tail_bound, bottom=-tail_bound, top=tail_bound)
    return outputs


class ConvFlow(nn.Module):
    def __init__(self, in_channels, filter_channels, kernel_size, n_layers, num_bins=10, tail_bound=5.0):
        super().__init__()
        self.filter_channels, self.num_bins, self.tail_bound = filter_channels, num_bins, tail_bound
        self.half_channels = in_channels // 2
        self.pre = nn.Conv1d(self.half_channels, filter_channels, 1)
        self.convs = DDSConv(filter_channels, kernel_size, n_layers)
        self.proj = nn.Conv1d(filter_channels, self.half_channels * (num_bins * 3 - 1), 1)

    def forward_reverse(self, x, x_mask, g):
        x0, x1 = torch.split(x, [self.half_channels] * 2, 1)
        h = self.pre(x0)
        h = self.convs(h, x_mask, g=g)
        h = self.proj(h) * x_mask
        b, c, t = x0.shape
        h = h.reshape(b, c, -1, t).permute(0, 1, 3, 2)
        uw = h[..., : self.num_bins] / math.sqrt(self.filter_channels)
        uh = h[..., self.num_bins: 2 * self.num_bins] / math.sqrt(self.filter_channels)
        ud = h[..., 2 * self.num_bins:]
        x1 = unconstrained_rqs_inverse(x1, uw, uh, ud, self.tail_bound)
        return torch.cat([x0, x1], 1) * x_mask


class ElementwiseAffine(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.m = nn.Parameter(torch.zeros(channels, 1))
        self.logs = nn.Parameter(torch.zeros(channels, 1))

    def forward_reverse(self, x, x_mask):
        return (x - self.m) * torch.exp(-self.logs) * x_mask


class Flip(nn.Module):
    def forward_reverse(self, x, *a, **k):
        return torch.flip(x, [1])


class StochasticDurationPredictor(nn.Module):
    def __init__(self, in_channels, kernel_size, n_flows, gin_channels, num_bins=10, tail_bound=5.0):
        super().__init__()
        filter_channels = in_channels  # upstream overrides filter_channels with in_channels
        self.flows = nn.ModuleList([ElementwiseAffine(2)])
        for _ in range(n_flows):
            self.flows.append(ConvFlow(2, filter_channels, kernel_size, 3, num_bins, tail_bound))
            self.flows.append(Flip())
        self.pre = nn.Conv1d(in_channels, filter_channels, 1)
        self.proj = nn.Conv1d(filter_channels, filter_channels, 1)
        self.convs = DDSConv(filter_channels, kernel_size, 3)
        self.cond = nn.Conv1d(gin_channels, filter_channels, 1)
        # training-only post_* modules are not part of the inference graph

    def forward_reverse(self, x, x_mask, g, noise):
        """noise: [B,2,T_x] already multiplied by noise_scale_w."""
        x = self.pre(x)
        x = x + self.cond(g)
        x = self.convs(x, x_mask)
        x = self.proj(x) * x_mask
        flows = list(reversed(self.flows))
        flows = flows[:-2] + [flows[-1]]  # upstream: "remove a useless vflow"
        z = noise
        for flow in flows:
            if isinstance(flow, ConvFlow):
                z = flow.forward_reverse(z, x_mask, g=x)
            elif isinstance(flow, ElementwiseAffine):
                z = flow.forward_reverse(z, x_mask)
            else:
                z = flow.forward_reverse(z)
        z0, _ = torch.split(z, [1, 1], 1)
        return z0


# ----------------------------------------------------------------------------------------------
# flows
# ----------------------------------------------------------------------------------------------

class TransformerCouplingLayer(nn.Module):
    def __init__(self, hp: HParams):
        super().__init__()
        self.half_channels = hp.inter_channels // 2
        self.pre = nn.Conv1d(self.half_channels, hp.hidden_channels, 1)
        self.enc = Encoder(hp.hidden_channels, hp.filter_channels, hp.n_heads, hp.n_layers_trans_flow,
                           hp.flow_kernel_size, hp.window_size, gin_channels=hp.gin_channels)
        self.post = nn.Conv1d(hp.hidden_channels, self.half_channels, 1)  # mean_only
        # upstream zero-inits post; the synthetic generator overrides it (SURVEY.md §8d "weights")

    def forward_reverse(self, x, x_mask, g):
        x0, x1 = torch.split(x, [self.half_channels] * 2, 1)
        h = self.pre(x0) * x_mask
        h = self.enc(h, x_mask, g=g)
        m = self.post(h) * x_mask
        x1 = (x1 - m) * x_mask
        return torch.cat([x0, x1], 1)


class WN(nn.Module):
    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels):
        super().__init__()
        self.hidden_channels, self.n_layers = hidden_channels, n_layers
        self.cond_layer = nn.Conv1d(gin_channels, 2 * hidden_channels * n_layers, 1)
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        for i in range(n_layers):
            dilation = dilation_rate ** i
            padding = int((kernel_size * dilation - dilation) / 2)
            self.in_layers.append(nn.Conv1d(hidden_channels, 2 * hidden_channels, kernel_size,
                                            dilation=dilation, padding=padding))
            rs = 2 * hidden_channels if i < n_layers - 1 else hidden_channels
            self.res_skip_layers.append(nn.Conv1d(hidden_channels, rs, 1))

    def forward(self, x, x_mask, g):
        output = torch.zeros_like(x)
        H = self.hidden_channels
        g = self.cond_layer(g)
        for i in range(self.n_layers):
            x_in = self.in_layers[i](x)
            g_l = g[:, i * 2 * H:(i + 1) * 2 * H, :]
            in_act = x_in + g_l
            acts = torch.tanh(in_act[:, :H, :]) * torch.sigmoid(in_act[:, H:, :])
            res_skip = self.res_skip_layers[i](acts)
            if i < self.n_layers - 1:
                x = (x + res_skip[:, :H, :]) * x_mask
                output = output + res_skip[:, H:, :]
            else:
                output = output + res_skip
        return output * x_mask


class ResidualCouplingLayer(nn.Module):
    def __init__(self, hp: HParams):
        super().__init__()
        self.half_channels = hp.inter_channels // 2
        self.pre = nn.Conv1d(self.half_channels, hp.hidden_channels, 1)
        self.enc = WN(hp.hidden_channels, hp.flow_kernel_size, 1, hp.wn_layers, hp.gin_channels)
        self.post = nn.Conv1d(hp.hidden_channels, self.half_channels, 1)

    def forward_reverse(self, x, x_mask, g):
        x0, x1 = torch.split(x, [self.half_channels] * 2, 1)
        h = self.pre(x0) * x_mask
        h = self.enc(h, x_mask, g=g)
        m = self.post(h) * x_mask
        x1 = (x1 - m) * x_mask
        return torch.cat([x0, x1], 1)


class CouplingBlock(nn.Module):
    """TransformerCouplingBlock / ResidualCouplingBlock: flows = [L, Flip] * n_flows."""

    def __init__(self, hp: HParams):
        super().__init__()
        self.flows = nn.ModuleList()
        for _ in range(hp.n_flow_layer):
            self.flows.append(TransformerCouplingLayer(hp) if hp.use_transformer_flow
                              else ResidualCouplingLayer(hp))
            self.flows.append(Flip())

    def forward_reverse(self, x, x_mask, g):
        for flow in reversed(self.flows):
            if isinstance(flow, Flip):
                x = flow.forward_reverse(x)
            else:
                x = flow.forward_reverse(x, x_mask, g)
        return x


# ----------------------------------------------------------------------------------------------
# HiFi-GAN generator
# ----------------------------------------------------------------------------------------------

class ResBlock1(nn.Module):
    def __init__(self, channels, kernel_size, dilation):
        super().__init__()
        self.convs1 = nn.ModuleList([
            nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=(kernel_size * d - d) // 2)
            for d in dilation])
        self.convs2 = nn.ModuleList([
            nn.Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=(kernel_size - 1) // 2)
            for _ in dilation])

    def forward(self, x):
        for c1, c2 in zip(self.convs1, self.convs2):
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = c1(xt)
            xt = F.leaky_relu(xt, LRELU_SLOPE)
            xt = c2(xt)
            x = xt + x
        return x


class Generator(nn.Module):
    def __init__(self, hp: HParams):
        super().__init__()
        self.num_kernels = len(hp.resblock_kernel_sizes)
        self.num_upsamples = len(hp.upsample_rates)
        C0 = hp.upsample_initial_channel
        self.conv_pre = nn.Conv1d(hp.inter_channels, C0, 7, 1, padding=3)
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(hp.upsample_rates, hp.upsample_kernel_sizes)):
            self.ups.append(nn.ConvTranspose1d(C0 // (2 ** i), C0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2))
        self.resblocks = nn.ModuleList()
        ch = C0
        for i in range(len(self.ups)):
            ch = C0 // (2 ** (i + 1))
            for k, d in zip(hp.resblock_kernel_sizes, hp.resblock_dilation_sizes):
                self.resblocks.append(ResBlock1(ch, k, d))
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3, bias=False)
        self.cond = nn.Conv1d(hp.gin_channels, C0, 1)

    def forward(self, x, g=None, collect: Optional[dict] = None):
        x = self.conv_pre(x)
        if g is not None:
            x = x + self.cond(g)
        if collect is not None:
            collect["dec_pre"] = x
        for i in range(self.num_upsamples):
            x = F.leaky_relu(x, LRELU_SLOPE)
            x = self.ups[i](x)
            if collect is not None:
                collect[f"dec_up{i}"] = x
            xs = None
            for j in range(self.num_kernels):
                r = self.resblocks[i * self.num_kernels + j](x)
                xs = r if xs is None else xs + r
            x = xs / self.num_kernels
            if collect is not None:
                collect[f"dec_stage{i}"] = x
        x = F.leaky_relu(x)  # default slope 0.01, as upstream
        x = self.conv_post(x)
        return torch.tanh(x)


# ----------------------------------------------------------------------------------------------
# SynthesizerTrn (JP-Extra), inference only
# ----------------------------------------------------------------------------------------------

class SynthesizerTrn(nn.Module):
    def __init__(self, hp: HParams):
        super().__init__()
        self.hp = hp
        self.enc_p = TextEncoder(hp)
        self.dec = Generator(hp)
        self.flow = CouplingBlock(hp)
        self.sdp = StochasticDurationPredictor(hp.hidden_channels, 3, hp.sdp_n_flows, hp.gin_channels,
                                               hp.sdp_num_bins, hp.sdp_tail_bound)
        self.dp = DurationPredictor(hp.hidden_channels, hp.dp_filter_channels, 3, hp.gin_channels)
        self.emb_g = nn.Embedding(hp.n_speakers, hp.gin_channels)

    @torch.no_grad()
    def infer(self, x, x_lengths, sid, tone, language, bert, style_vec, *, noise_sdp, noise_zp=None,
              noise_scale=0.667, length_scale=1.0, noise_scale_w=0.8, sdp_ratio=0.0,
              return_intermediates=False):
        """Restates upstream ``SynthesizerTrn.infer``; argument meaning as in
        /root/reference/scripts/convert/convert_model.py:97-110.

        noise_sdp: N(0,1) [B,2,T_x]; noise_zp: N(0,1) [B,192,>=T_y] (extra frames ignored) or a
        callable (B, C, T_y) -> tensor.
        """
        inter: Dict[str, torch.Tensor] = {}
        g = self.emb_g(sid).unsqueeze(-1)
        x, m_p, logs_p, x_mask = self.enc_p(x, x_lengths, tone, language, bert, style_vec, g=g)
        inter.update(enc_x=x, m_p=m_p, logs_p=logs_p)
        logw_sdp = self.sdp.forward_reverse(x, x_mask, g, noise_sdp * noise_scale_w)
        logw_dp = self.dp(x, x_mask, g=g)
        logw = logw_sdp * sdp_ratio + logw_dp * (1 - sdp_ratio)
        inter.update(logw_sdp=logw_sdp, logw_dp=logw_dp, logw=logw)
        w = torch.exp(logw) * x_mask * length_scale
        w_ceil = torch.ceil(w)
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        y_mask = sequence_mask(y_lengths, None).unsqueeze(1).to(x_mask.dtype)
        attn_mask = x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)
        attn = generate_path(w_ceil, attn_mask)
        inter.update(w=w, w_ceil=w_ceil, y_lengths=y_lengths, attn=attn)
        m_p = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)
        logs_p = torch.matmul(attn.squeeze(1), logs_p.transpose(1, 2)).transpose(1, 2)
        t_y = m_p.shape[2]
        if callable(noise_zp):
            eps = noise_zp(m_p.shape[0], m_p.shape[1], t_y)
        else:
            eps = noise_zp[:, :, :t_y]
        z_p = m_p + eps * torch.exp(logs_p) * noise_scale
        z = self.flow.forward_reverse(z_p, y_mask, g)
        inter.update(m_p_exp=m_p, logs_p_exp=logs_p, z_p=z_p, z=z)
        o = self.dec(z * y_mask, g=g, collect=inter if return_intermediates else None)
        if return_intermediates:
            return o, inter
        return o


# ----------------------------------------------------------------------------------------------
# synthetic weights (SURVEY.md §8d "weights")
# ----------------------------------------------------------------------------------------------

def init_synthetic_weights(model: SynthesizerTrn, seed: int = 0, target_frames_per_symbol: float = 689 / 241,
                           out_peak_scale: float = 1.0, decoder_init: str = "calibrated") -> None:
    """Random-init, tsukuyomi-shaped weights.

    PyTorch default inits everywhere except: the decoder, whose ups/resblocks/conv_post weights are
    variance-preserving ("calibrated": activations stay O(1) through all five stages and the
    waveform peaks near 0.4, so the 1e-3 max-abs bound is a real ~0.3 % test) or,
    with ``decoder_init="upstream"``, N(0, 0.01) as upstream's ``init_weights``; the upstream zero-initialised ``post`` convs of coupling layers and ConvFlow
    ``proj`` get a small non-zero init so that flow and SDP are not identities; ``dp.proj.bias``
    is set so durations average ``target_frames_per_symbol`` frames.
    """
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.startswith("dec.ups") or name.startswith("dec.resblocks") or name == "dec.conv_post.weight":
                if name.endswith("weight"):
                    if decoder_init == "upstream":          # upstream init_weights: N(0, 0.01)
                        std = 0.01 if not name.startswith("dec.conv_post") else 1.0 / math.sqrt(p[0].numel())
                    elif name.startswith("dec.ups"):        # [C_in, C_out, k], k/u taps hit each output
                        u = model.hp.upsample_rates[int(name.split(".")[2])]
                        std = 1.0 / math.sqrt(p.shape[0] * p.shape[2] / u)
                    elif name.startswith("dec.resblocks"):
                        std = 0.5 / math.sqrt(p.shape[1] * p.shape[2])
                    else:
                        std = 0.6 / math.sqrt(p.shape[1] * p.shape[2])
                    p.copy_(torch.randn(p.shape, generator=gen) * std)
                else:
                    p.copy_(torch.randn(p.shape, generator=gen) * 0.05)
            else:
                # re-draw default-style init from our generator for determinism
                if p.dim() >= 2:
                    fan_in = p[0].numel() if p.dim() > 1 else p.numel()
                    if "emb" in name and "emb_rel" not in name and "spk_emb" not in name and p.dim() == 2:
                        std = model.hp.hidden_channels ** -0.5 if name != "emb_g.weight" else 1.0
                        p.copy_(torch.randn(p.shape, generator=gen) * std)
                    elif "emb_rel" in name:
                        p.copy_(torch.randn(p.shape, generator=gen) * (p.shape[-1] ** -0.5))
                    else:
                        bound = 1.0 / math.sqrt(fan_in)
                        p.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * bound)
                elif name.endswith("gamma"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen))
                elif name.endswith("beta"):
                    p.copy_(0.1 * torch.randn(p.shape, generator=gen))
                else:
                    p.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * 0.05)
        # flow / sdp heads that upstream zero-initialises
        for f in model.flow.flows:
            if hasattr(f, "post"):
                f.post.weight.copy_(torch.randn(f.post.weight.shape, generator=gen) * 0.05)
                f.post.bias.copy_(torch.randn(f.post.bias.shape, generator=gen) * 0.05)
        for f in model.sdp.flows:
            if isinstance(f, ConvFlow):
                f.proj.weight.copy_(torch.randn(f.proj.weight.shape, generator=gen) * 0.3)
                f.proj.bias.copy_(torch.randn(f.proj.bias.shape, generator=gen) * 0.3)
            if isinstance(f, ElementwiseAffine):
                f.m.copy_(torch.randn(f.m.shape, generator=gen) * 0.1)
                f.logs.copy_(torch.randn(f.logs.shape, generator=gen) * 0.1)
        # steer durations: dp output ~ N(bias, small)
        model.dp.proj.weight.mul_(0.25)
        model.dp.proj.bias.fill_(math.log(target_frames_per_symbol) - 0.35)
        if out_peak_scale != 1.0:
            model.dec.conv_post.weight.mul_(out_peak_scale)


def build_model(hp: Optional[HParams] = None, seed: int = 0, **kw) -> SynthesizerTrn:
    hp = hp or HParams()
    torch.manual_seed(seed)
    m = SynthesizerTrn(hp).eval()
    init_synthetic_weights(m, seed=seed, **kw)
    return m


def state_dict_numpy(model: nn.Module) -> Dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def synthetic_inputs(hp: HParams, t_x: int, seed: int, batch: int = 1):
    """Synthetic phoneme/tone/language/BERT/style inputs with the reference's contract (§3.2)."""
    assert t_x % 2 == 1
    g = torch.Generator().manual_seed(seed)
    n_ph = (t_x - 1) // 2
    x = torch.zeros(batch, t_x, dtype=torch.long)
    tone = torch.zeros(batch, t_x, dtype=torch.long)
    lang = torch.zeros(batch, t_x, dtype=torch.long)
    x[:, 1::2] = torch.randint(1, hp.n_vocab, (batch, n_ph), generator=g)
    tone[:, 1::2] = torch.randint(0, 2, (batch, n_ph), generator=g) + 6
    lang[:, 1::2] = 1
    bert = torch.randn(batch, hp.bert_dim, t_x, generator=g)
    style = torch.randn(batch, hp.style_dim, generator=g) * 0.1
    return x, tone, lang, bert, style
