"""CPU oracle for the BERT-feature graph (deberta.onnx).

TEST INFRASTRUCTURE ONLY — see oracle/vits.py for the rules.

The reference exports ``AutoModelForMaskedLM("ku-nlp/deberta-v2-large-japanese-char-wwm")`` run with
``output_hidden_states=True`` and keeps ``hidden_states[-3]`` of batch element 0
(/root/reference/scripts/convert/convert_deberta.py:27-35); Rust feeds ``input_ids`` and
``attention_mask`` of shape [1, T_tok] (/root/reference/crates/sbv2_core/src/bert.rs:11-16).
HuggingFace ``transformers`` (5.5.0 in this image; unpinned in the reference) is importable, so the
oracle IS the code the reference exported: ``DebertaV2Model`` with the checkpoint's configuration
(restated from SURVEY.md §B because the checkpoint's config.json is not on disk) and random weights.
PARITY UNPINNED against ONNX Runtime (not installable offline); pinned against HF itself.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
from transformers import DebertaV2Config, DebertaV2Model


def deberta_config(hidden_size: int = 1024, num_hidden_layers: int = 24, num_attention_heads: int = 16,
                   intermediate_size: int = 4096, vocab_size: int = 22012) -> DebertaV2Config:
    return DebertaV2Config(
        vocab_size=vocab_size, hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
        num_attention_heads=num_attention_heads, intermediate_size=intermediate_size, hidden_act="gelu",
        layer_norm_eps=1e-7, relative_attention=True, position_buckets=256, max_position_embeddings=512,
        pos_att_type=["p2c", "c2p"], share_att_key=True, norm_rel_ebd="layer_norm", position_biased_input=False,
        conv_kernel_size=3, conv_act="gelu", type_vocab_size=0, hidden_dropout_prob=0.0,
        attention_probs_dropout_prob=0.0)


def tiny_config() -> DebertaV2Config:
    """Same structure, 4 layers x 2 heads x 64 (hidden 128): fast CPU/GPU parity tests."""
    return deberta_config(hidden_size=128, num_hidden_layers=4, num_attention_heads=2, intermediate_size=512, vocab_size=500)


def build_model(cfg: DebertaV2Config, seed: int = 1) -> DebertaV2Model:
    torch.manual_seed(seed)
    m = DebertaV2Model(cfg).eval()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("LayerNorm.weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("LayerNorm.bias") or name.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                std = 0.02 if "embeddings" in name else 1.0 / np.sqrt(fan_in)
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return m


def state_dict_numpy(model: DebertaV2Model) -> Dict[str, np.ndarray]:
    """Initializer names as an exported MaskedLM graph has them: prefix ``deberta.``."""
    return {"deberta." + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


@torch.no_grad()
def predict(model: DebertaV2Model, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """[B, S] ids/mask -> hidden_states[-3] [B, S, H] (the reference keeps batch element 0)."""
    out = model(input_ids=input_ids, attention_mask=attention_mask, output_hidden_states=True)
    return out.hidden_states[-3]
