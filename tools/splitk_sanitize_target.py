"""Small target for compute-sanitizer: a 3-layer DeBERTa of the large width (hidden 1024), one 9-token sentence in both
numerics modes -> the split-K cluster variants of the conv kernel (DSMEM exchange) and the multi-tile attention (200 tokens)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
from oracle import deberta as od  # noqa: E402
import sbv2_b200 as S  # noqa: E402
from sbv2_b200 import assets  # noqa: E402

cfg = od.deberta_config(num_hidden_layers=3, vocab_size=300)
hf = od.build_model(cfg, seed=1)
onnx = assets.deberta_onnx(od.state_dict_numpy(hf))
for mode in ("exact", "fp16"):
    if mode == "fp16":
        os.environ["SBV2_B200_BERT"] = "fp16"
    bert = S.Model(onnx, bert=True)
    for T in (9, 200):
        ids = torch.randint(3, cfg.vocab_size, (1, T), generator=torch.Generator().manual_seed(T))
        got = bert.predict(ids[0].numpy(), np.ones(T, np.int64))
        ref = od.predict(hf, ids, torch.ones_like(ids))[0].numpy()
        print(mode, T, "rel-Frobenius", np.linalg.norm(got - ref) / np.linalg.norm(ref), flush=True)
