"""DeBERTa-v2-large shape (cfg2: 32 x 128 tokens) forward passes for ncu. Usage: python tools/profile_bert.py [exact|fp16] [passes] [batch] [seq]"""
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

mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32
seq = int(sys.argv[4]) if len(sys.argv) > 4 else 128
if mode == "fp16":
    os.environ["SBV2_B200_BERT"] = "fp16"
cfg = od.deberta_config()
bert = S.Model(assets.deberta_onnx(od.state_dict_numpy(od.build_model(cfg, seed=1))), bert=True)
ids = torch.randint(3, cfg.vocab_size, (batch, seq), generator=torch.Generator().manual_seed(21)).numpy()
mask = np.ones_like(ids)
bert.predict_batch(ids, mask)
l0 = bert.launch_count
bert.enable_timing(True)
for _ in range(passes):
    bert.predict_batch(ids, mask)
print(mode, "launches per pass", (bert.launch_count - l0) // passes, "kernel ms", bert.region_ms("bert"))
