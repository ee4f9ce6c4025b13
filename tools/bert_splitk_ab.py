"""DeBERTa-v2-large shape, calls with few tokens: latency and error against the HF fp32 oracle, for the current
SBV2_B200_SPLITK / SBV2_B200_BERT settings (one process per setting: the switches are read once).
Usage: python tools/bert_splitk_ab.py [out.npz]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
from oracle import deberta as od  # noqa: E402
import sbv2_b200 as S  # noqa: E402
from sbv2_b200 import assets  # noqa: E402

cfg = od.deberta_config()
hf = od.build_model(cfg, seed=1)
bert = S.Model(assets.deberta_onnx(od.state_dict_numpy(hf)), bert=True)
tag = f"splitk={os.environ.get('SBV2_B200_SPLITK', '1')} mode={bert.describe()['numerics']}"
outs = {}
for T in (7, 40, 128, 300):
    ids = torch.randint(3, cfg.vocab_size, (1, T), generator=torch.Generator().manual_seed(40 + T))
    mask = np.ones(T, np.int64)
    got = bert.predict(ids[0].numpy(), mask)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        bert.predict(ids[0].numpy(), mask)
        ts.append((time.perf_counter() - t0) * 1e3)
    ref = od.predict(hf, ids, torch.ones_like(ids))[0].numpy()
    err = np.abs(got - ref).max()
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"{tag} T={T}: p50 {np.median(ts):.3f} ms, max-abs {err:.3e}, rel-Frobenius {rel:.3e}", flush=True)
    outs[f"t{T}"] = got
# three short sentences as one batch (three row tiles)
ids = torch.randint(3, cfg.vocab_size, (3, 24), generator=torch.Generator().manual_seed(77)).numpy()
mask = np.ones_like(ids)
mask[1, 9:] = 0
mask[2, 17:] = 0
got = bert.predict_batch(ids, mask)
ts = []
for _ in range(30):
    t0 = time.perf_counter()
    bert.predict_batch(ids, mask)
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"{tag} batch 3 x <=24: p50 {np.median(ts):.3f} ms", flush=True)
outs["b3"] = got
if len(sys.argv) > 1:
    np.savez(sys.argv[1], **outs)
