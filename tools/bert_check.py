"""GPU diagnostic for the DeBERTa path: tiny config vs HF, then the large config timing."""
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

cfg = od.tiny_config()
hf = od.build_model(cfg, seed=1)
m = S.Model(assets.deberta_onnx(od.state_dict_numpy(hf)), bert=True)
print(m.describe())
for s in (7, 64, 130, 300):
    ids = torch.randint(3, cfg.vocab_size, (1, s), generator=torch.Generator().manual_seed(s))
    ref = od.predict(hf, ids, torch.ones_like(ids))[0].numpy()
    got = m.predict(ids[0].numpy(), np.ones(s, np.int64))
    print(f"S={s}: max|ref| {np.abs(ref).max():.3f} max-abs err {np.abs(got - ref).max():.3e} rel-fro {np.linalg.norm(got - ref) / np.linalg.norm(ref):.3e}")
if "--large" in sys.argv:
    cfg = od.deberta_config()
    t = time.time()
    hf = od.build_model(cfg, seed=1)
    onnx = assets.deberta_onnx(od.state_dict_numpy(hf))
    print("built large model %.1fs, %.0f MB" % (time.time() - t, len(onnx) / 1e6))
    t = time.time()
    m = S.Model(onnx, bert=True)
    print("load %.1fs" % (time.time() - t), m.describe())
    ids = torch.randint(3, cfg.vocab_size, (32, 128), generator=torch.Generator().manual_seed(21))
    ref = od.predict(hf, ids[:2], torch.ones_like(ids[:2])).numpy()
    got = m.predict_batch(ids.numpy(), np.ones((32, 128), np.int64))
    print(f"large: max|ref| {np.abs(ref).max():.3f} max-abs err {np.abs(got[:2] - ref).max():.3e} rel-fro {np.linalg.norm(got[:2] - ref) / np.linalg.norm(ref):.3e}")
    for _ in range(3):
        t = time.time()
        m.predict_batch(ids.numpy(), np.ones((32, 128), np.int64))
        print("predict_batch 32x128: %.1f ms (host wall, incl. copies)" % ((time.time() - t) * 1e3))
