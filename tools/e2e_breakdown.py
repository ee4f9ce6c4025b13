"""Host-side breakdown of one end-to-end synthesize_batch call (upload / run / download)."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
import numpy as np
import bench
from oracle import vits as ov
import sbv2_b200 as S
from sbv2_b200 import assets

hp = ov.HParams()
oracle = ov.build_model(hp, seed=0)
model = S.Model(assets.synth_onnx(ov.state_dict_numpy(oracle), hp.upsample_rates, hp.resblock_dilation_sizes), bert=False)
utts, _ = bench.make_batch(hp, 32, seed=100)
for it in range(6):
    t0 = time.perf_counter()
    db = S.DeviceBatch(model, utts)
    t1 = time.perf_counter()
    n = db.run()
    t2 = time.perf_counter()
    out = db.download()
    t3 = time.perf_counter()
    db.close()
    t4 = time.perf_counter()
    o2 = model.synthesize_batch(utts)
    t5 = time.perf_counter()
    print(f"iter {it}: upload {1e3*(t1-t0):6.1f} ms  run {1e3*(t2-t1):6.1f}  download {1e3*(t3-t2):6.1f}  close {1e3*(t4-t3):5.1f} | synthesize_batch {1e3*(t5-t4):6.1f} ms  ({n/44100:.0f} s audio)")
