"""Batch-1 latency anatomy of the ~5 s utterance (the `latency` block of bench.py): wall-time percentiles, the library's
region split, and — when run under `ncu --metrics gpu__time_duration.sum` — the launch list of ONE call.
Usage: python tools/profile_latency.py [calls]    (calls = 0: one warm call + one profiled call, for ncu)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
import bench  # noqa: E402
from oracle import vits as ov  # noqa: E402
import sbv2_b200 as S  # noqa: E402
from sbv2_b200 import assets  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
hp = ov.HParams()
oracle = ov.build_model(hp, seed=0)
model = S.Model(assets.synth_onnx(ov.state_dict_numpy(oracle), hp.upsample_rates, hp.resblock_dilation_sizes), bert=False)
u = bench.one_utt(hp, 151, 77)
u = dict(u, bert=S.pinned_copy(u["bert"]))
for _ in range(3 if calls else 1):
    a = model.synthesize_batch([u])
l0 = model.launch_count
model.synthesize_batch([u])
print("launches per call", model.launch_count - l0, "audio_s", a[0].size / 44100)
if calls:
    ts = []
    for _ in range(calls):
        t = time.perf_counter()
        model.synthesize_batch([u])
        ts.append(time.perf_counter() - t)
    ts = np.array(ts) * 1e3
    print(f"wall ms: p50 {np.median(ts):.3f} p90 {np.percentile(ts, 90):.3f} p99 {np.percentile(ts, 99):.3f} min {ts.min():.3f}")
    model.enable_timing(True)
    reg = {"text": [], "flow": [], "decoder": []}
    for _ in range(20):
        model.synthesize_batch([u])
        for k in reg:
            reg[k].append(model.region_ms(k))
    print("region ms (CUDA events, median):", {k: round(float(np.median(v)), 3) for k, v in reg.items()})
