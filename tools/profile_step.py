"""One warm step + N profiled steps of the bench workload (for ncu). Usage: python tools/profile_step.py [steps] [batch] [transformer|wn]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
import bench  # noqa: E402
from oracle import vits as ov  # noqa: E402
import sbv2_b200 as S  # noqa: E402
from sbv2_b200 import assets  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
flow = sys.argv[3] if len(sys.argv) > 3 else "transformer"
hp = ov.HParams() if flow == "transformer" else ov.HParams(use_transformer_flow=False)
oracle = ov.build_model(hp, seed=0)
model = S.Model(assets.synth_onnx(ov.state_dict_numpy(oracle), hp.upsample_rates, hp.resblock_dilation_sizes), bert=False)
utts, _ = bench.make_batch(hp, batch, seed=100)
db = S.DeviceBatch(model, utts)
l0 = model.launch_count
db.run()
print("launches per step", model.launch_count - l0)
for _ in range(steps):
    db.run()
