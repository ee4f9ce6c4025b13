"""Batch vs single-utterance runs, bit for bit (the row-count-dependent kernel variants must not change any utterance).
Usage: python tools/variant_bitcheck.py   (env: SBV2_B200_TEXT_ATTN_RQ=1|4, SBV2_B200_LN_ROWS=1|4 force a variant)"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
import sbv2_b200 as S  # noqa: E402
from oracle import vits as ov  # noqa: E402


def run_gpu(model, u):
    return model.synthesize_with_noise(u["bert"][0].numpy(), u["x"][0].numpy(), u["sid"], u["tone"][0].numpy(),
                                       u["lang"][0].numpy(), u["style"][0].numpy(), u["sdp_ratio"], u["length_scale"],
                                       u["noise_scale"], u["noise_scale_w"], u["noise_sdp"][0].numpy(), u["noise_zp"][0].numpy())


hp = ov.HParams()
oracle, onnx = util.synth_assets(hp, seed=0)
model = S.Model(onnx, bert=False)
us = [util.make_utterance(hp, t, seed=200 + i, sdp_ratio=r, length_scale=ls)
      for i, (t, r, ls) in enumerate([(23, 0.0, 1.0), (151, 0.4, 1.0), (57, 0.0, 1.3), (5, 1.0, 0.8), (241, 0.2, 1.0)])]
singles = [run_gpu(model, u) for u in us]
again = [run_gpu(model, u) for u in us]
audios, durs, f2ps = model.synthesize_batch([util.to_api(u) for u in us], want_alignment=True)
env = {k: v for k, v in os.environ.items() if k.startswith("SBV2_B200_")}
for i in range(len(us)):
    same_run = np.array_equal(singles[i][0], again[i][0])
    d = np.abs(audios[i] - singles[i][0]).max() if audios[i].shape == singles[i][0].shape else -1
    print(env, "utt", i, "T_y", len(f2ps[i]), "single==single", same_run, "dur==", np.array_equal(durs[i], singles[i][1]),
          "batch-single max-abs", d)
