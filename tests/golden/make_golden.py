"""Mints the golden fixtures from the CPU oracle (oracle/vits.py, oracle/deberta.py).

The reference has no golden vectors; these pin the oracle so later sessions notice if
it drifts.  Run from the repo root: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import util  # noqa: E402
from util import ov  # noqa: E402


def synth_fixture(name, hp, t_x, weights_seed, input_seed, sdp_ratio):
    model = ov.build_model(hp, seed=weights_seed)
    u = util.make_utterance(hp, t_x, seed=input_seed, sdp_ratio=sdp_ratio)
    o, inter = util.oracle_run(model, u)
    attn = inter["attn"][0, 0].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), weights_seed=weights_seed, input_seed=input_seed, t_x=t_x,
                        sdp_ratio=sdp_ratio, durations=inter["w_ceil"][0, 0].numpy().astype(np.int32),
                        frame2ph=attn.argmax(1).astype(np.int32), logw=inter["logw"][0, 0].numpy(),
                        w=inter["w"][0, 0].numpy(), audio=o[0, 0].numpy().astype(np.float32))
    print(name, "T_y", attn.shape[0], "peak", float(o.abs().max()))


def bert_fixture(name, seed, s):
    from oracle import deberta as od
    cfg = od.tiny_config()
    m = od.build_model(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 5)
    ids = torch.randint(3, cfg.vocab_size, (1, s), generator=g)
    out = od.predict(m, ids, torch.ones_like(ids))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, ids=ids.numpy(), out=out[0].numpy().astype(np.float32))
    print(name, out.shape, float(out.abs().max()))


if __name__ == "__main__":
    synth_fixture("synth_tiny_tx23", ov.tiny_hparams(), 23, 0, 12, 0.4)
    bert_fixture("deberta_tiny_s7", 1, 7)
