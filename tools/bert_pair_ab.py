"""A/B of the DeBERTa GEMM packings: N block 256 (one CTA per tile) vs 128 with and without CTA pairs (cta_group::2)."""
import sys, os, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "sbv2-api_b200")
from oracle import deberta as od
import sbv2_b200 as S
from sbv2_b200 import assets
cfg = od.deberta_config()
onnx = assets.deberta_onnx(od.state_dict_numpy(od.build_model(cfg, seed=1)))
ref = None
# SBV2_B200_PAIR2 is read once per process: run one process per setting (argv: nb pair), e.g.
#   for c in "256 2" "128 2" "128 0"; do python tools/bert_pair_ab.py $c; done        (pair: 2 = off, 0 = DeBERTa asks for pairs)
nb, pair = int(sys.argv[1]), int(sys.argv[2])
os.environ["SBV2_B200_PAIR2"] = str(pair)
for mode in ("fp16", "exact"):
    for _once in (0,):
        os.environ["SBV2_B200_BERT"] = mode
        os.environ["SBV2_B200_BERT_NB"] = str(nb)
        m = S.Model(onnx, bert=True)
        m.enable_timing(True)
        ids = np.random.default_rng(0).integers(3, cfg.vocab_size, (32, 128)); mask = np.ones_like(ids)
        out = None
        for _ in range(3): out = m.predict_batch(ids, mask)
        ms = m.region_ms("bert")
        if ref is None or ref[0] != mode: ref = (mode, out)
        ids1 = np.arange(3, 10, dtype=np.int64)
        ts = []
        for _ in range(15):
            t = time.perf_counter(); m.predict(ids1, np.ones(7, np.int64)); ts.append(time.perf_counter() - t)
        print(mode, "nb", nb, "pair", "on" if pair == 0 else "off", "32x128 kernel ms %.3f" % ms, "max|diff vs first| %.3g" % float(np.abs(out - ref[1]).max()), "1x7 p50 ms %.3f" % (np.median(ts) * 1e3), flush=True)
        del m
