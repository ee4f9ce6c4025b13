"""Secondary measurements for the other BASELINE.json configurations (bench.py measures the headline
cfg 4): cfg1 short utterance latency, cfg2 DeBERTa 32x128, cfg3 HiFi-GAN decoder alone 32 x 10 s,
cfg5 one 60 s utterance, p50 latency of a ~5 s utterance.  Prints one JSON object per config."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
from oracle import vits as ov  # noqa: E402
from oracle import deberta as od  # noqa: E402
import sbv2_b200 as S  # noqa: E402
from sbv2_b200 import assets  # noqa: E402

SR = 44100
which = set(sys.argv[1:]) or {"cfg1", "cfg2", "cfg3", "cfg5", "p50"}
hp = ov.HParams()


def utt(t_x, seed):
    x, tone, lang, bert, style = ov.synthetic_inputs(hp, t_x, seed)
    return dict(bert=bert[0].numpy(), x_tst=x[0].numpy(), tones=tone[0].numpy(), lang_ids=lang[0].numpy(), style_vec=style[0].numpy(),
                sid=0, sdp_ratio=0.0, length_scale=1.0, noise_scale=0.677, noise_scale_w=0.8)


def wall(fn, n, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return np.array(ts)


if which & {"cfg1", "cfg3", "cfg5", "p50"}:
    oracle = ov.build_model(hp, seed=0)
    model = S.Model(assets.synth_onnx(ov.state_dict_numpy(oracle), hp.upsample_rates, hp.resblock_dilation_sizes), bert=False)

if "cfg1" in which:
    u = utt(23, 12)
    ts = wall(lambda: model.synthesize_batch([u]), 50)
    n = model.synthesize_batch([u])[0].size
    print(json.dumps({"config": "cfg1 short utterance T_x=23 batch 1", "audio_s": n / SR, "latency_ms_p50": float(np.median(ts) * 1e3),
                      "latency_ms_p90": float(np.percentile(ts, 90) * 1e3), "audio_s_per_s": n / SR / float(np.median(ts))}))

if "p50" in which:
    u = utt(151, 77)  # ~5 s
    ts = wall(lambda: model.synthesize_batch([u]), 100)
    n = model.synthesize_batch([u])[0].size
    print(json.dumps({"config": "p50 latency, one ~5 s utterance, batch 1, host in -> pinned host out", "audio_s": n / SR,
                      "latency_ms_p50": float(np.median(ts) * 1e3), "latency_ms_p99": float(np.percentile(ts, 99) * 1e3),
                      "audio_s_per_s": n / SR / float(np.median(ts))}))

if "cfg5" in which:
    u = utt(1801, 51)
    ts = wall(lambda: model.synthesize_batch([u]), 10)
    n = model.synthesize_batch([u])[0].size
    print(json.dumps({"config": "cfg5 one long-form utterance T_x=1801, batch 1", "audio_s": n / SR, "latency_ms_p50": float(np.median(ts) * 1e3),
                      "audio_s_per_s": n / SR / float(np.median(ts))}))

if "cfg3" in which:
    g = torch.Generator().manual_seed(31)
    zs = [torch.randn(192, 861, generator=g).numpy() for _ in range(32)]
    model.enable_timing(True)
    ts = wall(lambda: model.decode_batch(zs), 5)
    dec_ms = model.region_ms("decoder")
    audio_s = 32 * 861 * 512 / SR
    flop = 651.6e6 * 32 * 861
    print(json.dumps({"config": "cfg3 HiFi-GAN decoder alone, 32 x 861 frames (10 s)", "audio_s": audio_s, "decoder_ms": dec_ms,
                      "decoder_tflops": flop / dec_ms / 1e9, "audio_s_per_s_kernel": audio_s / (dec_ms * 1e-3),
                      "audio_s_per_s_e2e": audio_s / float(np.median(ts))}))

if "cfg2" in which:
    cfg = od.deberta_config()
    t = time.time()
    hf = od.build_model(cfg, seed=1)
    onnx = assets.deberta_onnx(od.state_dict_numpy(hf))
    bert = S.Model(onnx, bert=True)
    del onnx
    ids = torch.randint(3, cfg.vocab_size, (32, 128), generator=torch.Generator().manual_seed(21)).numpy()
    mask = np.ones_like(ids)
    got = bert.predict_batch(ids, mask)
    ref = od.predict(hf, torch.from_numpy(ids[:1]), torch.ones(1, 128, dtype=torch.long))[0].numpy()
    err = float(np.abs(got[0] - ref).max())
    rel = float(np.linalg.norm(got[0] - ref) / np.linalg.norm(ref))
    ts = wall(lambda: bert.predict_batch(ids, mask), 10)
    t1 = wall(lambda: bert.predict(ids[0, :7], np.ones(7, np.int64)), 30)
    ms = float(np.median(ts) * 1e3)
    print(json.dumps({"config": "cfg2 DeBERTa-v2-large shape, batch 32 x seq 128 (22 live layers)", "ms_per_batch_e2e": ms,
                      "tokens_per_s": 4096 / (ms * 1e-3), "tflops_e2e": 0.605e9 * 4096 / (ms * 1e-3) / 1e12, "max_abs_err_vs_hf": err,
                      "rel_fro_err_vs_hf": rel, "latency_ms_T_tok_7": float(np.median(t1) * 1e3)}))
