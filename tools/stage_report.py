"""GPU diagnostic: runs one utterance through the CUDA path and the oracle and prints the error of
every exposed intermediate (debug views of the C ABI).  Usage: python tools/stage_report.py [t_x] [sdp_ratio]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import util  # noqa: E402
from util import ov  # noqa: E402
import sbv2_b200 as S  # noqa: E402


def rep(name, got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    if got.shape != ref.shape:
        print(f"  {name:12s} SHAPE MISMATCH got {got.shape} ref {ref.shape}")
        return
    err = np.abs(got - ref).max() if got.size else 0.0
    print(f"  {name:12s} shape {str(got.shape):18s} max|ref| {np.abs(ref).max():9.4g}  max-abs err {err:9.3g}  rel {err / (np.abs(ref).max() + 1e-30):9.3g}")


def main():
    t_x = int(sys.argv[1]) if len(sys.argv) > 1 else 23
    ratio = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    full = os.environ.get("FULL", "1") == "1"
    hp = ov.HParams() if full else ov.tiny_hparams()
    model, onnx = util.synth_assets(hp, seed=0)
    t = time.time()
    m = S.Model(onnx, bert=False)
    print("load %.2fs" % (time.time() - t), m.describe())
    u = util.make_utterance(hp, t_x, seed=41, sdp_ratio=ratio)
    o, inter = util.oracle_run(model, u)
    audio, dur, f2p = m.synthesize_with_noise(**{k: v for k, v in util.to_api(u).items() if k not in ("sid",)}, sid=u["sid"]) \
        if False else m.synthesize_with_noise(u["bert"][0].numpy(), u["x"][0].numpy(), u["sid"], u["tone"][0].numpy(),
                                              u["lang"][0].numpy(), u["style"][0].numpy(), u["sdp_ratio"], u["length_scale"],
                                              u["noise_scale"], u["noise_scale_w"], u["noise_sdp"][0].numpy(), u["noise_zp"][0].numpy())
    print("launches", m.launch_count)
    tr = lambda t_: t_[0].transpose(0, 1).numpy()
    for name, ref in (("enc_x", tr(inter["enc_x"])), ("stats", np.concatenate([tr(inter["m_p"]), tr(inter["logs_p"])], 1)),
                      ("logw_dp", tr(inter["logw_dp"])), ("z_sdp", None), ("w", tr(inter["w"])), ("z_p", tr(inter["z_p"])),
                      ("z", tr(inter["z"])), ("dec_pre", tr(inter["dec_pre"])), ("dec_last", tr(inter["dec_stage4"]))):
        try:
            got = m.debug_fetch(name)
        except S.Sbv2Error as e:
            print("  ", name, "unavailable:", e.message)
            continue
        if name == "z_sdp":
            rep("logw_sdp", got[:, :1], tr(inter["logw_sdp"]))
        else:
            rep(name, got, ref)
    wc = inter["w_ceil"][0, 0].numpy().astype(np.int32)
    print("  durations equal:", np.array_equal(dur, wc), " T_y", len(f2p), "oracle", int(inter["y_lengths"][0]))
    attn = inter["attn"][0, 0].numpy()  # [T_y, T_x]
    ref_f2p = attn.argmax(1).astype(np.int32)
    print("  frame2ph equal:", np.array_equal(f2p, ref_f2p) if len(f2p) == len(ref_f2p) else "length differs")
    rep("audio", audio, o[0, 0].numpy())
    print("  audio peak", float(np.abs(o.numpy()).max()))


if __name__ == "__main__":
    main()
