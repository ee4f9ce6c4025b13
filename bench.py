#!/usr/bin/env python
"""Benchmark of the synthesis hot path (BASELINE.json metric: audio-sec/sec, JP-Extra 44.1 kHz,
batch 32 per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the full JP-Extra synthesizer (enc_p, duration predictors, length
regulator, flow, HiFi-GAN decoder) over one batch of 32 synthetic ~8 s utterances (BASELINE config 4
per-GPU shard: T_x odd in U{201..281}, BERT features given, sdp_ratio 0, device-side noise).
Weak scaling: every rank owns a model replica and its own batch, no collective on the data path.

  value : audio-seconds per second with the step's inputs already resident in HBM, timed with CUDA
          events on the model's stream, max over ranks.
  e2e   : the same metric through the C-ABI call a user makes (`sbv2_synthesize_batch`): host
          buffers in, waveforms in pinned host memory out, copies inside the timed region.
  roofline : HiFi-GAN decoder (tcgen05 implicit-GEMM convs), algorithmic FLOPs per step
          (651.6 MFLOP per latent frame) / measured decoder time.
  cpu_baseline : the CPU oracle (PyTorch restatement of the reference's ONNX graph; ONNX Runtime is
          not installable offline) timed on the host cores on a bounded sample of the same batch.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))

SR = 44100
BATCH = 32
DEC_FLOP_PER_FRAME = 651.6e6  # SURVEY.md §8d / §C.1 (flop-counter verified)


def make_batch(hp, batch: int, seed: int):
    """cfg4 per-GPU shard: T_x odd ~ U{201..281}; synthetic ids / BERT features / style."""
    from oracle import vits as ov
    rng = np.random.default_rng(seed)
    utts, raw = [], []
    for i in range(batch):
        t_x = int(rng.integers(100, 141)) * 2 + 1
        x, tone, lang, bert, style = ov.synthetic_inputs(hp, t_x, seed * 1000 + i)
        raw.append((x, tone, lang, bert, style, t_x))
        utts.append(dict(bert=bert[0].numpy(), x_tst=x[0].numpy(), tones=tone[0].numpy(), lang_ids=lang[0].numpy(),
                         style_vec=style[0].numpy(), sid=0, sdp_ratio=0.0, length_scale=1.0, noise_scale=0.677,
                         noise_scale_w=0.8))
    return utts, raw


class ClockSampler:
    """nvidia-smi in loop mode (-lms 100) for the duration of the timed region, as the profiling recipe does."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        time.sleep(0.15)

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[c.strip() for c in line.split(",")] for line in out.splitlines() if line.strip()]
        num = lambda x: x.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in rows if r and num(r[0])]
        mx = [float(r[1]) for r in rows if len(r) > 1 and num(r[1])]
        pw = [float(r[6]) for r in rows if len(r) > 6 and num(r[6])]
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(rows)}


def cpu_oracle_rate(model, raw, n_utts: int, threads: int):
    """audio-sec/sec of the CPU oracle, batch 1 per utterance like the reference (model.rs:66-79)."""
    import torch
    torch.set_num_threads(threads)
    audio_s, t0 = 0.0, None
    g = torch.Generator().manual_seed(7)
    # one short warm-up
    x, tone, lang, bert, style, t_x = raw[0]
    for i in range(n_utts + 1):
        x, tone, lang, bert, style, t_x = raw[i % len(raw)]
        if i == 1:
            t0 = time.perf_counter()
            audio_s = 0.0
        nsdp = torch.randn(1, 2, t_x, generator=g)
        o = model.infer(x, torch.tensor([t_x]), torch.tensor([0]), tone, lang, bert, style, noise_sdp=nsdp,
                        noise_zp=lambda b, c, t: torch.randn(b, c, t, generator=g), noise_scale=0.677, length_scale=1.0,
                        noise_scale_w=0.8, sdp_ratio=0.0)
        audio_s += o.shape[-1] / SR
    dt = time.perf_counter() - t0
    return audio_s / dt, audio_s, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-replicas", type=int, default=2,
                    help="model replicas per GPU driven by separate host threads in the e2e measurement (copies of one overlap compute of the other)")
    ap.add_argument("--tiny", action="store_true", help="reduced model (tests only; not a benchmark configuration)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1

    import torch
    from oracle import vits as ov

    hp = ov.tiny_hparams() if args.tiny else ov.HParams()
    workload = f"cfg4 shard: full JP-Extra pipeline, {args.batch} synthetic ~8 s utterances per GPU per step " \
               f"(T_x odd U{{201..281}}, BERT features given, sdp_ratio 0, transformer flow L=6)"

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return
        model = ov.build_model(hp, seed=0)
        _, raw = make_batch(hp, args.batch, seed=100)
        per_step = 1  # bounded sample: one utterance of the batch per step
        rate, audio_s, dt = cpu_oracle_rate(model, raw, max(1, args.steps * per_step), threads)
        line = {
            "impl": "reference", "metric": "audio-sec/sec", "value": rate, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample": f"{args.steps} utterances of the batch, batch 1 each (as the reference runs)"},
            "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} x 1 utterance (~8 s audio each) of the same synthetic batch; "
                                       "PyTorch-CPU restatement (ONNX Runtime unavailable offline)"},
            "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import sbv2_b200 as S
    from sbv2_b200 import assets
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    oracle = ov.build_model(hp, seed=0)
    onnx = assets.synth_onnx(ov.state_dict_numpy(oracle), hp.upsample_rates, hp.resblock_dilation_sizes)
    model = S.Model(onnx, bert=False, device=local_rank)
    model.seed(1234 + rank)
    utts, raw = make_batch(hp, args.batch, seed=100 + rank)
    stream = torch.cuda.ExternalStream(model.stream, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only: inputs resident in HBM.  R replicas (one CUDA stream + one host thread each) run their
    # own resident batch concurrently, so the mid-pipeline T_y read-back of one overlaps kernels of the other.
    R = max(1, args.e2e_replicas)
    replicas = [model] + [S.Model(onnx, bert=False, device=local_rank) for _ in range(R - 1)]
    del onnx
    for i, m in enumerate(replicas[1:]):
        m.seed(4321 + rank + i)
    dbs = [S.DeviceBatch(m, utts) for m in replicas]
    streams = [torch.cuda.ExternalStream(m.stream, device=local_rank) for m in replicas]
    samples = 0
    for db in dbs:
        for _ in range(args.warmup):
            samples = db.run()
    counts = [args.steps // R + (1 if i < args.steps % R else 0) for i in range(R)]
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sum(m.launch_count for m in replicas)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev_end = [torch.cuda.Event(enable_timing=True) for _ in range(R)]
    go = threading.Barrier(R)

    def kworker(i):
        go.wait()
        for _ in range(counts[i]):
            dbs[i].run()
        ev_end[i].record(streams[i])

    for st in streams:
        st.synchronize()
    ev0.record(streams[0])
    for st in streams[1:]:
        st.wait_event(ev0)  # nothing of any replica starts before ev0
    kthreads = [threading.Thread(target=kworker, args=(i,)) for i in range(R)]
    for th in kthreads:
        th.start()
    for th in kthreads:
        th.join()
    barrier()
    launches = sum(m.launch_count for m in replicas) - l0
    ms = max(ev0.elapsed_time(e) for e in ev_end)
    clocks = sampler.stop()
    db = dbs[0]
    # region split (CUDA events inside the library), measured on extra steps outside the timed region
    model.enable_timing(True)
    dec_ms = flow_ms = text_ms = 0.0
    n_reg = min(3, args.steps)
    for _ in range(n_reg):
        db.run()
        dec_ms += model.region_ms("decoder") / n_reg
        flow_ms += model.region_ms("flow") / n_reg
        text_ms += model.region_ms("text") / n_reg
    model.enable_timing(False)
    frames = samples // hp.hop
    audio_s_step = samples / SR

    # ---- end to end through the C ABI with host buffers
    # (a) one replica, one host thread: the latency-oriented number; (b) R replicas on this GPU, one host
    # thread each (ctypes releases the GIL): H2D / D2H of one replica overlap the kernels of the other.
    for _ in range(2):
        model.synthesize_batch(utts)
    barrier()
    t0 = time.perf_counter()
    single_audio = 0.0
    n_single = max(2, args.steps // 2)
    for _ in range(n_single):
        out = model.synthesize_batch(utts)
        single_audio += sum(a.size for a in out) / SR
    torch.cuda.synchronize()
    e2e_single = single_audio / (time.perf_counter() - t0)
    for m in replicas[1:]:
        for _ in range(2):
            m.synthesize_batch(utts)
    per_thread = [0.0] * R
    counts = [args.steps // R + (1 if i < args.steps % R else 0) for i in range(R)]

    def worker(i):
        for _ in range(counts[i]):
            o = replicas[i].synthesize_batch(utts)
            per_thread[i] += sum(a.size for a in o) / SR

    barrier()
    t0 = time.perf_counter()
    ethreads = [threading.Thread(target=worker, args=(i,)) for i in range(R)]
    for th in ethreads:
        th.start()
    for th in ethreads:
        th.join()
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    e2e_audio = sum(per_thread)
    barrier()
    h2d = sum(u["bert"].nbytes + 3 * 4 * u["x_tst"].size + u["style_vec"].nbytes + 8 * 4 for u in utts)
    d2h = int(samples * 4)

    t = torch.tensor([ms, e2e_dt * 1000.0], dtype=torch.float64, device="cuda")
    tot = torch.tensor([audio_s_step * args.steps, e2e_audio, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    audio_total, e2e_audio_total, launches_total = float(tot[0]), float(tot[1]), int(tot[2])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained")
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained; fp16 has the same tensor rate)"
        if not peak_tf:
            peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
        dec_flop = DEC_FLOP_PER_FRAME * frames
        achieved = dec_flop / (dec_ms * 1e-3) / 1e12 if dec_ms > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "decoder_traffic.json"))).get("dram_bytes_per_step")
        except Exception:
            pass
        line = {
            "metric": "audio-sec/sec", "value": audio_total / (ms_max * 1e-3), "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate on tcgen05 (decoder, flow; text encoder with two-term fp16 splits); f32 elsewhere",
            "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": args.batch, "audio_s_per_step_per_gpu": audio_s_step,
                       "frames_per_step_per_gpu": int(frames), "weights": "random-init tsukuyomi-shaped JP-Extra, seed 0",
                       "l2": "activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                       "region_ms_per_step_rank0": {"text": text_ms, "flow": flow_ms, "decoder": dec_ms}},
            "e2e": {"value": e2e_audio_total / (e2e_ms_max * 1e-3), "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": d2h, "replicas_per_gpu": R, "single_replica_rank0": e2e_single,
                    "api": "sbv2_synthesize_batch (host buffers in, pinned host waveforms out)"},
            "gpu_launches": launches_total,
            "replicas_per_gpu": R,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "umma_conv_kernel + umma_pair_kernel (HiFi-GAN decoder, timed region = whole decoder)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic, "peak_source": peak_src},
        }
        if not args.no_cpu_baseline:
            n = 3
            rate, a_s, dt = cpu_oracle_rate(oracle, raw, n, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                    "sample": f"{n} utterances (~{a_s:.0f} s audio) of the same batch, batch 1 each; "
                                              "PyTorch-CPU restatement (ONNX Runtime unavailable offline)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
