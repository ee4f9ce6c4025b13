"""End-to-end (host buffers in -> pinned host waveforms out) throughput of sbv2_synthesize_batch on one GPU for 1..3 model
replicas (one host thread + one stream each), with page-locked and with pageable inputs, plus the host-side phases of one
call (upload / run / download) while the other replica is busy.  Usage: python tools/e2e_replicas.py [steps]"""
import os
import sys
import threading
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sbv2-api_b200"))
import bench  # noqa: E402
from oracle import vits as ov  # noqa: E402
import sbv2_b200 as S  # noqa: E402
from sbv2_b200 import assets  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
hp = ov.HParams()
oracle = ov.build_model(hp, seed=0)
onnx = assets.synth_onnx(ov.state_dict_numpy(oracle), hp.upsample_rates, hp.resblock_dilation_sizes)
models = [S.Model(onnx, bert=False) for _ in range(3)]
utts, _ = bench.make_batch(hp, 32, seed=100)
pinned = [dict(u, bert=S.pinned_copy(u["bert"])) for u in utts]
audio_s = sum(a.size for a in models[0].synthesize_batch(utts)) / 44100
for m in models:
    for _ in range(2):
        m.synthesize_batch(pinned)
        m.synthesize_batch(utts)

for label, inp in (("pinned", pinned), ("pageable", utts)):
    for R in (1, 2, 3):
        counts = [steps // R + (1 if i < steps % R else 0) for i in range(R)]
        go = threading.Barrier(R + 1)

        def worker(i):
            go.wait()
            for _ in range(counts[i]):
                models[i].synthesize_batch(inp)

        ths = [threading.Thread(target=worker, args=(i,)) for i in range(R)]
        for th in ths:
            th.start()
        go.wait()
        t0 = time.perf_counter()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        print(f"{label:9s} R={R}: {steps * audio_s / dt:8.0f} audio-s/s  ({1e3 * dt / steps:.2f} ms per batch)")

# phases of one call on replica 0 while replica 1 loops
stop = False


def bg():
    while not stop:
        models[1].synthesize_batch(pinned)


for busy in (False, True):
    th = threading.Thread(target=bg)
    stop = not busy
    th.start()
    time.sleep(0.2 if busy else 0.0)
    up = run = down = 0.0
    n = 8
    for _ in range(n):
        t0 = time.perf_counter()
        db = S.DeviceBatch(models[0], pinned)
        t1 = time.perf_counter()
        db.run()
        t2 = time.perf_counter()
        db.download()
        t3 = time.perf_counter()
        db.close()
        up += t1 - t0
        run += t2 - t1
        down += t3 - t2
    stop = True
    th.join()
    print(f"phases, other replica {'busy' if busy else 'idle'}: upload {1e3 * up / n:.2f} ms, run (host returns) {1e3 * run / n:.2f} ms, "
          f"download (incl. wait for the GPU) {1e3 * down / n:.2f} ms")
