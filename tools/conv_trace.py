"""Per-role clock64 timeline of the persistent conv kernel (performance debugging).
Usage: python tools/conv_trace.py T cin cout k dil [res] [accum_mode]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sbv2-api_b200"))
import sbv2_b200 as S  # noqa: E402

fn = S.debug_lib().sbv2_debug_conv_trace
fn.restype = C.c_int
fn.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_int,
               C.POINTER(C.c_float), C.POINTER(C.c_int)]


def run(T, cin, cout, k, dil, res=0, acc=0, mt=16, show=True):
    nb = 2
    tr = np.zeros((nb, 64, 8), np.int64)
    ms = C.c_float()
    cfg = (C.c_int * 8)()
    st = fn(T, cin, cout, k, dil, mt, res, acc, tr.ctypes.data_as(C.POINTER(C.c_longlong)), nb, C.byref(ms), cfg)
    if st != 0:
        print("ERROR", S.debug_lib().sbv2_last_error().decode())
        return
    mtv, nbv, aslots, nst, sps, resid, smem, nkc = list(cfg)
    flops = 2.0 * T * cin * cout * k
    byts = T * (cin * 2 + cout * 2 * (1 + res) + (cout * 4 * (2 if acc >= 2 else 1) if acc else 0))
    print(f"T={T} cin={cin} cout={cout} k={k} d={dil} res={res} acc={acc}: {ms.value * 1e3:.1f} us  "
          f"{flops / ms.value / 1e9:.0f} TFLOP/s  {byts / ms.value / 1e6:.0f} GB/s algorithmic | mt={mtv} nb={nbv} a_slots={aslots} "
          f"stages={nst} sps={sps} b_resident={resid} smem={smem} nkc={nkc}")
    if not show:
        return
    t = tr[0]
    base = t[0, 0]
    names = ["prod_start", "prod_issued", "mma_acc_free", "mma_A_ready", "mma_commit", "epi_acc_full", "epi_done"]
    n = int((t[:, 4] > 0).sum())
    print("  block 0 items:", n, " (cycles relative to first producer start)")
    for i in list(range(min(n, 6))) + list(range(max(6, n - 3), n)):
        print("   item %2d: " % i + "  ".join(f"{names[e]}={t[i, e] - base:7d}" for e in range(7)))
    if n > 4:
        per = (t[n - 1, 4] - t[2, 4]) / (n - 3)
        print(f"  steady-state period {per:.0f} cycles/item; mma busy (A_ready->commit) {np.mean(t[2:n, 4] - t[2:n, 3]):.0f}; "
              f"epilogue (acc_full->done) {np.mean(t[2:n, 6] - t[2:n, 5]):.0f}; mma wait for A {np.mean(t[2:n, 3] - t[2:n, 2]):.0f}; "
              f"mma wait for acc {np.mean(t[3:n, 2] - t[2:n - 1, 4]):.0f}")


if __name__ == "__main__":
    if len(sys.argv) > 5:
        a = [int(x) for x in sys.argv[1:]]
        run(*a)
    else:
        run(1400000, 128, 128, 3, 1)
        run(1400000, 128, 128, 3, 1, res=1)
        run(1400000, 128, 128, 11, 5)
        run(175000, 256, 256, 3, 1)
        run(175000, 256, 256, 11, 5, res=1)
        run(2800000, 64, 64, 7, 3)
        run(2800000, 64, 64, 11, 1, res=1)
        run(5600000, 32, 32, 7, 1, res=1)
        run(11200000, 16, 16, 3, 1)
        run(11200000, 16, 16, 11, 1, res=1, acc=2, show=False)
        run(1400000, 128, 128, 7, 1, res=1, acc=3, show=False)
