"""clock64 timeline of the tensor-core flow attention kernel (performance debugging).
Usage: python tools/attn_trace.py [T] [n_utt] [heads]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sbv2-api_b200"))
import sbv2_b200 as S  # noqa: E402

fn = S.debug_lib().sbv2_debug_attn_trace
fn.restype = C.c_int
fn.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_float)]


def run(T, n, heads, show=True):
    tr = np.zeros((64, 8), np.int64)
    ms = C.c_float()
    st = fn(T, n, heads, tr.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(ms))
    if st:
        print("ERROR", S.debug_lib().sbv2_last_error().decode())
        return
    nkt = (T + 127) // 128
    flops = 4.0 * T * T * 96 * heads * n * 1.5  # pass A recomputes QK^T
    print(f"T={T} n={n} heads={heads}: {ms.value * 1e3:.1f} us/launch, {flops / ms.value / 1e9:.0f} TFLOP/s issued; "
          f"CTAs {nkt * heads * n} -> {nkt * heads * n / 148:.2f} waves")
    if not show:
        return
    names = ["ctl_K_ready", "ctl_S_free", "ctl_P_ready", "ctl_iter_end", "sm_S_ready", "sm_P_free", "sm_done"]
    base = tr[tr > 0].min()
    for i in range(2 * nkt):
        print("  tile %2d: " % i + "  ".join(f"{names[e]}={(tr[i, e] - base) if tr[i, e] else -1:7d}" for e in range(7)))


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    T = a[0] if len(a) > 0 else 861
    n = a[1] if len(a) > 1 else 32
    heads = a[2] if len(a) > 2 else 2
    run(T, n, heads)
    run(T, 1, 1, show=False)
    run(896, 37, 2, show=False)  # exactly 7 * 2 * 37 = 518 = 3.5 waves
