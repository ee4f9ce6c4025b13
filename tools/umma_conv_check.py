"""GPU diagnostic for the tcgen05 conv kernel: compares it with the fp32 CUDA-core conv on the same
fp16-representable inputs.  Usage: python tools/umma_conv_check.py [swap]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sbv2-api_b200"))
import sbv2_b200 as S  # noqa: E402

fn = S.debug_lib().sbv2_debug_conv_compare
pf = C.POINTER(C.c_float)
fn.restype = C.c_int
fn.argtypes = [pf, C.c_int64, C.c_int, pf, pf, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, pf, pf]


def run(T, cin, cout, k, dil, mt=4, res=0, swap=0, identity=False, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((T, cin)).astype(np.float16).astype(np.float32)
    if identity:
        w = np.zeros((cout, cin, k), np.float32)
        for c in range(min(cin, cout)):
            w[c, c, (k - 1) // 2] = 1.0
        b = np.zeros(cout, np.float32)
    else:
        w = (rng.standard_normal((cout, cin, k)) / np.sqrt(cin * k)).astype(np.float16).astype(np.float32)
        b = rng.standard_normal(cout).astype(np.float32) * 0.1
    ou = np.zeros((T, cout), np.float32)
    orf = np.zeros((T, cout), np.float32)
    st = fn(x.ctypes.data_as(pf), T, cin, w.ctypes.data_as(pf), b.ctypes.data_as(pf), cout, k, dil, mt, res,
            ou.ctypes.data_as(pf), orf.ctypes.data_as(pf))
    if st != 0:
        print(f"T={T} cin={cin} cout={cout} k={k} d={dil} mt={mt}: ERROR {S.debug_lib().sbv2_last_error().decode()}")
        return None
    ref = orf.copy()
    if res:
        ref = ref + np.where(x >= 0, x, x * 10.0)
    err = np.abs(ou - ref).max()
    tol = 2e-3 * max(1.0, np.abs(ref).max())
    flag = "OK " if err < tol else "BAD"
    print(f"{flag} T={T} cin={cin} cout={cout} k={k} d={dil} mt={mt} res={res} swap={swap} id={int(identity)}: max|ref| {np.abs(ref).max():.3f} "
          f"err {err:.3g}")
    if err >= tol and identity:
        np.set_printoptions(linewidth=200, precision=2, suppress=True)
        print(" x[0:3,:16]\n", x[0:3, :16], "\n out[0:3,:16]\n", ou[0:3, :16])
        # where does x[0,:] appear?
        for r in range(min(T, 4)):
            d = np.abs(ou - x[r:r + 1, :cout]).sum(1)
            print("  input row", r, "best matches output row", int(d.argmin()), "residual", float(d.min()))
        bad = np.abs(ou - ref) > tol
        print("  bad rows:", np.unique(np.nonzero(bad)[0])[:20], " bad cols:", np.unique(np.nonzero(bad)[1])[:32])
    return err


if __name__ == "__main__":
    swap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    run(200, 16, 16, 1, 1, mt=1, swap=swap, identity=True)
    run(200, 64, 64, 1, 1, mt=1, swap=swap, identity=True)
    run(200, 64, 64, 3, 1, mt=1, swap=swap, identity=True)
    run(200, 16, 16, 1, 1, mt=1, swap=swap)
    run(300, 64, 64, 3, 1, mt=1, swap=swap)
    run(300, 64, 64, 3, 1, mt=2, swap=swap)
    run(1000, 128, 128, 7, 3, mt=4, swap=swap)
    run(700, 256, 256, 11, 5, mt=2, swap=swap)
    run(1500, 32, 32, 11, 1, mt=4, swap=swap)
    run(1500, 16, 16, 3, 5, mt=4, swap=swap)
    run(300, 192, 512, 7, 1, mt=2, swap=swap)
    run(700, 128, 128, 3, 1, mt=4, res=1, swap=swap)
    run(900, 512, 256, 3, 1, mt=1, swap=swap)
    # cluster-multicast coverage (SBV2_B200_CLUSTER=2/4): many tiles, odd tile counts (dummy tiles), streaming weights
    run(647, 192, 768, 3, 1, mt=1, swap=swap)
    run(5000, 192, 768, 3, 1, mt=2, swap=swap)
    run(3000, 768, 192, 3, 1, mt=2, swap=swap)
    run(4096 + 128 * 3 + 5, 1024, 1024, 1, 1, mt=1, swap=swap)
    run(129, 1024, 1024, 1, 1, mt=1, swap=swap)
    run(9000, 128, 128, 11, 5, mt=2, res=1, swap=swap)
    run(9000, 256, 256, 7, 3, mt=1, swap=swap)
