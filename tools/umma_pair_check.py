"""GPU diagnostic for the fused ResBlock-pair kernel (umma_pair.cu): compares it with two unfused tensor-core
convs and with a numpy fp32 evaluation, on ragged batches; optional timing of both paths.
Usage: python tools/umma_pair_check.py [perf]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sbv2-api_b200"))
import sbv2_b200 as S  # noqa: E402

fn = S.debug_lib().sbv2_debug_pair_compare
pf = C.POINTER(C.c_float)
pi = C.POINTER(C.c_int)
fn.restype = C.c_int
fn.argtypes = [pf, pi, C.c_int, C.c_int, C.c_int, C.c_int, pf, pf, pf, pf, C.c_int, C.c_int, pf, pf, pf, pi, C.POINTER(C.c_longlong)]


def lrelu(v, s=0.1):
    return np.where(v >= 0, v, v * s)


def conv_np(x, w, b, dil):  # x [T, C], w [Co, Ci, k]
    T, _ = x.shape
    k = w.shape[2]
    pad = dil * (k - 1) // 2
    xp = np.zeros((T + 2 * pad, x.shape[1]), np.float32)
    xp[pad:pad + T] = x
    out = np.tile(b[None, :], (T, 1)).astype(np.float32)
    for j in range(k):
        out += xp[j * dil:j * dil + T] @ w[:, :, j].T
    return out


def run(lens, c, k, dil, mrf=0, iters=0, seed=0, check_np=True, trace=False):
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens, np.int32)
    tot = int(lens.sum())
    x = rng.standard_normal((tot, c)).astype(np.float32)
    w1 = (rng.standard_normal((c, c, k)) / np.sqrt(c * k)).astype(np.float16).astype(np.float32)
    w2 = (rng.standard_normal((c, c, k)) / np.sqrt(c * k)).astype(np.float16).astype(np.float32)
    b1 = (rng.standard_normal(c) * 0.1).astype(np.float32)
    b2 = (rng.standard_normal(c) * 0.1).astype(np.float32)
    of = np.zeros((tot, c), np.float32)
    orf = np.zeros((tot, c), np.float32)
    ms = np.zeros(2, np.float32)
    cfg = np.zeros(8, np.int32)
    tr = np.zeros((64, 16), np.int64)
    st = fn(x.ctypes.data_as(pf), lens.ctypes.data_as(pi), len(lens), c, k, dil, w1.ctypes.data_as(pf), b1.ctypes.data_as(pf),
            w2.ctypes.data_as(pf), b2.ctypes.data_as(pf), mrf, iters, of.ctypes.data_as(pf), orf.ctypes.data_as(pf),
            ms.ctypes.data_as(pf), cfg.ctypes.data_as(pi), tr.ctypes.data_as(C.POINTER(C.c_longlong)) if trace else None)
    tag = f"C={c} k={k} d={dil} mrf={mrf} n={len(lens)} rows={tot}"
    if st != 0:
        print(f"ERROR {tag}: {S.debug_lib().sbv2_last_error().decode()}")
        return False
    err_u = float(np.abs(of - orf).max())
    msg = f"{tag} cfg(mt,rows,aslots,nst,sps,res,smem,nkc)={cfg.tolist()} |fused-unfused| {err_u:.3g}"
    ok = err_u < 4e-3 * max(1.0, float(np.abs(orf).max()))
    if check_np:
        y = lrelu(x).astype(np.float16).astype(np.float32)  # stored activation
        xr = np.where(y >= 0, y, y * 10.0)
        ref = np.zeros_like(x)
        o = 0
        for n in lens:
            t1 = lrelu(conv_np(y[o:o + n], w1, b1, dil)).astype(np.float16).astype(np.float32)
            v = conv_np(t1, w2, b2, 1) + xr[o:o + n]
            ref[o:o + n] = v
            o += n
        if mrf:
            ya = y
            yb = x.astype(np.float16).astype(np.float32)
            ref = (np.where(ya >= 0, ya, ya * 10) + np.where(yb >= 0, yb, yb * 10) + ref) / 3.0
        ref = lrelu(ref)
        err = float(np.abs(of - ref).max())
        msg += f" |fused-numpy| {err:.3g} (max|ref| {np.abs(ref).max():.2f})"
        ok = ok and err < 4e-3 * max(1.0, float(np.abs(ref).max()))
    if iters:
        msg += f"  fused {ms[0] * 1e3:.1f} us, unfused {ms[1] * 1e3:.1f} us"
    print(("OK  " if ok else "BAD ") + msg)
    if trace:
        names = ["p1_go", "p1_issued", "p2_go", "p2_issued", "e1_go", "e1_end", "e2_go", "e2_end", "p2_enter", "p2_acc2_free", "w17_e1_go",
                 "w17_e1_end", "w17_e2_go", "w17_e2_end"]
        n = int((tr[:, 7] > 0).sum())
        base = tr[tr > 0].min()
        for i in list(range(max(3, n - 6), n)):
            print("   item %2d: " % i + " ".join(f"{names[e]}={tr[i, e] - base:7d}" for e in range(14)))
        if n > 6:
            print(f"   steady period {(tr[n - 1, 7] - tr[3, 7]) / (n - 4):.0f} cyc/item; e1 {np.mean(tr[3:n, 5] - tr[3:n, 4]):.0f}; e2 "
                  f"{np.mean(tr[3:n, 7] - tr[3:n, 6]):.0f}; p1 issue {np.mean(tr[3:n, 1] - tr[3:n, 0]):.0f}; p2 issue "
                  f"{np.mean(tr[3:n, 3] - tr[3:n, 2]):.0f}")
    return ok


if __name__ == "__main__":
    perf = len(sys.argv) > 1 and sys.argv[1] == "perf"
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        if os.environ.get("SBV2_B200_PAIR_GRID"):
            run([1014 * 60], 16, 11, 5, iters=0, check_np=False, trace=True)
            run([246 * 60], 64, 11, 5, iters=0, check_np=False, trace=True)
            sys.exit(0)
        for c, mul in ((16, 512),):
            for k, d in ((3, 1),):
                run([860 * mul] * 32, c, k, d, iters=3, check_np=False, trace=True)
        sys.exit(0)
    good = True
    if not perf:
        for c in (16, 32, 64, 128):
            for k, d in ((3, 1), (3, 5), (7, 3), (11, 1), (11, 5)):
                good &= run([50, 1, 700, 1300, 129, 2500], c, k, d)
        good &= run([900, 17, 3000], 32, 7, 5, mrf=1)
        good &= run([900, 17, 3000], 16, 11, 3, mrf=1)
        good &= run([40000] * 4, 64, 11, 5)
        good &= run([300000] * 3, 16, 3, 1, check_np=False)
        print("ALL OK" if good else "FAILURES")
        sys.exit(0 if good else 1)
    # bench-like shapes: 32 utterances, T_y ~ 860 frames -> rows = T_y * {64, 128, 256, 512}... per stage
    for c, mul in ((128, 64), (64, 128), (32, 256), (16, 512)):
        rows = 860 * mul
        for k, d in ((3, 1), (3, 5), (7, 3), (11, 5)):
            run([rows] * 32, c, k, d, iters=5, check_np=False)
        run([rows] * 32, c, 11, 5, mrf=1, iters=5, check_np=False)
