"""What comes next for the decoder's narrow stages, checked on the CPU: folding R time steps of a
conv1d into the N dimension of the tensor-core GEMM ("block-Toeplitz" weights).

  out[t, co] = sum_j sum_ci x[t + (j - (k-1)/2) d, ci] w[j, ci, co]            (the conv as the kernels compute it)

For dilation d the time axis splits into d residue classes; inside a class, u = (t - c) / d is contiguous and the conv is an
ordinary k-tap conv over u.  Folding R consecutive u into N:

  A[r, (j', ci)]        = x_class[R r + j' - (k-1)/2, ci]        j' = 0 .. R + k - 2      (K = (R + k - 1) C)
  W_T[(j', ci), (i, co)] = w[j' - i, ci, co] if 0 <= j' - i < k else 0                    (N = R C)
  out_class[R r + i, co] = (A W_T)[r, (i, co)]

The script (1) verifies the identity numerically against the direct conv for the decoder's shapes, and (2) prices both
mappings with the measured issue model of tcgen05.mma M = 128, kind::f16 (profiles/r1_mma_issue_rate.log: 48 cycles per
MMA for N <= 64, 64 cycles at N = 128, 128 cycles at N = 256; K = 16 per instruction).
Usage: python tools/toeplitz_model.py"""
import numpy as np


def conv_direct(x, w, d):
    T, C = x.shape
    k = w.shape[0]
    out = np.zeros((T, w.shape[2]), np.float64)
    for j in range(k):
        s = (j - (k - 1) // 2) * d
        lo, hi = max(0, -s), min(T, T - s)
        out[lo:hi] += x[lo + s:hi + s] @ w[j]
    return out


def conv_toeplitz(x, w, d, R):
    T, C = x.shape
    k, _, Co = w.shape
    out = np.zeros((T, Co), np.float64)
    WT = np.zeros(((R + k - 1) * C, R * Co), np.float64)
    for jp in range(R + k - 1):
        for i in range(R):
            j = jp - i
            if 0 <= j < k:
                WT[jp * C:(jp + 1) * C, i * Co:(i + 1) * Co] = w[j]
    h = (k - 1) // 2
    for c in range(d):  # residue classes of the time axis
        xc = x[c::d]
        U = xc.shape[0]
        rows = (U + R - 1) // R
        xp = np.zeros((rows * R + k - 1, C), np.float64)
        xp[h:h + U] = xc
        A = np.stack([xp[R * r:R * r + R + k - 1].reshape(-1) for r in range(rows)])
        D = (A @ WT).reshape(rows * R, Co)
        out[c::d] = D[:U]
    return out


def issue_cycles(n):
    return 48 if n <= 64 else (64 if n <= 128 else 128)


def main():
    rng = np.random.default_rng(0)
    worst = 0.0
    for C, k, d, R in [(16, 3, 1, 4), (16, 7, 3, 4), (16, 11, 5, 8), (32, 11, 5, 2), (32, 7, 1, 4), (64, 3, 3, 2)]:
        x = rng.standard_normal((997, C))
        w = rng.standard_normal((k, C, C))
        e = np.abs(conv_toeplitz(x, w, d, R) - conv_direct(x, w, d)).max()
        worst = max(worst, e)
    print(f"identity check on 6 shapes (incl. dilations 3 and 5): max |toeplitz - direct| = {worst:.2e}")
    print()
    print("MMA issue time per ResBlock pair (conv1 dilated + conv2), bench workload: 21 867 latent frames per step,")
    print("rows = frames * upsampling; cycles at 148 SMs, 1.85 GHz; 'now' = M=128 x N=C MMAs, one per tap and 16 channels")
    frames, sms, ghz = 21867, 148, 1.85
    total_now = total_new = 0.0
    for C, up in [(64, 128), (32, 256), (16, 512)]:
        rows = frames * up
        for k in (3, 7, 11):
            for R in (1, 2, 4, 8):
                if R * C > 256:
                    continue
                n = R * C
                # per 128 GEMM rows (= 128 R time steps): (R + k - 1) * C / 16 MMAs per conv
                mmas_per_conv = (rows / (128 * R)) * (R + k - 1) * C / 16
                us = 2 * mmas_per_conv * issue_cycles(n) / sms / (ghz * 1e3)
                tag = "now" if R == 1 else f"R={R}"
                print(f"  C={C:3d} k={k:2d} {tag:4s} N={n:3d}: {us:7.1f} us per pair")
            best = min(2 * (rows / (128 * R)) * (R + k - 1) * C / 16 * issue_cycles(R * C) / sms / (ghz * 1e3)
                       for R in (1, 2, 4, 8) if R * C <= 256)
            now = 2 * (rows / 128) * k * C / 16 * 48 / sms / (ghz * 1e3)
            total_now += 3 * now   # three dilations per kernel size
            total_new += 3 * best
    print()
    print(f"issue floor of the three narrow stages: {total_now / 1e3:.2f} ms now -> {total_new / 1e3:.2f} ms with the best fold per shape "
          f"(measured today: 9.8 ms = issue floor of the k = 7 / 11 pairs + ~0.1 ms each, ~0.24 ms for every epilogue-bound k = 3 pair; "
          f"the two epilogues per output element, ~0.24 ms x 27 pairs = 6.5 ms per step, are the bound after the fold)")


if __name__ == "__main__":
    main()
