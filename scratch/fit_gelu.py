"""Fit q(a) so that Phi(-a) ~= 2^q(a) on a in [0, A]; gelu(x) = max(x,0) - |x| * 2^q(|x|).
Minimise the ABSOLUTE error of a * 2^q(a) (what the activation output sees) with Lawson-style reweighting."""
import numpy as np
from scipy.special import ndtr, log_ndtr
A = 6.0
for deg in (6, 7, 8, 9):
    a = np.linspace(0, A, 40001)
    target = log_ndtr(-a) / np.log(2.0)              # log2 Phi(-a)
    base_w = np.maximum(a, 0.02) * ndtr(-a) * np.log(2.0)   # d(out)/dq
    # also demand relative accuracy near zero: weight floor
    w = base_w.copy()
    lam = np.ones_like(a)
    for itr in range(60):
        V = np.vander(a / A, deg + 1, increasing=True)
        W = (w * lam)[:, None]
        coef, *_ = np.linalg.lstsq(V * W, target * w * lam, rcond=None)
        err = (V @ coef - target) * w
        lam = lam * (1 + 4 * np.abs(err) / np.abs(err).max()); lam /= lam.mean()
    c = coef / (A ** np.arange(deg + 1))
    # evaluate in float32 Horner
    xs = np.linspace(-9, 9, 2000001).astype(np.float32)
    ax = np.minimum(np.abs(xs), np.float32(A))
    q = np.full_like(ax, np.float32(c[-1]))
    for k in range(deg - 1, -1, -1):
        q = (q * ax + np.float32(c[k])).astype(np.float32)
    out = np.maximum(xs, 0) - np.abs(xs) * np.exp2(q.astype(np.float32))
    ref = xs.astype(np.float64) * ndtr(xs.astype(np.float64))
    e = np.abs(out - ref)
    rel_small = np.abs(out - ref)[np.abs(xs) < 1] / np.maximum(np.abs(ref[np.abs(xs) < 1]), 1e-30)
    print(deg, "max abs err", e.max(), "at", xs[e.argmax()], "max rel err |x|<1", rel_small.max())
    print("   coef", ", ".join(f"{v:.9e}f" for v in c))
