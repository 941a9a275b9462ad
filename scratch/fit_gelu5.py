import numpy as np
from scipy.special import ndtr, log_ndtr
def ulp16(v):
    v = np.abs(v)
    e = np.floor(np.log2(np.maximum(v, 2.0**-14)))
    return 2.0 ** (e - 10)
deg = 5
for A, abs_target, ulp_target in ((6.5, 1e-6, 1.0), (6.5, 2e-6, 1.0), (6.5, 1e-6, 2.0), (7.0, 1.5e-6, 0.75)):
    a = np.linspace(0, A, 60001)
    target = log_ndtr(-a) / np.log(2.0) + 1.0
    out_mag = a * ndtr(-a)
    s = np.maximum(out_mag, 1e-30) * np.log(2.0)
    w = np.maximum(np.maximum(s / abs_target, s / (ulp_target * ulp16(out_mag))), 1e-3)
    lam = np.ones_like(a)
    for itr in range(300):
        V = np.vander(a / A, deg + 1, increasing=True)[:, 1:]
        W = (w * lam)[:, None]
        coef, *_ = np.linalg.lstsq(V * W, target * w * lam, rcond=None)
        err = (V @ coef - target) * w
        lam = lam * (1 + 4 * np.abs(err) / np.abs(err).max()); lam /= lam.mean()
    c = np.concatenate([[-1.0], coef / (A ** np.arange(1, deg + 1))])
    xs = np.concatenate([np.linspace(-12, 12, 4000001), np.linspace(-60000, 60000, 200001)]).astype(np.float32)
    ax = np.abs(xs)
    q = np.full_like(ax, np.float32(c[-1]))
    for k in range(deg - 1, -1, -1):
        q = (q * ax + np.float32(c[k])).astype(np.float32)
    with np.errstate(over='ignore', invalid='ignore'):
        out = np.maximum(xs, 0) - ax * np.exp2(q.astype(np.float32))
    ref = xs.astype(np.float64) * ndtr(xs.astype(np.float64))
    e = np.abs(out - ref)
    neg = xs < 0
    ul = e / ulp16(ref)
    sm = (ax < 1) & (ref != 0)
    print(A, abs_target, ulp_target, "-> max abs err %.3g  max fp16-ulps(neg side) %.3g  rel|x|<1 %.3g lead %.3g nan %d" % (np.nanmax(e), np.nanmax(ul[neg]), (e[sm]/np.abs(ref[sm])).max(), c[-1], np.isnan(out).sum()))
    print("  coef", ", ".join(f"{v:.9e}f" for v in c))
