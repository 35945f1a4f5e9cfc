"""Coefficients of gelu_logistic2 (csrc/common.cuh): gelu(x) = x / (1 + exp(-x p(x^2))), p of degree 4 fitted (Lawson-
weighted least squares = minimax) to the exact erf GELU on |x| <= 9; prints the coefficients scaled by -log2(e) and the
maximum absolute error of an fp32 evaluation against fp64 over |x| <= 14.  CPU only (numpy + scipy)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import ndtr


def gelu(x):
    return x * ndtr(x)


def q(c, x):
    t, p = x * x, c[-1]
    for k in range(len(c) - 2, -1, -1):
        p = p * t + c[k]
    return x * p


xs = np.linspace(1e-3, 9.0, 8000)


def resid(c):
    v = q(c, xs)
    with np.errstate(over="ignore"):
        return np.concatenate([xs / (1 + np.exp(-v)) - gelu(xs), -xs / (1 + np.exp(v)) - gelu(-xs)])


c = np.array([1.5957, 0.0713, 0.0, 0.0, 0.0])
w = np.ones(2 * len(xs))
for _ in range(100):
    c = least_squares(lambda cc: resid(cc) * w, c, xtol=1e-15, ftol=1e-15, gtol=1e-15).x
    e = np.abs(resid(c))
    w = w * (0.5 + e / e.max())
    w /= w.mean()
cf = (-c * np.log2(np.e)).astype(np.float32)
x = np.linspace(-14, 14, 2800001).astype(np.float32)
t, p = (x * x).astype(np.float32), np.full_like(x, cf[-1])
for k in range(len(cf) - 2, -1, -1):
    p = (p * t + cf[k]).astype(np.float32)
with np.errstate(over="ignore"):
    y = (x * (np.float32(1) / (np.float32(1) + np.exp2((x * p).astype(np.float32)).astype(np.float32)))).astype(np.float32)
err = np.abs(y.astype(np.float64) - gelu(x.astype(np.float64)))
print("p(t) coefficients * -log2(e), constant term first:", [f"{float(v):.9g}" for v in cf])
print(f"max |error| (fp32 evaluation, |x| <= 14): {err.max():.3e} at x = {float(x[err.argmax()]):.3f}")
