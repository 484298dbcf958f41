"""Fit of the polynomial used by the GEMM epilogue's exact-erf GELU (metatransformer_b200/csrc/mt_gemm.cu: gelu_erf).

    gelu(x) = x*Phi(x) = relu(x) - 0.5*|x|*erfc(|x|/sqrt(2)),      erfc(t/sqrt(2)) ~= 2^(t*p(t)),  t = min(|x|, 12)

p = degree-4 polynomial fitted (iteratively re-weighted least squares ~ minimax) so that the *GELU* absolute error is
minimised.  Run to re-derive the coefficients and print the achieved error (TEST INFRASTRUCTURE; needs scipy).
"""
import numpy as np
from numpy.polynomial import polynomial as P
from scipy.optimize import least_squares
from scipy.special import erfc

COEFFS = [-1.1510006189346313, -0.4595957100391388, -0.052146803587675095, 0.007198806386440992,
          -0.00048811757005751133]   # c0..c4 as compiled into the kernel


def gelu_err(c, t):
    return 0.5 * t * (np.exp2(t * P.polyval(t, c)) - erfc(t / np.sqrt(2)))


def fit(n=4, T=6.0):
    t = np.linspace(1e-6, T, 20001)
    h = np.log2(np.maximum(erfc(t / np.sqrt(2)), 1e-300)) / t
    c = P.polyfit(t[t < 5.5], h[t < 5.5], n)
    w = np.ones_like(t)
    for _ in range(60):
        c = least_squares(lambda cc: gelu_err(cc, t) * w * 1e6, c, method="lm", max_nfev=2000).x
        e = np.abs(gelu_err(c, t))
        w = w * (1 + 2 * e / e.max())
        w /= w.mean()
    return c


def max_error(c):
    t = np.linspace(0, 12, 400001)
    c32 = np.asarray(c, dtype=np.float32)
    t32 = t.astype(np.float32)
    p = np.float32(c32[-1]) * np.ones_like(t32)
    for k in range(len(c32) - 2, -1, -1):
        p = p * t32 + c32[k]
    e = np.exp2((t32 * p).astype(np.float64))
    return float(np.abs(0.5 * t * (e - erfc(t / np.sqrt(2)))).max())


if __name__ == "__main__":
    print("compiled coefficients: max |gelu error| =", max_error(COEFFS))
    c = fit()
    print("refit:", [float(np.float32(v)) for v in c], "max |gelu error| =", max_error(c))
