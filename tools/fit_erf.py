"""Fit of the single-exp2 erf used by the CUDA epilogues (csrc/common.cuh erf_abs_fast):
erf(t) = 1 - 2^p(t), p(t) = t*(c1 + c2 t + ... + c5 t^4) ~ log2(erfc(t)) on [0, 4], Lawson-weighted least squares
so that the ABSOLUTE error of erf is equi-oscillating; reports the fp32 max error over [0, 6]."""
import numpy as np
from scipy.special import erf, erfc

DEG = 5
t = np.linspace(0, 4.0, 4001)
y = np.log2(erfc(t))
A = np.stack([t ** k for k in range(1, DEG + 1)], 1)
w = erfc(t) * np.log(2)
ww = np.ones_like(t)
for _ in range(200):
    c = np.linalg.lstsq(A * (w * ww)[:, None], y * w * ww, rcond=None)[0]
    r = np.abs((1 - np.exp2(A @ c)) - erf(t))
    ww = ww * (0.5 + r / r.max())
    ww /= ww.mean()
c32 = c.astype(np.float32)
tt = np.linspace(0, 6, 600001).astype(np.float32)
p = np.zeros_like(tt)
for k in range(DEG - 1, -1, -1):
    p = ((p + c32[k]).astype(np.float32) * tt).astype(np.float32)
e = (np.float32(1) - np.exp2(p)).astype(np.float32)
err = np.abs(e.astype(np.float64) - erf(tt.astype(np.float64)))
print("coefficients c1..c5:", [float(x) for x in c32])
print("max |erf error| (fp32 Horner):", err.max(), "at t =", tt[err.argmax()])
