"""Fit the odd polynomial used by the GEMM epilogue for erf-GELU:  erf(x/sqrt2) ~= xc * Q(xc^2), xc = clamp(x, +-XC).
Weighted least squares on Chebyshev nodes (near-minimax), rescaled so that the clamped value is exactly 1.
Prints C++ constants and the max error of gelu() evaluated in float32 arithmetic."""
import math
import numpy as np

XC, DEG = 3.0 * math.sqrt(2.0), 8
erf = np.vectorize(math.erf)
n = 2000
u = (np.cos(np.pi * (np.arange(n) + 0.5) / n) + 1) / 2 * XC * XC
x = np.sqrt(u)
f = erf(x / math.sqrt(2)) / x
A = np.vander(u, DEG + 1, increasing=True)
coef, *_ = np.linalg.lstsq(A * x[:, None], f * x, rcond=None)
coef = coef / (XC * np.polyval(coef[::-1], XC * XC))          # exact saturation at the clamp
c32 = coef.astype(np.float32)
print("constexpr float kGeluClamp = %.9ef;" % XC)
print("constexpr float kGeluC[%d] = {%s};" % (DEG + 1, ", ".join("%.9ef" % c for c in c32)))

xs = np.linspace(-8, 8, 1600001).astype(np.float32)
xc = np.clip(xs, np.float32(-XC), np.float32(XC))
u32 = xc * xc
p = np.full_like(xs, c32[DEG])
for k in range(DEG - 1, -1, -1):
    p = p * u32 + c32[k]
e = xc * p
h = np.float32(0.5) * xs
g = h * e + h
ref = 0.5 * xs.astype(np.float64) * (1 + erf(xs.astype(np.float64) / math.sqrt(2)))
print("max |erf err| %.3e" % np.abs(e - erf(xs.astype(np.float64) / math.sqrt(2))).max())
print("max |gelu err| %.3e ; max rel-to-bf16-ulp %.3e" % (np.abs(g - ref).max(), (np.abs(g - ref) / (np.abs(ref) * 2 ** -8 + 1e-30)).max()))
