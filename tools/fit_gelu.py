"""Fits erf(z) ~= tanh(z * (a0 + a1 z^2 + a2 z^4)) (minimax over z in [0, 4.5]) - the constants of gelu_f in csrc/ptx.cuh
are a_k / (sqrt 2)^(2k+1).  Prints the coefficients and the resulting erf / GELU errors."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

z = np.linspace(0, 4.5, 20001)
f = lambda p, z=z: np.tanh(z * (p[0] + z * z * (p[1] + z * z * p[2])))
r = minimize(lambda p: np.max(np.abs(f(p) - erf(z))), [1.1283792, 0.1, 0.0], method="Nelder-Mead",
             options=dict(xatol=1e-10, fatol=1e-12, maxiter=40000, maxfev=40000))
a = r.x
print("a =", a, "max |erf err| =", r.fun)
s2 = np.sqrt(2.0)
print("b =", a[0] / s2, a[1] / (2 * s2), a[2] / (4 * s2))
x = np.linspace(-8, 8, 40001)
x2 = np.minimum(x * x, 40.0)
b = [a[0] / s2, a[1] / (2 * s2), a[2] / (4 * s2)]
g = 0.5 * x * (1 + np.tanh(x * (b[0] + x2 * (b[1] + x2 * b[2]))))
print("max |gelu err| =", np.max(np.abs(g - 0.5 * x * (1 + erf(x / s2)))))
