#!/usr/bin/env python
"""Fits erf-form GELU as 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) (csrc/ptx.cuh: gelu_tanh_fit) and reports its error
against the exact 0.5 x (1 + erf(x / sqrt 2)) of nn.GELU() (backbone/vit.py:127,132)."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

v = np.linspace(-8, 8, 200001)
g = 0.5 * v * (1 + erf(v / np.sqrt(2)))
f = lambda p: np.abs(0.5 * v * (1 + np.tanh(v * (p[0] + p[1] * v * v + p[2] * v ** 4))) - g).max()
r = minimize(f, [0.7978845608, 0.0356774, 0.0], method="Nelder-Mead", options=dict(xatol=1e-12, fatol=1e-12, maxiter=20000))
print("coefficients", r.x, "max abs error", r.fun)
print("classic 0.044715 tanh-GELU max abs error", f([0.7978845608, 0.7978845608 * 0.044715, 0.0]))
