"""CPU: the GELU the fc1 epilogue evaluates (csrc/ptx.cuh gelu_tanh_fit), restated in float32 numpy with the same
operation order, against the exact erf GELU of nn.GELU() (reference backbone/vit.py:127,132).

Pins two claims: max |error| <= 3e-5 on [-8, 8] (below the bf16 rounding that follows wherever |GELU| >= 0.02), and -- the
round-1 advisor finding -- no sign flip of the polynomial for large |x| (x^2 is clamped to 64)."""
import re
import os

import numpy as np
from scipy.special import erf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coefficients():
    src = open(os.path.join(ROOT, "easy_vitpose_b200", "csrc", "ptx.cuh")).read()
    body = src[src.index("float gelu_tanh_fit(float x)"):]
    body = body[:body.index("}")]
    c = [np.float32(v) for v in re.findall(r"(-?\d\.\d+e[-+]\d+)f", body)]
    clamp = np.float32(re.search(r"fminf\(x \* x, (\d+\.\d+)f\)", body).group(1))
    assert len(c) == 3
    return c, clamp


def gelu_fit_f32(x: np.ndarray) -> np.ndarray:
    (c2, c1, c0), clamp = _coefficients()
    x = x.astype(np.float32)
    x2 = np.minimum(x * x, clamp)
    p = (c2 * x2 + c1).astype(np.float32)
    p = (p * x2 + c0).astype(np.float32)
    t = np.tanh((x * p).astype(np.float32)).astype(np.float32)
    hx = np.float32(0.5) * x
    return (hx * t + hx).astype(np.float32)


def gelu_erf(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.float64)
    return 0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))


def test_fit_error_on_the_fitted_range():
    x = np.linspace(-8, 8, 400001)
    err = np.abs(gelu_fit_f32(x) - gelu_erf(x))
    print("max |gelu_fit - gelu_erf| on [-8,8]:", err.max())
    assert err.max() < 3e-5
    # where |GELU| >= 0.02 the fit error is below half a bf16 step of the result (2^-9 relative); nearer to zero
    # (x in about [-4, -2.5], |GELU| < 0.02) it stays an ABSOLUTE 2.6e-5, i.e. 1e-5 of the hidden activation's range
    big = np.abs(gelu_erf(x)) >= 0.02
    assert np.all(err[big] <= 2.0 ** -9 * np.abs(gelu_erf(x))[big])


def test_no_sign_flip_outside_the_fitted_range():
    """without the clamp GELU(11.5) came out as 0.0003 and GELU(-12) as -12."""
    x = np.concatenate([np.linspace(-40, -8, 6401), np.linspace(8, 40, 6401), [-1e4, 1e4, -65504.0, 65504.0]])
    y, ref = gelu_fit_f32(x), gelu_erf(x)
    assert np.all(np.abs(y - ref) <= 1e-6 * np.maximum(1.0, np.abs(ref)))
    for v in (11.0, 11.5, 12.0, 15.0, 30.0):
        assert gelu_fit_f32(np.array([v]))[0] == np.float32(v)
        assert gelu_fit_f32(np.array([-v]))[0] == 0.0
