"""oracle/shims/numdifftools -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Stand-in for numdifftools==0.9.40 (poetry.lock:76-77), which the reference
imports at sco_py/expr.py:1 and calls at expr.py:67 (`nd.Jacobian(f)(x)`) and
expr.py:108 (`nd.Hessian(f)(x.flatten())`).  The package is absent from this
image, so its published scheme is restated: central differences on a short
geometric step sequence (base step EPS**(1/scale) scaled by
max(log1p|x|, 1), step ratio 2) followed by one Richardson extrapolation
(SURVEY.md Appendix B2).  `scale` follows the documented defaults: 2.5 for
first derivatives, 7.8 for the second-order (Hessian) rule.

Shape contract relied on by the reference: Jacobian of R^n -> R^m is (m, n)
even for m == 1; Hessian of a flat x is (n, n).

Parity with upstream numdifftools is pinned only through the reference's own
tests (tests/sco_osqp/test_expr.py:71-78,101-110,151-161,177-211 at
np.allclose defaults / atol 1e-4); the exact step sequence is NOT pinned.
The CUDA path (sco_py_b200/csrc) restates this same scheme so the two agree
to rounding.
"""
import numpy as np

EPS = np.finfo(float).eps
JAC_BASE_STEP = EPS ** (1.0 / 2.5)
HESS_BASE_STEP = EPS ** (1.0 / 7.8)


def _step_nom(x):
    return np.maximum(np.log1p(np.abs(x)), 1.0)


def jacobian_fd(fun, x):
    """(4 D(h) - D(2h)) / 3 with D the central difference; returns (m, n)."""
    x = np.asarray(x, dtype=float)
    shape = x.shape
    xf = x.ravel().copy()
    n = xf.size
    h0 = JAC_BASE_STEP * _step_nom(xf)
    cols = []
    for j in range(n):
        est = []
        for mult in (1.0, 2.0):
            h = h0[j] * mult
            xp = xf.copy()
            xm = xf.copy()
            xp[j] += h
            xm[j] -= h
            hh = xp[j] - xm[j]  # exactly representable step
            fp = np.asarray(fun(xp.reshape(shape)), dtype=float).ravel()
            fm = np.asarray(fun(xm.reshape(shape)), dtype=float).ravel()
            est.append((fp - fm) / hh)
        cols.append((4.0 * est[0] - est[1]) / 3.0)
    return np.stack(cols, axis=1)


def hessian_fd(fun, x):
    """Central second differences at steps h and 2h, Richardson-combined."""
    x = np.asarray(x, dtype=float)
    shape = x.shape
    xf = x.ravel().copy()
    n = xf.size
    h0 = HESS_BASE_STEP * _step_nom(xf)

    def f(v):
        return float(np.asarray(fun(v.reshape(shape)), dtype=float).ravel()[0])

    f0 = f(xf)
    ests = []
    for mult in (1.0, 2.0):
        h = h0 * mult
        H = np.zeros((n, n))
        for i in range(n):
            e = np.zeros(n)
            e[i] = 2.0 * h[i]
            H[i, i] = (f(xf + e) - 2.0 * f0 + f(xf - e)) / (4.0 * h[i] * h[i])
            for j in range(i + 1, n):
                ei = np.zeros(n)
                ej = np.zeros(n)
                ei[i] = h[i]
                ej[j] = h[j]
                v = (f(xf + ei + ej) - f(xf + ei - ej) - f(xf - ei + ej) + f(xf - ei - ej)) / (
                    4.0 * h[i] * h[j])
                H[i, j] = H[j, i] = v
        ests.append(H)
    return (4.0 * ests[0] - ests[1]) / 3.0


class Jacobian(object):
    def __init__(self, fun, **kwargs):
        self.fun = fun

    def __call__(self, x, *args, **kwargs):
        return jacobian_fd(self.fun, x)


class Gradient(Jacobian):
    def __call__(self, x, *args, **kwargs):
        return jacobian_fd(self.fun, x).ravel()


class Derivative(Jacobian):
    def __call__(self, x, *args, **kwargs):
        x = np.asarray(x, dtype=float)
        return jacobian_fd(self.fun, x.reshape(-1)).reshape(x.shape) if x.ndim else float(
            jacobian_fd(self.fun, x.reshape(1)).ravel()[0])


class Hessian(object):
    def __init__(self, fun, **kwargs):
        self.fun = fun

    def __call__(self, x, *args, **kwargs):
        return hessian_fd(self.fun, x)
