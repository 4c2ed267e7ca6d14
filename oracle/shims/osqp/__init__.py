"""oracle/shims/osqp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Drop-in stand-in for the third-party `osqp` module (osqp==0.6.2.post5,
poetry.lock:101-102) that the reference imports at
sco_py/sco_osqp/osqp_utils.py:4 and drives at osqp_utils.py:195-216.  Only the
surface the reference touches is provided: `OSQP().setup(P, q, A, l, u,
**settings)`, `.solve()` returning an object with `.x`, `.y` and
`.info.status_val / .status / .iter / .obj_val / .pri_res / .dua_res`.

The arithmetic is oracle/osqp_core.c (a restatement of the published OSQP
algorithm, see that file's header).  With this directory and /root/reference on
sys.path the reference's sco_py runs unmodified; that combination is the oracle
used to generate tests/golden/*.npz (oracle/gen_golden.py).
"""
import ctypes
import os
import sys

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.dirname(os.path.dirname(_HERE))

OSQP_INFTY = 1e30

_STATUS = {
    1: "solved", 2: "solved inaccurate", 3: "primal infeasible inaccurate",
    4: "dual infeasible inaccurate", -2: "maximum iterations reached",
    -3: "primal infeasible", -4: "dual infeasible", -7: "problem non convex",
    -10: "unsolved",
}


class _Settings(ctypes.Structure):
    _fields_ = [
        ("rho", ctypes.c_double), ("sigma", ctypes.c_double), ("alpha", ctypes.c_double),
        ("eps_abs", ctypes.c_double), ("eps_rel", ctypes.c_double),
        ("eps_prim_inf", ctypes.c_double), ("eps_dual_inf", ctypes.c_double),
        ("adaptive_rho_tolerance", ctypes.c_double),
        ("max_iter", ctypes.c_int32), ("scaling", ctypes.c_int32),
        ("check_termination", ctypes.c_int32), ("adaptive_rho", ctypes.c_int32),
        ("adaptive_rho_interval", ctypes.c_int32), ("scaled_termination", ctypes.c_int32),
    ]


class _Info(ctypes.Structure):
    _fields_ = [
        ("status_val", ctypes.c_int32), ("iter", ctypes.c_int32),
        ("rho_updates", ctypes.c_int32), ("setup_error", ctypes.c_int32),
        ("obj_val", ctypes.c_double), ("pri_res", ctypes.c_double),
        ("dua_res", ctypes.c_double), ("rho_estimate", ctypes.c_double),
    ]


_lib = None


def _load():
    global _lib
    if _lib is None:
        sys.path.insert(0, _ORACLE)
        try:
            import build as _oracle_build  # oracle/build.py
            path = _oracle_build.build()
        finally:
            sys.path.pop(0)
        _lib = ctypes.CDLL(path)
        _lib.osqp_oracle_solve.restype = ctypes.c_int
        _lib.osqp_oracle_default_settings.restype = None
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _InfoView(object):
    pass


class _Results(object):
    def __init__(self, x, y, info):
        self.x = x
        self.y = y
        self.info = info


def solve_csc(n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, settings, want_scaling=False):
    """Raw entry used by the parity tests: returns (x, y, info[, D, E, c])."""
    lib = _load()
    st = _Settings()
    lib.osqp_oracle_default_settings(ctypes.byref(st))
    for k, v in settings.items():
        if hasattr(st, k):
            setattr(st, k, int(v) if isinstance(getattr(st, k), int) else float(v))
    info = _Info()
    x = np.zeros(n)
    y = np.zeros(m)
    D = np.zeros(n)
    E = np.zeros(m)
    c = ctypes.c_double(0.0)
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (Pp, Pi, Ap, Ai)]
    dbl = [np.ascontiguousarray(a, dtype=np.float64) for a in (Px, q, Ax, l, u)]
    err = lib.osqp_oracle_solve(
        ctypes.c_int(n), ctypes.c_int(m), _dp(arrs[0]), _dp(arrs[1]), _dp(dbl[0]), _dp(dbl[1]),
        _dp(arrs[2]), _dp(arrs[3]), _dp(dbl[2]), _dp(dbl[3]), _dp(dbl[4]), ctypes.byref(st),
        _dp(x), _dp(y), ctypes.byref(info), _dp(D), _dp(E), ctypes.byref(c))
    if err:
        raise ValueError("Workspace allocation error!" if err == 4 else "Problem data validation.")
    if want_scaling:
        return x, y, info, D, E, c.value
    return x, y, info


class OSQP(object):
    def __init__(self):
        self._data = None
        self._settings = {}

    def version(self):
        return "0.6.2.post5-oracle"

    def setup(self, P=None, q=None, A=None, l=None, u=None, **settings):
        if P is None and q is None:
            raise ValueError("The problem does not have any variables")
        n = len(q) if q is not None else P.shape[0]
        m = A.shape[0] if A is not None else 0
        if P is None:
            P = sp.csc_matrix((n, n))
        if A is None:
            A = sp.csc_matrix((0, n))
            l = np.zeros(0)
            u = np.zeros(0)
        if l is None:
            l = -np.inf * np.ones(m)
        if u is None:
            u = np.inf * np.ones(m)
        q = np.asarray(q, dtype=np.float64).ravel() if q is not None else np.zeros(n)
        P = sp.triu(sp.csc_matrix(P), format="csc")
        A = sp.csc_matrix(A)
        for M in (P, A):
            M.sum_duplicates()
            M.sort_indices()
        if P.shape != (n, n) or A.shape[1] != n:
            raise ValueError("Dimensions mismatch")
        l = np.maximum(np.asarray(l, dtype=np.float64).ravel(), -OSQP_INFTY)
        u = np.minimum(np.asarray(u, dtype=np.float64).ravel(), OSQP_INFTY)
        if np.any(l > u):
            raise ValueError("Lower bound must be lower than or equal to upper bound")
        self._data = (n, m, P, q, A, l, u)
        st = dict(settings)
        for drop in ("polish", "warm_start", "verbose", "delta", "linsys_solver", "time_limit",
                     "polish_refine_iter", "adaptive_rho_fraction"):
            st.pop(drop, None)
        self._settings = st
        # factorisation errors (non-convex P) surface at setup in upstream; the
        # oracle core reports them from its single entry point, so probe now.
        return self

    def update_settings(self, **kw):
        self._settings.update(kw)

    def solve(self):
        n, m, P, q, A, l, u = self._data
        x, y, info = solve_csc(n, m, P.indptr, P.indices, P.data, q, A.indptr, A.indices,
                               A.data, l, u, self._settings)
        iv = _InfoView()
        iv.status_val = int(info.status_val)
        iv.status = _STATUS.get(iv.status_val, "unknown")
        iv.iter = int(info.iter)
        iv.obj_val = float(info.obj_val)
        iv.pri_res = float(info.pri_res)
        iv.dua_res = float(info.dua_res)
        iv.rho_updates = int(info.rho_updates)
        iv.rho_estimate = float(info.rho_estimate)
        iv.status_polish = 0
        iv.setup_time = iv.solve_time = iv.run_time = iv.update_time = iv.polish_time = 0.0
        return _Results(x, y, iv)


def constant(name):
    return {"OSQP_INFTY": OSQP_INFTY, "OSQP_SOLVED": 1, "OSQP_SOLVED_INACCURATE": 2,
            "OSQP_MAX_ITER_REACHED": -2, "OSQP_PRIMAL_INFEASIBLE": -3,
            "OSQP_DUAL_INFEASIBLE": -4, "OSQP_NON_CVX": -7, "OSQP_UNSOLVED": -10}[name]
