"""Expression classes with the names, constructor arguments and host-side behaviour of
sco_py/expr.py (Expr :22, AffExpr :164, QuadExpr :189, AbsExpr :221, HingeExpr :243, CompExpr :267,
EqExpr :300, LEqExpr :335, LExpr :374, BoundExpr :421), plus the closed *device families* the B200
engine can evaluate.

A reference program describes a nonlinear constraint with a black-box callable,
``LEqExpr(Expr(f, grad), val)``.  A GPU kernel cannot call Python, so the drop-in asks for ONE
change: say which family `f` belongs to,

    Expr(f, grad)                      ->  QuadFormExpr(P, a)          f_j = 0.5 x'P_j x + a_j'x
                                           CircleDistExpr(T, c, R)     f_tk = R_k - |p_t - c_k|
                                           FK7Expr(n)                  flange position of a 7-link chain
                                           SymExpr(rows, n)            anything written with sco_py_b200.sym
                                                                       (finite-difference derivatives, like a
                                                                       black box without `grad`)

Everything else (Variable, Prob.add_obj_expr / add_cnt_expr, BoundExpr, Solver.solve) is written as
with the reference.  The family classes are still `Expr`s -- eval / grad / convexify work on the host
exactly like the black box they replace (used by the parity tests to build the reference's own Prob
from the same objects) -- but `Solver.solve` never calls them: it ships their parameters to the
device.  A plain black-box `Expr` inside a constraint makes `Solver.solve` raise (there is no CPU
fallback).

Host-side numerical derivatives (only reached through `Expr.grad` / `Expr.hess` of a black box
without analytic derivatives) use the scheme of the device FK7 kernel: central differences at h and
2h, Richardson-combined (expr.py:61-69,102-109 call numdifftools with its defaults).
The rounded-x caches of the reference (expr.py:13,31-41: quirk C-6 of SURVEY.md) are not replicated.
"""
import numpy as np

from . import families_host, sym
from .structure import FAM_CIRCLE2D, FAM_FK7, FAM_QUADFORM, FAM_VM

DEFAULT_TOL = 1e-4


def _central_jacobian(fun, x):
    """(m, n) Jacobian of fun: R^n -> R^m at flat x; Richardson combination of two central differences."""
    x = np.asarray(x, dtype=float).ravel()
    f0 = np.atleast_1d(np.asarray(fun(x), dtype=float)).ravel()
    J = np.zeros((f0.size, x.size))
    for j in range(x.size):
        h = families_host.FD_BASE_STEP * max(np.log1p(abs(x[j])), 1.0)
        est = []
        for hh in (h, 2.0 * h):
            xp, xm = x.copy(), x.copy()
            xp[j] += hh
            xm[j] -= hh
            est.append((np.ravel(fun(xp)) - np.ravel(fun(xm))) / (xp[j] - xm[j]))
        J[:, j] = (4.0 * est[0] - est[1]) / 3.0
    return J


HESS_BASE_STEP = np.finfo(float).eps ** (1.0 / 7.8)  # numdifftools default for second derivatives


def _central_hessian(fun, x):
    """(n, n) Hessian of a scalar fun at flat x: central second differences at h and 2h,
    Richardson-combined -- the scheme of the device kernel (sco_families.cuh vm_fd2)."""
    x = np.asarray(x, dtype=float).ravel()
    n = x.size
    f = lambda v: float(np.ravel(fun(v))[0])
    f0 = f(x)
    h0 = HESS_BASE_STEP * np.maximum(np.log1p(np.abs(x)), 1.0)
    est = []
    for mult in (1.0, 2.0):
        h = h0 * mult
        H = np.zeros((n, n))
        for i in range(n):
            ei = np.zeros(n)
            ei[i] = h[i]
            H[i, i] = (f(x + 2 * ei) - 2.0 * f0 + f(x - 2 * ei)) / (4.0 * h[i] * h[i])
            for j in range(i + 1, n):
                ej = np.zeros(n)
                ej[j] = h[j]
                H[i, j] = H[j, i] = (f(x + ei + ej) - f(x + ei - ej) - f(x - ei + ej) + f(x - ei - ej)) / (4.0 * h[i] * h[j])
        est.append(H)
    return (4.0 * est[0] - est[1]) / 3.0


class Expr(object):
    """Black-box expression f with optional analytic gradient / Hessian (expr.py:22-156)."""

    family = None  # device families override

    def __init__(self, f, grad=None, hess=None, **kwargs):
        self.f = f
        self._grad = grad
        self._hess = hess

    def eval(self, x):
        return self.f(x)

    def _flat(self, x):
        shape = np.shape(x)
        return lambda v: np.ravel(self.f(np.reshape(v, shape)))

    def grad(self, x, num_check=False, atol=DEFAULT_TOL):
        if self._grad is None:
            assert not num_check
            return _central_jacobian(self._flat(x), x)
        g = self._grad(x)
        if num_check and not np.allclose(_central_jacobian(self._flat(x), x), g, atol=atol):
            raise Exception("Numerical and analytical gradients aren't close.")
        return g

    def hess(self, x, num_check=False, atol=DEFAULT_TOL):
        if self._hess is not None:
            h = self._hess(x)
            if num_check and not np.allclose(self._num_hess(x), h, atol=atol):
                raise Exception("Numerical and analytical hessians aren't close.")
            return h
        assert not num_check
        return self._num_hess(x)

    def _num_hess(self, x):
        return _central_hessian(self._flat(x), x)

    def convexify(self, x, degree=1):
        """Affine (degree 1) or convex quadratic (degree 2) model at x (expr.py:130-156)."""
        if degree == 1:
            A = self.grad(x)
            return AffExpr(A, self.eval(x) - A.dot(x))
        if degree == 2:
            H = self.hess(x)
            lam = np.linalg.eigvalsh(H).min()
            if lam < 0:
                H = H - lam * np.eye(H.shape[0])
            g = self.grad(x)
            return QuadExpr(H, g - x.T.dot(H), 0.5 * x.T.dot(H).dot(x) - g.dot(x) + self.eval(x))
        raise NotImplementedError


class AffExpr(Expr):
    """A x + b (expr.py:159-181; note grad returns A' as the reference does)."""

    def __init__(self, A, b):
        assert b.shape[0] == A.shape[0]
        self.A = A
        self.b = b
        self.x_shape = (A.shape[1], 1)

    def eval(self, x):
        return self.A.dot(x) + self.b

    def grad(self, x):
        return self.A.T

    def hess(self, x):
        return np.zeros((self.x_shape[0],) * 2)


class QuadExpr(Expr):
    """0.5 x'Qx + A x + b, scalar (expr.py:184-213)."""

    def __init__(self, Q, A, b):
        assert A.shape[0] == 1, "Can only define scalar quadrative expressions"
        assert Q.shape[0] == Q.shape[1] == A.shape[1]
        assert b.shape[0] == 1
        self.Q = Q
        self.A = A
        self.b = b
        self.x_shape = (A.shape[1], 1)

    def eval(self, x):
        return 0.5 * x.T.dot(self.Q.dot(x)) + self.A.dot(x) + self.b

    def grad(self, x):
        assert x.shape == self.x_shape
        return 0.5 * (self.Q + self.Q.T).dot(x) + self.A.T

    def hess(self, x):
        return self.Q.copy()


class AbsExpr(Expr):
    """|expr| (expr.py:216-235)."""

    def __init__(self, expr):
        self.expr = expr

    def eval(self, x):
        return np.absolute(self.expr.eval(x))

    def grad(self, x):
        raise NotImplementedError

    def hess(self, x):
        raise NotImplementedError


class HingeExpr(Expr):
    """max(expr, 0) (expr.py:238-259)."""

    def __init__(self, expr):
        self.expr = expr

    def eval(self, x):
        return np.maximum(self.expr.eval(x), 0.0)

    def grad(self, x):
        raise NotImplementedError

    def hess(self, x):
        raise NotImplementedError


class CompExpr(Expr):
    """expr compared with val (expr.py:262-297)."""

    def __init__(self, expr, val):
        self.expr = expr
        self.val = val.copy()

    def eval(self, x, tol=DEFAULT_TOL):
        raise NotImplementedError

    def grad(self, x):
        raise Exception("The gradient is not well defined for comparison expressions")

    def hess(self, x):
        raise Exception("The hessian is not well defined for comparison expressions")

    def convexify(self, x, degree=1):
        raise NotImplementedError

    def _model(self, x):
        aff = self.expr.convexify(x, degree=1)
        aff.b = aff.b - self.val
        return aff


class EqExpr(CompExpr):
    """expr == val; its convex model is the l1 penalty |A x + b - val| (expr.py:300-332)."""

    def eval(self, x, tol=DEFAULT_TOL, negated=False):
        assert tol >= 0.0
        ok = np.allclose(self.expr.eval(x), self.val, atol=tol)
        return (not ok) if negated else ok

    def convexify(self, x, degree=1):
        assert degree == 1
        return AbsExpr(self._model(x))


class LEqExpr(CompExpr):
    """expr <= val; its convex model is the hinge penalty max(A x + b - val, 0) (expr.py:335-371)."""

    def eval(self, x, tol=DEFAULT_TOL, negated=False):
        assert tol >= 0.0
        v = self.expr.eval(x)
        if negated:
            return not np.all(v <= self.val - tol)
        return np.all(v <= self.val + tol)

    def convexify(self, x, degree=1):
        assert degree == 1
        return HingeExpr(self._model(x))


class LExpr(CompExpr):
    """expr < val (expr.py:374-410).  The OSQP backend of the reference cannot use it (affine LExpr
    constraints are silently dropped, nonlinear ones crash: quirk C-9), so `Prob.add_cnt_expr` of this
    package rejects it."""

    def eval(self, x, tol=DEFAULT_TOL, negated=False):
        assert tol >= 0.0
        v = self.expr.eval(x)
        if negated:
            return not np.all(v < self.val - tol)
        return np.all(v < self.val + tol)

    def convexify(self, x, degree=1):
        assert degree == 1
        return HingeExpr(self._model(x))


class BoundExpr(object):
    """An expression bound to a Variable; the variable's ordering matters (expr.py:413-437)."""

    def __init__(self, expr, var):
        self.expr = expr
        self.var = var

    def eval(self):
        return self.expr.eval(self.var.get_value())

    def convexify(self, degree=1):
        assert self.var.get_value() is not None
        return BoundExpr(self.expr.convexify(self.var.get_value(), degree), self.var)


# ------------------------------------------------------------------------------ device families
class DeviceFamilyExpr(Expr):
    """An `Expr` whose f / grad are one of the closed families the kernels of sco_families.cuh
    evaluate.  Subclasses define `family`, `m`, `jw`, `ipar`, `params()` and `shared_params`."""

    shared_params = False  # True: the parameters are the same for every problem of a batch

    def __init__(self):
        super().__init__(self._f, self._g)

    def params(self):
        raise NotImplementedError


class QuadFormExpr(DeviceFamilyExpr):
    """f_j(x) = 0.5 x'P_j x + a_j'x, j < m, over ALL n variables of the problem (the form
    BASELINE.json's QCQP config writes as Expr(f, grad); `QuadExpr` itself cannot be a constraint for
    n > 1, quirk C-8)."""

    family = FAM_QUADFORM

    def __init__(self, P, a):
        P = np.asarray(P, dtype=float)
        a = np.asarray(a, dtype=float)
        assert P.ndim == 3 and P.shape[1] == P.shape[2] and a.shape == P.shape[:2]
        self.P = 0.5 * (P + np.transpose(P, (0, 2, 1)))  # only the symmetric part enters x'Px
        self.a = a
        self.m, self.n = a.shape
        self.jw = self.n
        self.ipar = [self.n, self.m, 0, 0, 0, 0, 0, 0]
        super().__init__()

    def _f(self, x):
        v = x[:, 0]
        return (0.5 * np.einsum("i,jik,k->j", v, self.P, v) + self.a @ v).reshape(-1, 1)

    def _g(self, x):
        return self.P @ x[:, 0] + self.a

    def params(self):
        iu = np.triu_indices(self.n)
        return np.concatenate([self.P[:, iu[0], iu[1]].ravel(), self.a.ravel()])


class CircleDistExpr(DeviceFamilyExpr):
    """f_{t,k}(x) = R_k - |p_t - c_k| for the T 2-D way-points p_t = x[2t:2t+2] and K discs: the
    obstacle-avoidance constraint of the point-robot config.  Row t*K + k."""

    family = FAM_CIRCLE2D

    def __init__(self, T, centres, radii):
        self.T = int(T)
        self.centres = np.asarray(centres, dtype=float).reshape(-1, 2)
        self.radii = np.asarray(radii, dtype=float).ravel()
        self.K = self.radii.size
        assert self.centres.shape[0] == self.K
        self.m = self.T * self.K
        self.jw = 2
        self.ipar = [self.T, self.K, 0, 0, 0, 0, 0, 0]
        super().__init__()

    def _diff(self, x):
        p = x[:2 * self.T, 0].reshape(self.T, 1, 2)
        return p - self.centres[None, :, :]

    def _f(self, x):
        d = np.sqrt((self._diff(x) ** 2).sum(axis=2))
        return (self.radii[None, :] - d).reshape(-1, 1)

    def _g(self, x):
        diff = self._diff(x)
        d = np.sqrt((diff ** 2).sum(axis=2))
        J = np.zeros((self.m, x.shape[0]))
        for t in range(self.T):
            for k in range(self.K):
                J[t * self.K + k, 2 * t:2 * t + 2] = -diff[t, k] / d[t, k]
        return J

    def params(self):
        return np.concatenate([self.centres.ravel(), self.radii])


class FK7Expr(DeviceFamilyExpr):
    """Flange position (3 rows) of the 7-link modified-DH chain of families_host on the LAST seven
    variables of an n-vector; no analytic gradient -- the Jacobian is finite-differenced, on the device
    by the FK7 kernel and on the host by `Expr.grad` (the arm config of BASELINE.json)."""

    family = FAM_FK7
    shared_params = True

    def __init__(self, n):
        self.n = int(n)
        self.m = 3
        self.jw = 7
        self.ipar = [self.n // 7, 0, 0, 0, 0, 0, 0, 0]
        Expr.__init__(self, self._f, None)

    def _f(self, x):
        return families_host.fk7_pos(np.asarray(x).reshape(-1)[-7:]).reshape(3, 1)

    def params(self):
        return families_host.fk7_table()


class SymExpr(DeviceFamilyExpr):
    """Rows written in the expression language of sco_py_b200.sym, over ALL n variables.

    analytic=False (default): no analytic derivatives -- Jacobians (constraints) and gradient + Hessian (a
    non-quadratic objective, expr.py:102-156) are finite-differenced with the numdifftools scheme, on the device by the
    VM family and on the host by `Expr.grad` / `Expr.hess`: exactly what the reference does with `Expr(f)`.
    analytic=True: the counterpart of `Expr(f, grad)` (expr.py:86-88): the program is differentiated in forward mode,
    on the device (vm_eval_dual) and on the host (sym.jacobian); a non-quadratic objective keeps its numerical
    Hessian, as in the reference when only `grad` is given (expr.py:102-109)."""

    family = FAM_VM

    def __init__(self, rows, n, analytic=False):
        self.n = int(n)
        self.rows = list(rows)
        self.m = len(self.rows)
        self.analytic = bool(analytic)
        self.program, self.n_instr = sym.compile_rows(self.rows)
        ins = self.program[self.m:].reshape(-1, 2)
        used = ins[ins[:, 0] == sym.PUSH_X, 1]
        if used.size and (used.min() < 0 or used.max() >= self.n):  # the device would read past x
            raise ValueError("SymExpr over %d variables uses variable index %d" % (self.n, int(used.max())))
        self.jw = self.n
        self.ipar = [self.n, self.m, self.n_instr, int(self.analytic), 0, 0, 0, 0]
        Expr.__init__(self, self._f, self._g if self.analytic else None)

    def _f(self, x):
        return sym.eval_program(self.program, self.m, np.asarray(x).ravel()).reshape(self.m, 1)

    def _g(self, x):
        return sym.jacobian(self.program, self.m, self.n, np.asarray(x).ravel())

    def params(self):
        return self.program
