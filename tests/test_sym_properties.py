"""Property tests of the expression language (sco_py_b200/sym.py) on random expression DAGs:
  * the compiled stack program (with temporaries for shared sub-expressions) evaluates to what the same
    tree gives when it is evaluated directly in Python, bit for bit;
  * the oracle's independent interpreter (oracle/families.py) agrees bit for bit, values and derivatives;
  * forward-mode derivatives agree with central differences away from the kinks of abs / min / max.
These are the host-side halves of the device VM's parity tests (tests/test_c1_toy.py, test_geometry.py)."""
import math

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as hst

import families  # oracle/families.py
from sco_py_b200 import sym

N = 3


class Node(object):
    """A random expression: builds the Sym tree and evaluates itself in plain Python, in the same operation order."""

    def __init__(self, kind, args=(), value=None):
        self.kind, self.args, self.value = kind, args, value
        self._sym = None

    def sym(self, X):  # one Sym per Node: a Node used twice becomes a shared sub-expression
        if self._sym is None:
            a = [c.sym(X) for c in self.args]
            k = self.kind
            self._sym = (X[self.value] if k == "x" else sym.Sym.wrap(self.value) if k == "c" else
                         a[0] + a[1] if k == "+" else a[0] - a[1] if k == "-" else a[0] * a[1] if k == "*" else
                         a[0] / (2.0 + a[1] * a[1]) if k == "/" else -a[0] if k == "neg" else
                         a[0] ** self.value if k == "pow" else sym.sqrt(a[0] * a[0] + 1.0) if k == "sqrt" else
                         sym.log(1.0 + a[0] * a[0]) if k == "log" else sym.exp(0.1 * sym.sin(a[0])) if k == "exp" else
                         sym.sin(a[0]) if k == "sin" else sym.cos(a[0]) if k == "cos" else abs(a[0]) if k == "abs" else
                         sym.minimum(a[0], a[1]) if k == "min" else sym.maximum(a[0], a[1]))
        return self._sym

    def ev(self, v):
        a = [c.ev(v) for c in self.args]
        k = self.kind
        if k == "x": return float(v[self.value])
        if k == "c": return float(self.value)
        if k == "+": return a[0] + a[1]
        if k == "-": return a[0] - a[1]
        if k == "*": return a[0] * a[1]
        if k == "/": return a[0] / (2.0 + a[1] * a[1])
        if k == "neg": return -a[0]
        if k == "pow":
            r = 1.0
            for _ in range(self.value):  # repeated products, left to right (sym.POWI)
                r = r * a[0]
            return r
        if k == "sqrt": return math.sqrt(a[0] * a[0] + 1.0)
        if k == "log": return math.log(1.0 + a[0] * a[0])
        if k == "exp": return math.exp(0.1 * math.sin(a[0]))
        if k == "sin": return math.sin(a[0])
        if k == "cos": return math.cos(a[0])
        if k == "abs": return abs(a[0])
        if k == "min": return a[0] if a[0] < a[1] else a[1]
        return a[0] if a[0] > a[1] else a[1]

    def smooth(self):
        return self.kind not in ("abs", "min", "max") and all(c.smooth() for c in self.args)


leaves = hst.one_of(hst.integers(0, N - 1).map(lambda j: Node("x", value=j)),
                    hst.floats(-2.0, 2.0, allow_nan=False).map(lambda c: Node("c", value=round(c, 3))))


def grow(children):
    un = hst.tuples(hst.sampled_from(["neg", "sqrt", "log", "exp", "sin", "cos", "abs"]), children).map(
        lambda t: Node(t[0], (t[1],)))
    pw = hst.tuples(hst.integers(0, 3), children).map(lambda t: Node("pow", (t[1],), value=t[0]))
    bi = hst.tuples(hst.sampled_from(["+", "-", "*", "/", "min", "max"]), children, children).map(
        lambda t: Node(t[0], (t[1], t[2])))
    sq = children.map(lambda c: Node("*", (c, c)))  # the same node twice: a shared sub-expression (TEE / LOAD)
    return hst.one_of(un, pw, bi, sq)


exprs = hst.recursive(leaves, grow, max_leaves=12)
points = hst.lists(hst.floats(-1.5, 1.5, allow_nan=False), min_size=N, max_size=N)
SET = dict(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))


@settings(**SET)
@given(hst.lists(exprs, min_size=1, max_size=3), points)
def test_compiled_programs_evaluate_like_the_tree(rows, pt):
    X = sym.variables(N)
    for r in rows:
        for node in _all(r):
            node._sym = None
    try:
        prog, n_instr = sym.compile_rows([r.sym(X) for r in rows])
    except ValueError:  # deeper than the VM's stack / more temporaries than its slots: refused, not mis-compiled
        return
    v = np.array(pt)
    ref = np.array([r.ev(v) for r in rows])
    got = sym.eval_program(prog, len(rows), v)
    assert np.array_equal(got, ref), (got, ref)
    assert np.array_equal(families.vm_f(v.reshape(N, 1), prog, len(rows))[:, 0], got)
    J = sym.jacobian(prog, len(rows), N, v)
    assert np.array_equal(families.vm_grad(v.reshape(N, 1), prog, len(rows)), J)


@settings(**SET)
@given(exprs, points)
def test_forward_mode_matches_central_differences_on_smooth_programs(row, pt):
    if not row.smooth():
        return
    X = sym.variables(N)
    for node in _all(row):
        node._sym = None
    try:
        prog, _ = sym.compile_rows([row.sym(X)])
    except ValueError:
        return
    v = np.array(pt)
    J = sym.jacobian(prog, 1, N, v)[0]
    h = 1e-6
    for j in range(N):
        e = np.zeros(N)
        e[j] = h
        fd = (sym.eval_program(prog, 1, v + e)[0] - sym.eval_program(prog, 1, v - e)[0]) / (2 * h)
        assert abs(J[j] - fd) <= 1e-6 * max(1.0, abs(fd), abs(J[j])) + 1e-7, (j, J[j], fd)


def _all(node, seen=None):
    seen = {} if seen is None else seen
    if id(node) not in seen:
        seen[id(node)] = node
        for c in node.args:
            _all(c, seen)
    return list(seen.values())


# ------------------------------------------------------------------ the device VM on the same random programs
import pytest  # noqa: E402


def _problem(rows_sym, pt, analytic):
    from sco_py_b200.expr import BoundExpr, LEqExpr, QuadExpr, SymExpr
    from sco_py_b200.sco_b200.osqp_utils import OSQPVar
    from sco_py_b200.sco_b200.prob import Prob
    from sco_py_b200.sco_b200.variable import Variable
    prob = Prob()
    ov = np.array([[OSQPVar("x%d" % j)] for j in range(N)], dtype=object)
    for v in ov[:, 0]:
        prob.add_osqp_var(v)
    var = Variable(ov, value=np.asarray(pt, dtype=float).reshape(N, 1))
    prob.add_var(var)
    prob.add_obj_expr(BoundExpr(QuadExpr(np.eye(N), np.zeros((1, N)), np.zeros((1, 1))), var))
    prob.add_cnt_expr(BoundExpr(LEqExpr(SymExpr(rows_sym, N, analytic=analytic), np.zeros((len(rows_sym), 1))), var))
    return prob


@pytest.mark.gpu
@settings(max_examples=10, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(hst.lists(exprs, min_size=2, max_size=2), hst.lists(points, min_size=4, max_size=4))
def test_device_vm_evaluates_random_programs(rows, pts):
    """Values to 1e-10 and forward-mode Jacobians to 1e-8 of random expression DAGs (temporaries, abs / min / max, integer
    powers) on the device against the host interpreter; affine offsets b = f - J x of the convexification."""
    from sco_py_b200 import batch
    from sco_py_b200.engine import Engine
    X = sym.variables(N)
    for r in rows:
        for node in _all(r):
            node._sym = None
    try:
        rows_sym = [r.sym(X) for r in rows]
        prog, _ = sym.compile_rows(rows_sym)
    except ValueError:
        return
    st, params, x0, _ = batch.compile_batch([_problem(rows_sym, p, True) for p in pts])
    eng = Engine(st)
    f, J, b, _ = [t.cpu().numpy() for t in eng.convexify(params, x0)]
    eng.close()
    for i, p in enumerate(pts):
        v = np.array(p)
        fh = sym.eval_program(prog, 2, v)
        Jh = sym.jacobian(prog, 2, N, v)
        scale = max(1.0, np.abs(fh).max())
        # device libm differs from the host's in the last ulp or two; cancellation inside a random expression amplifies it
        assert np.abs(f[i] - fh).max() <= 1e-10 * scale, (i, f[i], fh)
        assert np.abs(J[i].reshape(2, N) - Jh).max() <= 1e-8 * max(1.0, np.abs(Jh).max()), (i, J[i], Jh)
        assert np.abs(b[i] - (fh - Jh @ v)).max() <= 1e-8 * max(scale, np.abs(Jh).max())
