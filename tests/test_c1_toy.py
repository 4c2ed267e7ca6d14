"""BASELINE.json configs[0]: the toy non-convex problems of the reference's own solver test
(tests/sco_osqp/test_solver.py:91-169, hyper-parameters :15-25, pass mark atol 5e-4 on the optimum :87).

Five of the nine have objective and constraints that are quadratic forms (prob0, 2, 3, 5, 6); they are
written here with the device families (QuadExpr objective, QuadFormExpr constraints -- a linear g is a
quadratic form with P = 0) exactly as helper_test_prob (:32-87) assembles them: one quadratic
objective, one linear LEq row block, one nonlinear LEq block g, one nonlinear Eq block h, where absent
pieces are the reference's placeholders (g = -1e5, h = 0).  The other four need a black-box objective
(quartic, Rosenbrock, log): finite-difference Hessians are not on the device yet.
"""
import numpy as np
import pytest

import sqp_port
from sco_py_b200 import batch
from sco_py_b200 import workloads as W
from sco_py_b200.expr import AffExpr, BoundExpr, EqExpr, LEqExpr, QuadExpr, QuadFormExpr
from sco_py_b200.sco_b200.osqp_utils import OSQPVar
from sco_py_b200.sco_b200.prob import Prob
from sco_py_b200.sco_b200.solver import Solver
from sco_py_b200.sco_b200.variable import Variable

N = 2
ANG = (np.arange(1, 7) * 2 * np.pi / 6).reshape(6, 1)
HEX_A = np.hstack((np.cos(ANG), np.sin(ANG)))
HEX_q = -np.array([[np.cos(np.pi / 6), np.sin(np.pi / 6)]])


def quadform(rows):
    """rows: list of (P 2x2, a 2, const) -> (QuadFormExpr, val) with f_j = 0.5 x'P_j x + a_j'x <= / == -const."""
    P = np.array([r[0] for r in rows], dtype=float)
    a = np.array([r[1] for r in rows], dtype=float)
    val = -np.array([[r[2]] for r in rows], dtype=float)
    return QuadFormExpr(P, a), val


Z = np.zeros((2, 2))
PROBLEMS = {
    # name: (x0, x_true, Q, q, c, A_ineq, b_ineq, g rows, h rows)
    "prob0": ([1.0, 1.0], [1.5, 1.5], 2 * np.eye(2), np.zeros((1, 2)), 0.0, None, None,
              [(Z, [-1.0, -1.0], 3.0)], None),                                   # g = 3 - x1 - x2
    "prob2": ([10.0, 1.0], [0.0, 0.0], 2 * np.array([[1.0, -1.0], [-1.0, 1.0]]), np.array([[0.0, 1.0]]), 1e-5,
              None, None, [(Z, [0.0, -1.0], 0.0)], None),                        # f = x2 + 1e-5 + (x2-x1)^2 ; g = -x2
    "prob3": ([10.0, 1.0], [1.0, 1.0], 2 * np.diag([1.0, 0.0]), np.array([[-2.0, 0.0]]), 1.0, None, None, None,
              [(np.array([[-20.0, 0.0], [0.0, 0.0]]), [0.0, 10.0], 0.0)]),       # h = 10 (x2 - x1^2)
    "prob5": ([0.0, 0.0], [1.0, np.tan(np.pi / 6)], Z, HEX_q, 0.0, HEX_A, np.ones((6, 1)), None, None),
    "prob6": ([0.0, 0.0], [1.0, np.tan(np.pi / 6)], 0.1 * np.eye(2), HEX_q, 0.0, None, None,
              [(Z, 0.01 * HEX_A[k], -0.01) for k in range(6)], None),            # g = 0.01 (A x - 1)
}


def build(name, x0=None):
    x0_, x_true, Q, q, c, A_ineq, b_ineq, g, h = PROBLEMS[name]
    x0 = np.array(x0_ if x0 is None else x0, dtype=float).reshape(2, 1)
    prob = Prob()
    ov = np.array([[OSQPVar("x1")], [OSQPVar("x2")]], dtype=object)
    for v in ov[:, 0]:
        prob.add_osqp_var(v)
    var = Variable(ov, value=x0)
    prob.add_var(var)
    prob.add_obj_expr(BoundExpr(QuadExpr(Q, q, np.array([[c]])), var))
    if A_ineq is None:
        A_ineq, b_ineq = np.zeros((1, N)), np.zeros((1, 1))
    prob.add_cnt_expr(BoundExpr(LEqExpr(AffExpr(A_ineq, -b_ineq), np.zeros(b_ineq.shape)), var))
    ge, gval = quadform(g if g is not None else [(Z, [0.0, 0.0], -1e5)])       # neginffunc
    prob.add_cnt_expr(BoundExpr(LEqExpr(ge, gval), var))
    he, hval = quadform(h if h is not None else [(Z, [0.0, 0.0], 0.0)])        # zerofunc
    prob.add_cnt_expr(BoundExpr(EqExpr(he, hval), var))
    return prob, var, np.array(x_true)


def _solver():
    s = Solver()
    for k, v in W.SOLVER_SETTINGS.items():
        setattr(s, k, v)
    return s


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_oracle_reaches_the_reference_answers(name):
    """The modelling above is the reference's problem: the oracle port lands on x_true at atol 5e-4."""
    prob, var, x_true = build(name)
    st, params, x0, _ = batch.compile_batch([prob])
    r = sqp_port.solve(st, params[0], x0[0], solver=W.SOLVER_SETTINGS)
    assert np.allclose(r["x"], x_true, atol=5e-4), (name, r["x"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_device_reaches_the_reference_answers(name):
    prob, var, x_true = build(name)
    _solver().solve(prob, method="penalty_sqp")
    assert np.allclose(var.get_value()[:, 0], x_true, atol=5e-4), (name, var.get_value()[:, 0])


@pytest.mark.gpu
def test_jittered_batch_of_toy_problems():
    """B copies with x0 jittered by N(0, 0.1^2), seed 1 -- one launch, same optimum."""
    rng = np.random.default_rng(1)
    for name in ("prob0", "prob3"):
        built = [build(name, x0=np.array(PROBLEMS[name][0]) + rng.normal(0.0, 0.1, 2)) for _ in range(16)]
        _solver().solve_batch([b[0] for b in built], method="penalty_sqp")
        for prob, var, x_true in built:
            assert np.allclose(var.get_value()[:, 0], x_true, atol=5e-4), (name, var.get_value()[:, 0])
