"""The reference-facing API layer (sco_py_b200.expr + sco_py_b200.sco_b200): host logic on CPU,
the device path on the GPU box."""
import numpy as np
import pytest

import api_builder
import sqp_port
from sco_py_b200 import batch
from sco_py_b200 import expr as E
from sco_py_b200 import workloads as W
from sco_py_b200.sco_b200.osqp_utils import OSQPVar
from sco_py_b200.sco_b200.prob import Prob
from sco_py_b200.sco_b200.solver import Solver
from sco_py_b200.sco_b200.variable import Variable


def _solver():
    s = Solver()
    for k, v in W.SOLVER_SETTINGS.items():
        setattr(s, k, v)
    return s


# ------------------------------------------------------------------ CPU: compiler and host classes
@pytest.mark.parametrize("name,B", [("qcqp", 3), ("point_robot", 2), ("arm", 2)])
def test_compiled_batch_is_the_same_problem(name, B):
    """API objects -> compile_batch gives a structure on which the oracle reproduces, bit for bit,
    what it computes on the workload's own structure."""
    st, params, x0 = W.GENERATORS[name](B)
    probs = [api_builder.build_prob(st, params[i], x0[i])[0] for i in range(B)]
    st2, p2, x2, _ = batch.compile_batch(probs)
    assert (st2.n, st2.m_lin, st2.m_nl, st2.n_slack) == (st.n, st.m_lin, st.m_nl, st.n_slack)
    assert np.array_equal(x2, x0)
    for i in range(B):
        a = sqp_port.PortProblem(st, params[i], x0[i])
        b = sqp_port.PortProblem(st2, p2[i], x2[i])
        assert np.array_equal(a.Q, b.Q) and np.array_equal(a.q, b.q)
        a.convexify()
        b.convexify()
        for bi in range(len(st.blocks)):
            assert np.array_equal(a.J[bi], b.J[bi]) and np.array_equal(a.b[bi], b.b[bi])
        if st.m_lin:
            assert np.array_equal(a.A_lin.toarray(), b.A_lin.toarray())
            assert np.array_equal(a.l_lin, b.l_lin) and np.array_equal(a.u_lin, b.u_lin)
    if name != "qcqp":  # the smoothness objective and the FK table are shared by the whole batch
        assert st2.Q.shared


def test_variable_ordering_follows_the_names():
    """QP column order = stable sort by var_name (osqp_utils.py:136-143), whatever the insertion order."""
    prob = Prob()
    names = ["b", "a", "c"]
    ov = np.array([[OSQPVar(nm)] for nm in names], dtype=object)
    for v in ov[::-1, 0]:
        prob.add_osqp_var(v)
    var = Variable(ov, np.array([[2.0], [1.0], [3.0]]))
    prob.add_obj_expr(E.BoundExpr(E.QuadExpr(np.diag([20.0, 10.0, 30.0]), np.zeros((1, 3)), np.zeros((1, 1))), var))
    cp = batch.compile_problem(prob)
    assert [v.var_name for v in cp.ovars] == ["a", "b", "c"]
    assert np.array_equal(cp.x0, [1.0, 2.0, 3.0])
    assert np.array_equal(np.diag(cp.Q), [10.0, 20.0, 30.0])
    batch.scatter_solution(cp, np.array([7.0, 8.0, 9.0]))
    assert np.array_equal(var.get_value()[:, 0], [8.0, 7.0, 9.0]) and ov[0, 0].val == 8.0


def test_unsupported_inputs_fail_loudly():
    ov = np.array([[OSQPVar("x0")], [OSQPVar("x1")]], dtype=object)
    var = Variable(ov, np.zeros((2, 1)))

    def fresh():
        p = Prob()
        for v in ov[:, 0]:
            p.add_osqp_var(v)
        p.add_obj_expr(E.BoundExpr(E.QuadExpr(np.eye(2), np.zeros((1, 2)), np.zeros((1, 1))), var))
        return p

    p = fresh()
    p.add_cnt_expr(E.BoundExpr(E.LEqExpr(E.Expr(lambda x: x[:1] ** 2), np.zeros((1, 1))), var))
    with pytest.raises(batch.UnsupportedProblem, match="black-box"):
        batch.compile_batch([p])
    p = fresh()  # AffExpr objective terms are accepted (quirk C-4 weight on the device): their own field
    p.add_obj_expr(E.BoundExpr(E.AffExpr(np.array([[1.0, 2.0], [0.5, 0.0]]), np.array([[0.25], [0.5]])), var))
    st, params, _, _ = batch.compile_batch([p])
    assert st.qa.off >= 0 and np.array_equal(params[0, st.qa.off:st.qa.off + 2], [1.5, 2.0])
    assert params[0, st.c.off] == 0.75
    p = fresh()
    p.add_cnt_expr(E.BoundExpr(E.LEqExpr(E.QuadFormExpr(np.zeros((1, 3, 3)), np.ones((1, 3))), np.zeros((1, 1))), var))
    with pytest.raises(batch.UnsupportedProblem, match="3 variables"):
        batch.compile_batch([p])
    with pytest.raises(NotImplementedError):
        fresh().add_cnt_expr(E.BoundExpr(E.LExpr(E.AffExpr(np.ones((1, 2)), np.zeros((1, 1))), np.ones((1, 1))), var))
    with pytest.raises(Exception, match="This method is not supported."):
        Solver().solve(fresh(), method="nope")
    with pytest.raises(ValueError):
        Variable(np.array([[OSQPVar("y")]], dtype=object)).update()


def test_expr_classes_behave_like_the_reference_ones():
    """The checks of tests/sco_osqp/test_expr.py that pin shapes and values."""
    x = np.array([[1.0], [2.0]])
    aff = E.AffExpr(np.array([[1.0, 2.0], [3.0, 4.0]]), np.array([[1.0], [1.0]]))
    assert np.array_equal(aff.eval(x), [[6.0], [12.0]]) and np.array_equal(aff.grad(x), aff.A.T)
    quad = E.QuadExpr(np.array([[2.0, 0.0], [0.0, 4.0]]), np.array([[1.0, 1.0]]), np.array([[3.0]]))
    assert quad.eval(x)[0, 0] == 0.5 * (2 + 16) + 3 + 3
    assert np.array_equal(quad.grad(x), [[3.0], [9.0]]) and np.array_equal(quad.hess(x), quad.Q)
    f = E.Expr(lambda v: v ** 3)
    for pt in (1.0, -1.0, 2.0, 0.0):
        assert np.allclose(f.grad(np.array([[pt]])), 3 * pt ** 2)
    cvx = E.Expr(lambda v: -v ** 2).convexify(np.array([[1.0]]), degree=2)  # negative curvature clipped
    assert np.allclose(cvx.Q, 0.0)
    le = E.LEqExpr(aff, np.array([[6.0], [12.0]]))
    assert le.eval(x) and not le.eval(x + 1e-2)
    hinge = le.convexify(x)
    assert isinstance(hinge, E.HingeExpr) and np.allclose(hinge.eval(x), 0.0)
    eq = E.EqExpr(aff, np.array([[6.0], [12.0]]))
    assert eq.eval(x) and isinstance(eq.convexify(x), E.AbsExpr)
    fam = E.QuadFormExpr(np.random.default_rng(0).standard_normal((3, 2, 2)), np.ones((3, 2)))
    assert np.allclose(fam.grad(x), E.Expr(fam.f).grad(x), atol=1e-6)


# ------------------------------------------------------------------ GPU: the drop-in path
@pytest.mark.gpu
@pytest.mark.parametrize("name,B,tol", [("qcqp", 4, 1e-4), ("point_robot", 2, 1e-4)])
def test_solver_solve_batch_matches_the_oracle(name, B, tol):
    st, params, x0 = W.GENERATORS[name](B)
    built = [api_builder.build_prob(st, params[i], x0[i]) for i in range(B)]
    ok = _solver().solve_batch([p for p, _ in built], method="penalty_sqp")
    for i, (prob, var) in enumerate(built):
        ref = sqp_port.solve(st, params[i], x0[i], solver=W.SOLVER_SETTINGS)
        assert ok[i] == ref["success"]
        x = var.get_value()[:, 0]
        assert np.abs(x - ref["x"]).max() <= tol * max(1.0, np.abs(ref["x"]).max())
        assert abs(prob.get_max_cnt_violation() - ref["max_vio"]) <= 1e-5


@pytest.mark.gpu
def test_solver_solve_single_problem_and_settings():
    st, params, x0 = W.gen_qcqp(1, n=8, m=6)
    prob, var = api_builder.build_prob(st, params[0], x0[0])
    assert _solver().solve(prob, method="penalty_sqp") is True
    ref = sqp_port.solve(st, params[0], x0[0], solver=W.SOLVER_SETTINGS)
    assert np.abs(var.get_value()[:, 0] - ref["x"]).max() <= 1e-4


@pytest.mark.gpu
def test_prob_stage_methods_known_answers():
    """The step-by-step surface on known answers of tests/sco_osqp/test_prob.py: the projection of
    (1, 1)... onto linear constraints (:48-75) and a QP optimum under a trust region."""
    ov = np.array([[OSQPVar("x0")], [OSQPVar("x1")]], dtype=object)

    def problem(value, cnts):
        p = Prob()
        for v in ov[:, 0]:
            p.add_osqp_var(v)
            v.set_lower_bound(-np.inf)
            v.set_upper_bound(np.inf)
        var = Variable(ov, np.array(value, dtype=float).reshape(2, 1))
        p.add_var(var)
        for c in cnts:
            p.add_cnt_expr(E.BoundExpr(c, var))
        return p, var

    # x <= 0 componentwise, from (1, 1): closest feasible point is (0, 0)
    p, var = problem([1.0, 1.0], [E.LEqExpr(E.AffExpr(np.eye(2), np.zeros((2, 1))), np.zeros((2, 1)))])
    assert p.find_closest_feasible_point()
    assert np.allclose(var.get_value(), 0.0, atol=1e-5)
    # x0 == -1, x1 <= 0 from (1, -1): (-1, -1)
    p, var = problem([1.0, -1.0], [E.EqExpr(E.AffExpr(np.array([[1.0, 0.0]]), np.zeros((1, 1))), -np.ones((1, 1))),
                                   E.LEqExpr(E.AffExpr(np.array([[0.0, 1.0]]), np.zeros((1, 1))), np.zeros((1, 1)))])
    assert p.find_closest_feasible_point()
    assert np.allclose(var.get_value()[:, 0], [-1.0, -1.0], atol=1e-5)
    # min (x0 - 2)^2 + (x1 + 10)^2 inside a trust region of 1 around (0, 0): (1, -1); without: (2, -10)
    p, var = problem([0.0, 0.0], [])
    p.add_obj_expr(E.BoundExpr(E.QuadExpr(2 * np.eye(2), np.array([[-4.0, 20.0]]), np.array([[104.0]])), var))
    assert p.optimize()
    assert np.allclose(var.get_value()[:, 0], [2.0, -10.0], atol=1e-4)
    var._value = np.zeros((2, 1))
    p.save()
    p.add_trust_region(1.0)
    assert p.optimize()
    assert np.allclose(var.get_value()[:, 0], [1.0, -1.0], atol=1e-5)
    assert abs(p.get_value(1.0) - (1.0 + 81.0)) <= 1e-6


@pytest.mark.gpu
def test_prob_stage_methods_penalty_iteration():
    """convexify / update_obj / add_trust_region / optimize / merits step by step on a QCQP, against the
    oracle port's first SQP iteration; the compounded weight of quirk C-1 shows in the second one."""
    st, params, x0 = W.gen_qcqp(1, n=8, m=6)
    prob, var = api_builder.build_prob(st, params[0], x0[0])
    pp = sqp_port.PortProblem(st, params[0], x0[0])
    mu, delta = 3.0, 0.5
    assert abs(prob.get_value(mu) - pp.get_value(mu)) <= 1e-9 * max(1.0, abs(pp.get_value(mu)))
    assert abs(prob.get_max_cnt_violation() - pp.get_max_cnt_violation()) <= 1e-12
    for it in range(2):  # second iteration: the QP weight is mu * mu (quirk C-1), rows are doubled (C-3)
        prob.convexify()
        prob.update_obj(mu)
        prob.save()
        prob.add_trust_region(delta)
        pp.convexify()
        pp.update_obj(mu)
        xs = pp.x[:, 0].copy()
        ok_ref = pp.solve_qp(xs - delta, xs + delta, True, sqp_port.DEFAULT_OSQP)
        assert prob.optimize() == ok_ref
        x1 = var.get_value()[:, 0]
        assert np.abs(x1 - pp.x[:, 0]).max() <= 1e-6
        assert abs(prob.get_approx_value(mu) - pp.get_approx_value(mu)) <= 1e-6 * max(1.0, abs(pp.get_approx_value(mu)))
        assert abs(prob.get_value(mu) - pp.get_value(mu)) <= 1e-6 * max(1.0, abs(pp.get_value(mu)))
        assert np.allclose(prob.get_value(mu, vectorize=True), pp.get_value(mu, vectorize=True), atol=1e-6)


@pytest.mark.gpu
def test_edge_cases_empty_batch_qp_only_problem_and_size_limits():
    from sco_py_b200.engine import Engine, make_settings
    from sco_py_b200.structure import Field
    # empty batch: valid, returns empty results (the reference has no batch notion; ragged ends of a
    # sharded batch produce it)
    st, params, x0 = W.gen_qcqp(2, n=8, m=6)
    eng = Engine(st)
    out = eng.solve_batch(params[:0], x0[:0], make_settings(solver=W.SOLVER_SETTINGS))
    assert out["x"].shape == (0, 8) and out["verdict"].numel() == 0
    out = eng.solve_batch_host(params[:0], x0[:0], make_settings(solver=W.SOLVER_SETTINGS))
    assert out["x"].shape == (0, 8)
    eng.close()
    # a problem without nonlinear constraints is one QP inside the SQP loop: min |x - (2,-10)|^2, x0 <= 1
    ov = np.array([[OSQPVar("x0")], [OSQPVar("x1")]], dtype=object)
    p = Prob()
    for v in ov[:, 0]:
        p.add_osqp_var(v)
    var = Variable(ov, np.zeros((2, 1)))
    p.add_obj_expr(E.BoundExpr(E.QuadExpr(2 * np.eye(2), np.array([[-4.0, 20.0]]), np.array([[104.0]])), var))
    p.add_cnt_expr(E.BoundExpr(E.LEqExpr(E.AffExpr(np.array([[1.0, 0.0]]), np.zeros((1, 1))), np.ones((1, 1))), var))
    st2, params2, x02, _ = batch.compile_batch([p])
    ref = sqp_port.solve(st2, params2[0], x02[0], solver=W.SOLVER_SETTINGS)
    assert _solver().solve(p, method="penalty_sqp") == ref["success"]
    assert np.allclose(var.get_value()[:, 0], ref["x"], atol=1e-5) and np.allclose(ref["x"], [1.0, -10.0], atol=1e-4)
    # size limits are refused with a message, not mis-solved: a working set beyond the shared memory of one SM
    # (rows wider than 32 entries are fine: tests/test_caps.py)
    st3, params3, x03 = W.gen_qcqp(1, n=200, m=4)
    with pytest.raises(RuntimeError, match="exceeds shared memory"):
        Engine(st3)
    # the largest dense kind (n = m = 32) goes through the two-warp solver
    st4, params4, x04 = W.gen_qcqp(2, n=32, m=32)
    eng = Engine(st4)
    assert eng.team == 64
    out = eng.solve_batch(params4, x04, make_settings(solver=W.SOLVER_SETTINGS))
    for i in range(2):
        ref = sqp_port.solve(st4, params4[i], x04[i], solver=W.SOLVER_SETTINGS)
        assert (int(out["verdict"][i]) == 1) == ref["success"]
        assert np.abs(out["x"][i].cpu().numpy() - ref["x"]).max() <= 1e-4 * max(1.0, np.abs(ref["x"]).max())
    eng.close()


@pytest.mark.gpu
def test_solve_batch_buckets_mixed_structures():
    """Solver.solve_batch on problems of different structures: one launch per structure, results in
    the caller's order."""
    built, refs = [], []
    for n, m, first in ((8, 6, 0), (12, 10, 0), (8, 6, 1), (12, 10, 1), (8, 6, 2)):
        st, params, x0 = W.gen_qcqp(1, n=n, m=m, first=first)
        built.append(api_builder.build_prob(st, params[0], x0[0]))
        refs.append(sqp_port.solve(st, params[0], x0[0], solver=W.SOLVER_SETTINGS))
    ok = _solver().solve_batch([p for p, _ in built], method="penalty_sqp")
    for (prob, var), ref, o in zip(built, refs, ok):
        assert o == ref["success"]
        assert np.abs(var.get_value()[:, 0] - ref["x"]).max() <= 1e-4 * max(1.0, np.abs(ref["x"]).max())
