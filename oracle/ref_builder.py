"""oracle/ref_builder.py -- TEST INFRASTRUCTURE.

Builds the reference's own objects (sco_py.sco_osqp Prob / Variable / OSQPVar,
sco_py.expr Expr / QuadExpr / AffExpr / EqExpr / LEqExpr / BoundExpr) for one
structured problem and solves it with the reference's `Solver.solve(prob,
method="penalty_sqp")` (sco_py/sco_osqp/solver.py:30-59), exactly the usage of
tests/sco_osqp/test_solver.py:52-84.  Requires /root/reference (this container
only) plus the oracle shims for `osqp` / `numdifftools`.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SCO_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sco_py"))


def import_reference():
    for p in (os.path.join(_HERE, "shims"), _HERE, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import sco_py.expr as expr
    import sco_py.sco_osqp.osqp_utils as osqp_utils
    import sco_py.sco_osqp.prob as prob
    import sco_py.sco_osqp.solver as solver
    import sco_py.sco_osqp.variable as variable
    return dict(expr=expr, osqp_utils=osqp_utils, prob=prob, solver=solver, variable=variable)


def build_prob(ref, st, row, x0):
    """Returns (prob, var) -- reference objects for problem (st, row, x0)."""
    import sqp_port as port  # oracle/sqp_port.py: reuse the family callables
    ex = ref["expr"]
    n = st.n
    pp = port.PortProblem(st, row, x0)
    prob = ref["prob"].Prob()
    ovars = np.empty((n, 1), dtype=object)
    for j in range(n):
        if pp.lb0 is not None:
            ovars[j, 0] = ref["osqp_utils"].OSQPVar("x%05d" % j, float(pp.lb0[j]), float(pp.ub0[j]))
        else:
            ovars[j, 0] = ref["osqp_utils"].OSQPVar("x%05d" % j)  # sorts before "z+_pos_osqp_var"
        prob.add_osqp_var(ovars[j, 0])
    var = ref["variable"].Variable(ovars, value=np.asarray(x0, dtype=float).reshape(n, 1))
    prob.add_var(var)
    prob.add_obj_expr(ex.BoundExpr(ex.QuadExpr(pp.Q, pp.q.reshape(1, n), np.array([[pp.c]])), var))
    if pp.qa is not None:  # AffExpr objective term (quirk C-4)
        prob.add_obj_expr(ex.BoundExpr(ex.AffExpr(pp.qa.reshape(1, n), np.zeros((1, 1))), var))
    if pp.obj_prog is not None:  # black-box objective term, as tests/sco_osqp/test_solver.py:66-68 adds it
        import families as fam
        og = (lambda x: fam.vm_grad(x, pp.obj_prog, 1)) if getattr(st, "obj_prog_flags", 0) & 1 else None
        prob.add_obj_expr(ex.BoundExpr(ex.Expr(lambda x: fam.vm_f(x, pp.obj_prog, 1), og), var))
    if st.m_lin:
        A = pp.A_lin.toarray()
        eq = np.isclose(pp.l_lin, pp.u_lin) & np.isfinite(pp.l_lin)
        # contiguous runs of equal type become one Eq / LEq expression
        r = 0
        while r < st.m_lin:
            e = r
            while e < st.m_lin and eq[e] == eq[r]:
                e += 1
            Ab = A[r:e]
            zero = np.zeros((e - r, 1))
            if eq[r]:
                cnt = ex.EqExpr(ex.AffExpr(Ab, zero), pp.u_lin[r:e].reshape(-1, 1))
            else:
                assert np.all(np.isneginf(pp.l_lin[r:e]))
                cnt = ex.LEqExpr(ex.AffExpr(Ab, zero), pp.u_lin[r:e].reshape(-1, 1))
            prob.add_cnt_expr(ex.BoundExpr(cnt, var))
            r = e
    for b in pp.blocks:
        blk = b.blk
        grad = None if blk.family == port.FAM_FK7 or (blk.family == port.FAM_VM and not blk.ipar[3]) else b.grad
        e = ex.Expr(b.f, grad)
        cnt = (ex.EqExpr if blk.cnt_type == port.CNT_EQ else ex.LEqExpr)(e, b.val)
        gids = None
        if st.n_groups > 1 or blk.group_mask != 1:
            gids = ["g%02d" % g for g in range(st.n_groups) if (blk.group_mask >> g) & 1]
        prob.add_cnt_expr(ex.BoundExpr(cnt, var), gids)
    return prob, var


def solve_with_reference(ref, st, row, x0, solver=None, osqp_kw=None):
    prob, var = build_prob(ref, st, row, x0)
    solv = ref["solver"].Solver()
    for k, v in (solver or {}).items():
        setattr(solv, k, v)
    kw = {}
    for k, v in (osqp_kw or {}).items():
        kw[{"eps_abs": "osqp_eps_abs", "eps_rel": "osqp_eps_rel", "max_iter": "osqp_max_iter"}.get(k, k)] = v
    ok = solv.solve(prob, method="penalty_sqp", **kw)
    x = var.get_value()[:, 0]
    mu_final = None
    return dict(x=x.copy(), success=bool(ok), max_vio=float(prob.get_max_cnt_violation()),
                prob=prob, var=var, nonconverged=list(prob.nonconverged_groups))
