"""Builds sco_py_b200 API objects (Variable / Prob / BoundExpr / family exprs) for one structured
problem of the synthetic workloads -- the mirror of oracle/ref_builder.py, which builds the
reference's own objects from the same numbers."""
import numpy as np

from sco_py_b200 import expr as E
from sco_py_b200.sco_b200.osqp_utils import OSQPVar
from sco_py_b200.sco_b200.prob import Prob
from sco_py_b200.sco_b200.variable import Variable
from sco_py_b200.structure import CNT_EQ, FAM_CIRCLE2D, FAM_FK7, FAM_QUADFORM


def family_expr(st, blk, row):
    par = st.get(blk.par, row, 10 ** 9)
    n = st.n
    if blk.family == FAM_QUADFORM:
        m, ntri = blk.m, n * (n + 1) // 2
        iu = np.triu_indices(n)
        P = np.zeros((m, n, n))
        tri = np.asarray(par[:m * ntri]).reshape(m, ntri)
        for j in range(m):
            P[j][iu] = tri[j]
            P[j] = P[j] + P[j].T - np.diag(np.diag(P[j]))
        a = np.asarray(par[m * ntri:m * ntri + m * n]).reshape(m, n)
        return E.QuadFormExpr(P, a)
    if blk.family == FAM_CIRCLE2D:
        T, K = blk.ipar[0], blk.ipar[1]
        return E.CircleDistExpr(T, np.asarray(par[:2 * K]).reshape(K, 2), par[2 * K:3 * K])
    if blk.family == FAM_FK7:
        return E.FK7Expr(n)
    raise NotImplementedError(blk.family)


def build_prob(st, row, x0):
    n = st.n
    prob = Prob()
    ov = np.empty((n, 1), dtype=object)
    for j in range(n):
        ov[j, 0] = OSQPVar("x%05d" % j)
        prob.add_osqp_var(ov[j, 0])
    var = Variable(ov, np.asarray(x0, dtype=float).reshape(n, 1))
    prob.add_var(var)
    Q = np.asarray(st.get(st.Q, row, n * n)).reshape(n, n)
    q = np.asarray(st.get(st.q, row, n)).reshape(1, n)
    c = np.asarray(st.get(st.c, row, 1)).reshape(1, 1)
    prob.add_obj_expr(E.BoundExpr(E.QuadExpr(Q, q, c), var))
    if st.m_lin:
        A = np.zeros((st.m_lin, n))
        for r in range(st.m_lin):
            for p in range(st.lin_rowptr[r], st.lin_rowptr[r + 1]):
                A[r, st.lin_col[p]] = st.lin_val[p]
        lo = np.asarray(st.get(st.lin_l, row, st.m_lin))
        hi = np.asarray(st.get(st.lin_u, row, st.m_lin))
        eq = np.isfinite(lo) & (lo == hi)
        r = 0
        while r < st.m_lin:  # contiguous runs of one kind become one Eq / LEq expression
            e = r
            while e < st.m_lin and eq[e] == eq[r]:
                e += 1
            cls = E.EqExpr if eq[r] else E.LEqExpr
            prob.add_cnt_expr(E.BoundExpr(cls(E.AffExpr(A[r:e], np.zeros((e - r, 1))), hi[r:e].reshape(-1, 1)), var))
            r = e
    for blk in st.blocks:
        val = np.asarray(st.get(blk.val, row, blk.m)).reshape(-1, 1)
        cls = E.EqExpr if blk.cnt_type == CNT_EQ else E.LEqExpr
        prob.add_cnt_expr(E.BoundExpr(cls(family_expr(st, blk, row), val), var))
    return prob, var


def build_qcqp_variant(i, n=8, m=6, aff=False, bounds=False, groups=None):
    """QCQP i of the synthetic family (workloads.gen_qcqp) with optional extras of the reference API:
    an AffExpr objective term (quirk C-4), user bounds on the scalar variables, constraint groups
    (`groups` = list of (row slice, group ids))."""
    from sco_py_b200 import workloads as W
    st, params, x0 = W.gen_qcqp(1, n=n, m=m, first=i)
    fam = family_expr(st, st.blocks[0], params[0])
    val = np.asarray(st.get(st.blocks[0].val, params[0], m)).reshape(-1, 1)
    rng = np.random.default_rng(900 + i)
    prob = Prob()
    ov = np.empty((n, 1), dtype=object)
    for j in range(n):
        if bounds:
            lo = float(x0[0, j] - rng.uniform(-0.2, 1.0))  # the start value may lie outside its box
            ov[j, 0] = OSQPVar("x%05d" % j, lo, lo + float(rng.uniform(0.1, 2.0)))
        else:
            ov[j, 0] = OSQPVar("x%05d" % j)
        prob.add_osqp_var(ov[j, 0])
    var = Variable(ov, x0[0].reshape(n, 1))
    prob.add_var(var)
    Q = np.asarray(st.get(st.Q, params[0], n * n)).reshape(n, n)
    q = np.asarray(st.get(st.q, params[0], n)).reshape(1, n)
    prob.add_obj_expr(E.BoundExpr(E.QuadExpr(Q, q, np.zeros((1, 1))), var))
    if aff:
        prob.add_obj_expr(E.BoundExpr(E.AffExpr(0.3 * rng.standard_normal((1, n)), np.array([[0.25]])), var))
    if groups is None:
        prob.add_cnt_expr(E.BoundExpr(E.LEqExpr(fam, val), var))
    else:
        for sl, gids in groups:
            prob.add_cnt_expr(E.BoundExpr(E.LEqExpr(E.QuadFormExpr(fam.P[sl], fam.a[sl]), val[sl]), var), group_ids=gids)
    return prob, var
