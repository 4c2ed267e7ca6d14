"""oracle/sqp_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE ("port" oracle).

NumPy restatement of the reference's penalty-SQP path for ONE problem given in
structured form (the same structure/params the CUDA engine receives):

  Solver._penalty_sqp / _min_merit_fn      sco_py/sco_osqp/solver.py:62-283
  Prob.find_closest_feasible_point         sco_py/sco_osqp/prob.py:369-412
  Prob.convexify / update_obj              prob.py:414-426, 461-512, 522-544
  Prob.get_value / get_approx_value /
       get_max_cnt_violation               prob.py:547-630
  osqp_utils.optimize (P, q, A, l, u)      sco_py/sco_osqp/osqp_utils.py:113-221
  Variable.add_trust_region/save/restore   sco_py/sco_osqp/variable.py:37-73

including the OSQP-backend quirks of SURVEY.md Appendix C: compounded penalty
weight (C-1, prob.py:424-426), frozen row sparsity (C-2, prob.py:488-504),
duplicated penalty rows (C-3, prob.py:508-509) and "failed QP => y-converged"
(C-5, solver.py:140,151-153).  The QP arithmetic is oracle/osqp_core.c.

It is validated against the UNMODIFIED reference (run from /root/reference on
the oracle shims) by tests/test_oracle_port.py, and its outputs on the golden
problems are committed under tests/golden/ (oracle/gen_golden.py).
It travels to the GPU box (the reference does not) and is the checker of the
`-m gpu` parity tests and the `cpu_baseline` / `--impl reference` arm of bench.py.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.join(_HERE, "shims") not in sys.path:
    sys.path.insert(0, os.path.join(_HERE, "shims"))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import families as fam  # noqa: E402  (oracle/families.py)
import numdifftools as nd  # noqa: E402  (oracle/shims/numdifftools)
import osqp as osqp_shim  # noqa: E402  (oracle/shims/osqp)

FAM_QUADFORM, FAM_CIRCLE2D, FAM_FK7, FAM_VM = 1, 2, 3, 4
CNT_LEQ, CNT_EQ = 0, 1

DEFAULT_SOLVER = dict(  # sco_py/sco_osqp/solver.py:17-28
    improve_ratio_threshold=0.25, min_trust_region_size=1e-4, min_approx_improve=1e-8,
    max_iter=50, trust_shrink_ratio=0.1, trust_expand_ratio=1.5, cnt_tolerance=1e-4,
    max_merit_coeff_increases=1, merit_coeff_increase_ratio=10.0,
    initial_trust_region_size=1.0, initial_penalty_coeff=1e3)
DEFAULT_OSQP = dict(  # sco_py/sco_osqp/osqp_utils.py:10-15
    max_iter=100000, sigma=5e-10, rho=0.1, adaptive_rho=False, eps_abs=1e-6, eps_rel=1e-9)


def _field(st, f, row, size):
    if f.off < 0:
        return np.zeros(size)
    src = st.shared if f.shared else row
    return np.asarray(src[f.off:f.off + size], dtype=float)


class _BlockFn(object):
    """f / grad of one nonlinear constraint block at the problem's parameters."""

    def __init__(self, st, blk, row):
        self.blk = blk
        n = st.n
        self.n = n
        self.val = _field(st, blk.val, row, blk.m)[:, None]
        if blk.family == FAM_QUADFORM:
            m = blk.m
            ntri = n * (n + 1) // 2
            par = _field(st, blk.par, row, m * ntri + m * n)
            self.Pm = fam.unpack_sym(par[:m * ntri].reshape(m, ntri), n)
            self.a = par[m * ntri:].reshape(m, n)
            self.f = lambda x: fam.quadform_f(x, self.Pm, self.a)
            self.grad = lambda x: fam.quadform_grad(x, self.Pm, self.a)
        elif blk.family == FAM_CIRCLE2D:
            T, K = blk.ipar[0], blk.ipar[1]
            par = _field(st, blk.par, row, 3 * K)
            cen = par[:2 * K].reshape(K, 2)
            rad = par[2 * K:]
            self.f = lambda x: fam.circle_f(x, cen, rad, T)
            self.grad = lambda x: fam.circle_grad(x, cen, rad, T)
        elif blk.family == FAM_FK7:
            self.f = fam.fk7_f
            # no analytic gradient: expr.py:86-87 -> numdifftools.Jacobian
            self.grad = lambda x: nd.Jacobian(lambda v: fam.fk7_f(v.reshape(n, 1)).ravel())(x)
        elif blk.family == FAM_VM:
            m, n_instr = blk.m, blk.ipar[2]
            prog = _field(st, blk.par, row, m + 2 * n_instr)
            self.f = lambda x: fam.vm_f(x, prog, m)
            if blk.ipar[3]:  # Expr(f, grad): the user's analytic gradient (expr.py:86-88)
                self.grad = lambda x: fam.vm_grad(x, prog, m)
            else:  # a black box without analytic gradient: expr.py:86-87 -> numdifftools.Jacobian
                self.grad = lambda x: nd.Jacobian(lambda v: fam.vm_f(v, prog, m).ravel())(x)
        else:
            raise NotImplementedError(blk.family)

    def violation(self, x):
        v = self.f(x) - self.val
        if self.blk.cnt_type == CNT_EQ:
            return np.abs(v)  # prob.py:585-586
        return np.maximum(v, 0.0)  # prob.py:587-590


class PortProblem(object):
    def __init__(self, st, row, x0):
        self.st = st
        n = st.n
        self.n = n
        self.Q = _field(st, st.Q, row, n * n).reshape(n, n)
        self.q = _field(st, st.q, row, n)
        self.c = float(_field(st, st.c, row, 1)[0])
        # AffExpr objective terms (prob.py:97-103 files them under _quad_obj_exprs): exact value qa'x, QP weight `wa`
        self.qa = _field(st, st.qa, row, n) if getattr(st, "qa", None) is not None and st.qa.off >= 0 else None
        self.wa = 0.0
        self.lb0 = self.ub0 = None  # user bounds of the scalar variables (OSQPVar lb / ub)
        if getattr(st, "lb0", None) is not None and st.lb0.off >= 0:
            self.lb0, self.ub0 = _field(st, st.lb0, row, n), _field(st, st.ub0, row, n)
        self.nonconverged = []  # prob.nonconverged_groups (group indices, reference order incl. its duplicates)
        self.m_lin = st.m_lin
        if st.m_lin:
            self.A_lin = sp.csr_matrix((st.lin_val, st.lin_col, st.lin_rowptr), shape=(st.m_lin, n))
            self.l_lin = _field(st, st.lin_l, row, st.m_lin)
            self.u_lin = _field(st, st.lin_u, row, st.m_lin)
        else:
            self.A_lin = sp.csr_matrix((0, n))
            self.l_lin = self.u_lin = np.zeros(0)
        self.blocks = [_BlockFn(st, b, row) for b in st.blocks]
        # non-quadratic objective term (prob.py:97-103 -> _nonquad_obj_exprs), a stack program
        self.obj_prog = None
        if getattr(st, "obj_prog_len", 0):
            self.obj_prog = _field(st, st.obj_prog, row, 1 + 2 * st.obj_prog_len)
        self.Hq = None  # its convex quadratic model (expr.py:143-153): 0.5 x'Hq x + aq x + bq
        self.aq = None
        self.bq = 0.0
        self.x = np.asarray(x0, dtype=float).reshape(n, 1).copy()
        self.x_saved = None
        # penalty-QP bookkeeping
        self.pi = 1.0          # compounded slack weight (quirk C-1)
        self.kdup = 0          # number of copies of every penalty row (quirk C-3)
        self.masks = None      # frozen sparsity (quirk C-2)
        self.J = None
        self.b = None
        self.stats = dict(qp_solves=0, admm_iters=0, sqp_iters=0, last_status=0)
        self.trace = []

    # -- objective / merit ------------------------------------------------
    def objective(self, x):
        # QuadExpr.eval, expr.py:205-206 (+ Expr.eval of the non-quadratic term, prob.py:573-574)
        v = float(0.5 * x[:, 0] @ (self.Q @ x[:, 0]) + self.q @ x[:, 0] + self.c)
        if self.qa is not None:  # AffExpr.eval, expr.py:173-174
            v += float(self.qa @ x[:, 0])
        if self.obj_prog is not None:
            v += float(fam.vm_f(x, self.obj_prog, 1)[0, 0])
        return v

    def objective_model(self, x):
        # quadratic terms + the degree-2 model of the non-quadratic one (prob.py:624-626)
        v = float(0.5 * x[:, 0] @ (self.Q @ x[:, 0]) + self.q @ x[:, 0] + self.c)
        if self.qa is not None:
            v += float(self.qa @ x[:, 0])
        if self.obj_prog is not None:
            v += float(0.5 * x[:, 0] @ (self.Hq @ x[:, 0]) + self.aq @ x[:, 0] + self.bq)
        return v

    def group_vec(self, per_block_sums):
        g = np.zeros(self.st.n_groups)
        for blk, s in zip(self.st.blocks, per_block_sums):
            for gi in range(self.st.n_groups):
                if (blk.group_mask >> gi) & 1:
                    g[gi] += s
        return g

    def get_value(self, mu, vectorize=False):  # prob.py:547-579
        sums = [float(np.sum(b.violation(self.x))) for b in self.blocks]
        if vectorize:
            return self.group_vec(sums)
        value = self.objective(self.x)
        for s in sums:
            value += mu * s
        return value

    def get_approx_value(self, mu, vectorize=False):  # prob.py:605-630
        sums = []
        for b, J, bb in zip(self.blocks, self.J, self.b):
            v = J @ self.x + bb
            pen = np.abs(v) if b.blk.cnt_type == CNT_EQ else np.maximum(v, 0.0)
            sums.append(float(np.sum(pen)))
        if vectorize:
            return self.group_vec(sums)
        value = self.objective_model(self.x)
        for s in sums:
            value += mu * s
        return value

    def get_max_cnt_violation(self):  # prob.py:592-603
        mv = 0.0
        for b in self.blocks:
            mv = max(mv, float(np.amax(b.violation(self.x))))
        return mv

    # -- convexification + penalty bookkeeping ------------------------------
    def convexify(self):  # prob.py:522-544 ; expr.py:139-142, 327-328, 366-367
        if self.obj_prog is not None:  # Expr.convexify degree 2, expr.py:143-153
            flat = lambda v: fam.vm_f(v, self.obj_prog, 1).ravel()
            H = nd.Hessian(flat)(self.x[:, 0].copy())
            lam = float(np.min(np.linalg.eigvalsh(H)))
            if lam < 0:
                H = H - np.eye(self.n) * lam
            if getattr(self.st, "obj_prog_flags", 0) & 1:  # Expr(f, grad): analytic gradient, numerical Hessian
                g = fam.vm_grad(self.x, self.obj_prog, 1)
            else:
                g = nd.Jacobian(flat)(self.x)  # (1, n)
            self.Hq = H
            self.aq = (g - self.x.T @ H)[0]
            self.bq = float(0.5 * self.x[:, 0] @ (H @ self.x[:, 0]) - g[0] @ self.x[:, 0] + flat(self.x)[0])
        self.J, self.b = [], []
        for b in self.blocks:
            J = np.asarray(b.grad(self.x), dtype=float)
            bb = -J @ self.x + b.f(self.x) - b.val
            self.J.append(J)
            self.b.append(bb)

    def update_obj(self, mu):  # prob.py:414-426
        # `quirks` switches the OSQP-backend behaviours off one by one ("intended semantics", the
        # maths of the Gurobi backend: sco_gurobi/prob.py:316-318,365-373): fixed weight, fresh
        # sparsity, single copy of the penalty rows
        q = getattr(self, "quirks", None) or {}
        if q.get("freeze_sparsity", True):
            if self.masks is None:  # _lazy_spawn_osqp_cnts, prob.py:434-444
                self.masks = [(J != 0.0) for J in self.J]
        else:
            self.masks = [np.ones_like(J, dtype=bool) for J in self.J]
        self.kdup = self.kdup + 1 if q.get("duplicate_rows", True) else 1
        self.pi = self.pi * mu if q.get("compound_penalty", True) else mu
        # quirk C-4 (prob.py:220-221,240-249 then 424-426): another copy of the AffExpr objective coefficients is
        # appended to the penalty terms, then every stored coefficient is multiplied by the penalty coefficient
        self.wa = (self.wa + 1.0) * mu if q.get("aff_obj_quirk", True) else 1.0

    # -- QP ------------------------------------------------------------------
    def solve_qp(self, lbx, ubx, penalty, osqp_kw, P_override=None, q_override=None):
        """osqp_utils.optimize (osqp_utils.py:113-221) on the current term lists."""
        n = self.n
        ns = sum(b.blk.m * (2 if b.blk.cnt_type == CNT_EQ else 1) for b in self.blocks) if penalty else 0
        nq = n + ns
        P = np.zeros((nq, nq))
        qv = np.zeros(nq)
        if P_override is not None:
            P[:n, :n] = P_override
            qv[:n] = q_override
        else:
            P[:n, :n] = 0.5 * (self.Q + self.Q.T)
            qv[:n] = self.q
            if self.qa is not None:
                qv[:n] += self.wa * self.qa
            if self.obj_prog is not None and self.Hq is not None:  # prob.py:348-367
                P[:n, :n] += 0.5 * (self.Hq + self.Hq.T)
                qv[:n] += self.aq
        rows = []
        lo = []
        hi = []
        if self.m_lin:
            rows.append(sp.hstack([self.A_lin, sp.csr_matrix((self.m_lin, ns))], format="csr"))
            lo.append(self.l_lin)
            hi.append(self.u_lin)
        if penalty:
            qv[n:] = self.pi
            pen_rows = []
            pen_lo = []
            pen_hi = []
            so = n
            for b, J, bb, M in zip(self.blocks, self.J, self.b, self.masks):
                m = b.blk.m
                blkA = np.zeros((m, nq))
                blkA[:, :n] = J * M
                if b.blk.cnt_type == CNT_EQ:  # prob.py:280-315
                    blkA[np.arange(m), so + np.arange(m)] = -1.0
                    blkA[np.arange(m), so + m + np.arange(m)] = 1.0
                    pen_lo.append(-bb[:, 0])
                    so += 2 * m
                else:  # prob.py:251-278
                    blkA[np.arange(m), so + np.arange(m)] = -1.0
                    pen_lo.append(np.full(m, -np.inf))
                    so += m
                pen_hi.append(-bb[:, 0])
                pen_rows.append(sp.csr_matrix(blkA))
            for _ in range(self.kdup):  # quirk C-3
                rows.extend(pen_rows)
                lo.extend(pen_lo)
                hi.extend(pen_hi)
        rows.append(sp.identity(nq, format="csr"))
        lo.append(np.concatenate([lbx, np.zeros(ns)]))
        hi.append(np.concatenate([ubx, np.full(ns, np.inf)]))
        A = sp.vstack(rows, format="csc")
        l = np.concatenate(lo)
        u = np.concatenate(hi)
        m = osqp_shim.OSQP()
        m.setup(P=sp.csc_matrix(np.triu(P)), q=qv, A=A, l=l, u=u, rho=osqp_kw["rho"],
                sigma=osqp_kw["sigma"], eps_abs=osqp_kw["eps_abs"], eps_rel=osqp_kw["eps_rel"],
                delta=1e-7, polish=False, adaptive_rho=osqp_kw["adaptive_rho"], warm_start=True,
                verbose=False, max_iter=osqp_kw["max_iter"])
        res = m.solve()
        self.stats["qp_solves"] += 1
        self.stats["admm_iters"] += res.info.iter
        self.stats["last_status"] = res.info.status_val
        self.last_qp = (P, qv, A, l, u, res)
        if res.info.status_val not in (1, 2):  # prob.py:197-198
            return False
        self.x = res.x[:n].reshape(n, 1).copy()
        return True

    def find_closest_feasible_point(self):  # prob.py:369-412 (default OSQP settings, solver.py:81)
        n = self.n
        x0 = self.x[:, 0]
        inf = np.full(n, np.inf)
        lo, hi = (-inf, inf) if self.lb0 is None else (self.lb0, self.ub0)  # whatever is on the OSQPVars
        return self.solve_qp(lo, hi, False, DEFAULT_OSQP, P_override=2.0 * np.eye(n),
                             q_override=-2.0 * x0)


def solve(st, row, x0, solver=None, osqp_kw=None, max_sqp_iters=10000, keep_trace=False, quirks=None):
    """Returns dict(x, success, merit, objective, max_vio, stats, trace)."""
    S = dict(DEFAULT_SOLVER)
    S.update(solver or {})
    O = dict(DEFAULT_OSQP)
    O.update(osqp_kw or {})
    p = PortProblem(st, row, x0)
    p.quirks = quirks
    delta0 = S["initial_trust_region_size"]
    mu = S["initial_penalty_coeff"]
    delta = delta0
    success = False

    def result(ok):
        return dict(x=p.x[:, 0].copy(), success=bool(ok), merit=p.get_value(mu),
                    objective=p.objective(p.x), max_vio=p.get_max_cnt_violation(),
                    stats=dict(p.stats), trace=p.trace, mu=mu, nonconverged=list(p.nonconverged))

    if not p.find_closest_feasible_point():  # solver.py:81-82
        return result(False)
    for _ in range(S["max_merit_coeff_increases"]):  # solver.py:84
        success = _min_merit_fn(p, S, O, mu, delta, max_sqp_iters, keep_trace)
        if p.get_max_cnt_violation() > S["cnt_tolerance"]:  # solver.py:94-96
            mu *= S["merit_coeff_increase_ratio"]
            delta = delta0
        else:
            return result(success)  # solver.py:97-101
    return result(False)  # solver.py:105


def _min_merit_fn(p, S, O, mu, delta, max_sqp_iters, keep_trace):  # solver.py:108-253
    ng = p.st.n_groups if p.blocks else 0
    overlap = p.st.group_overlap
    sqp_iter = 1
    while True:
        p.stats["sqp_iters"] += 1
        p.convexify()
        p.update_obj(mu)
        merit = p.get_value(mu)
        merit_vec = p.get_value(mu, True) if ng else np.zeros(0)
        p.x_saved = p.x.copy()
        while True:
            lbx = p.x_saved[:, 0] - delta  # variable.py:43-45
            ubx = p.x_saved[:, 0] + delta
            ok = p.solve_qp(lbx, ubx, True, O)  # result ignored, solver.py:140
            model_merit = p.get_approx_value(mu)
            model_vec = p.get_approx_value(mu, True) if ng else np.zeros(0)
            new_merit = p.get_value(mu)
            approx = merit - model_merit
            if not approx:
                approx += 1e-12
            approx_vec = merit_vec - model_vec
            violated = merit_vec > S["cnt_tolerance"]
            if approx_vec.shape == (0,):
                approx_vec = np.array([approx])
                violated = approx_vec > -np.inf
            exact = merit - new_merit
            ratio = exact / approx
            if keep_trace:
                p.trace.append(dict(sqp_iter=sqp_iter, delta=delta, merit=merit, model=model_merit,
                                    new=new_merit, qp_ok=ok, status=p.stats["last_status"],
                                    x=p.x[:, 0].copy()))
            if approx < -1e-5:  # _bad_model
                p.x = p.x_saved.copy()
                return False
            if approx < S["min_approx_improve"]:  # _y_converged
                p.x = p.x_saved.copy()
                return True
            nonconv = []
            for g in range(ng):  # solver.py:209-225
                if violated[g] and approx_vec[g] < S["min_approx_improve"]:
                    ov = False
                    if overlap is not None:
                        for g2 in range(ng):
                            if g2 != g and overlap[g, g2] and approx_vec[g2] > S["min_approx_improve"]:
                                ov = True
                                break
                    if not ov:
                        nonconv.append(g)
            p.nonconverged = list(nonconv)  # solver.py:208: cleared at every evaluation of this block
            if nonconv:
                p.x = p.x_saved.copy()
                for g in range(ng):  # solver.py:231-233: appended again, without the overlap rule
                    if violated[g] and approx_vec[g] < S["min_approx_improve"]:
                        p.nonconverged.append(g)
                return True
            if exact < 0 or ratio < S["improve_ratio_threshold"]:  # _shrink_trust_region
                p.x = p.x_saved.copy()
                delta *= S["trust_shrink_ratio"]
            else:
                delta *= S["trust_expand_ratio"]
                break
            if delta < S["min_trust_region_size"]:  # _x_converged
                return True
        sqp_iter += 1
        if sqp_iter > max_sqp_iters:
            return False
