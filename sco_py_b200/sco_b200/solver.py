"""Solver: same attributes (sco_py/sco_osqp/solver.py:17-28) and `solve` signature (:30-59) as the
reference's OSQP backend.  `solve` hands the whole penalty SQP -- closest feasible point,
convexification, penalty-QP assembly, ADMM, merit / trust-region / penalty-increase state machine
(:62-253) -- to ONE kernel launch of the B200 engine; `solve_batch` does the same for B problems of
shared structure, which is what the engine is for."""
import numpy as np

from .. import batch
from . import osqp_utils


def nonconverged_list(gids, mask):
    """prob.nonconverged_groups as the reference leaves it (solver.py:209-235): the groups of the first loop in
    gid2ind order (sorted ids, prob.py:538-542), then -- the reference appends again -- every violated group whose
    model improvement is below the threshold, in sorted order.  `mask` is sco_batch_io.d_nonconverged."""
    first = [g for k, g in enumerate(gids) if (mask >> k) & 1]
    second = [g for k, g in enumerate(gids) if (mask >> (16 + k)) & 1]
    return first + second if first else []


class Solver(object):
    def __init__(self):
        self.improve_ratio_threshold = 0.25
        self.min_trust_region_size = 1e-4
        self.min_approx_improve = 1e-8
        self.max_iter = 50                      # never read by the reference either (quirk C-10)
        self.trust_shrink_ratio = 0.1
        self.trust_expand_ratio = 1.5
        self.cnt_tolerance = 1e-4
        self.max_merit_coeff_increases = 1
        self.merit_coeff_increase_ratio = 10
        self.initial_trust_region_size = 1
        self.initial_penalty_coeff = 1e3
        self._engines = {}
        self.last_report = None

    _ATTRS = ("improve_ratio_threshold", "min_trust_region_size", "min_approx_improve", "trust_shrink_ratio",
              "trust_expand_ratio", "cnt_tolerance", "max_merit_coeff_increases", "merit_coeff_increase_ratio",
              "initial_trust_region_size", "initial_penalty_coeff")

    def solve(self, prob, method=None, tol=None, verbose=False, osqp_eps_abs=osqp_utils.DEFAULT_EPS_ABS,
              osqp_eps_rel=osqp_utils.DEFAULT_EPS_REL, osqp_max_iter=osqp_utils.DEFAULT_MAX_ITER,
              rho=osqp_utils.DEFAULT_RHO, adaptive_rho=osqp_utils.DEFAULT_ADAPTIVE_RHO,
              sigma=osqp_utils.DEFAULT_SIGMA):
        """Returns whether the solve succeeded; the solution is written into the Variables."""
        return self.solve_batch([prob], method=method, tol=tol, verbose=verbose, osqp_eps_abs=osqp_eps_abs,
                                osqp_eps_rel=osqp_eps_rel, osqp_max_iter=osqp_max_iter, rho=rho,
                                adaptive_rho=adaptive_rho, sigma=sigma)[0]

    def solve_batch(self, probs, method=None, tol=None, verbose=False, osqp_eps_abs=osqp_utils.DEFAULT_EPS_ABS,
                    osqp_eps_rel=osqp_utils.DEFAULT_EPS_REL, osqp_max_iter=osqp_utils.DEFAULT_MAX_ITER,
                    rho=osqp_utils.DEFAULT_RHO, adaptive_rho=osqp_utils.DEFAULT_ADAPTIVE_RHO,
                    sigma=osqp_utils.DEFAULT_SIGMA, device=0):
        """B problems of shared structure in one launch -> list of B bools."""
        if tol is not None:
            self.min_trust_region_size = tol
            self.min_approx_improve = tol
            self.cnt_tolerance = tol
        if method != "penalty_sqp":
            raise Exception("This method is not supported.")
        from ..engine import Engine, make_settings
        settings = make_settings(
            solver={k: getattr(self, k) for k in self._ATTRS},
            osqp=dict(eps_abs=osqp_eps_abs, eps_rel=osqp_eps_rel, max_iter=osqp_max_iter, rho=rho,
                      adaptive_rho=adaptive_rho, sigma=sigma))
        # problems of different structures are bucketed: one launch per structure (BASELINE.json configs[4])
        compiled = [batch.compile_problem(p) for p in probs]
        buckets = {}
        for i, cp in enumerate(compiled):
            buckets.setdefault(batch.group_key(cp), []).append(i)
        B = len(probs)
        out = dict(x=[None] * B, verdict=np.zeros(B, np.int32), merit=np.zeros(B), objective=np.zeros(B),
                   max_vio=np.zeros(B), stats=np.zeros((B, 4), np.int32), nonconverged=np.zeros(B, np.int32))
        for idx in buckets.values():
            st, params, x0, cps = batch.compile_batch([probs[i] for i in idx], compiled=[compiled[i] for i in idx])
            key = (batch.signature(st), device)
            if key not in self._engines:
                self._engines[key] = Engine(st, device=device)
            res = self._engines[key].solve_batch_host(params, x0, settings)
            for k, i in enumerate(idx):
                batch.scatter_solution(cps[k], res["x"][k])
                probs[i].nonconverged_groups = nonconverged_list(cps[k].gids, int(res["nonconverged"][k]))
                probs[i]._stage = None
                out["x"][i] = res["x"][k]
                for name in ("verdict", "merit", "objective", "max_vio", "stats", "nonconverged"):
                    out[name][i] = res[name][k]
                # prob.py:204 runs the callback after every successful QP, on the caller's thread; the whole SQP
                # is one kernel here, so the hook runs once per problem when its result has been delivered
                probs[i]._callback()
        self.last_report = out
        if verbose:
            print("sco_b200: %d problems, %d converged, mean ADMM iterations %.0f"
                  % (len(probs), int((out["verdict"] == 1).sum()), float(out["stats"][:, 2].mean())))
        return [bool(v == 1) for v in out["verdict"]]
