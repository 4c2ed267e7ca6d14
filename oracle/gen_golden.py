"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/*.npz.

Runs the UNMODIFIED reference (`/root/reference/sco_py`, imported on top of the
oracle shims for the two third-party modules that are absent from this image,
see oracle/shims/) through its own public entry point
`Solver.solve(prob, method="penalty_sqp")` (sco_py/sco_osqp/solver.py:30-59) on
the first problems of every synthetic workload (sco_py_b200/workloads.py) and
records, per problem, what the reference left behind:

  x        Variable.get_value() after solve              (variable.py:25-33)
  success  the bool returned by Solver.solve             (solver.py:97-105)
  max_vio  Prob.get_max_cnt_violation()                  (prob.py:592-603)
  qp_*     every (P, q, A, l, u) the reference handed to osqp.OSQP().setup
           (osqp_utils.py:195-214) for the first problem of the config, with
           the x / status / iteration count that came back -- the QP-level
           known answers for oracle/osqp_core.c.

The reference cannot travel to the GPU box; these vectors do.  Only this
container (where /root/reference exists) can regenerate them:

    python oracle/gen_golden.py            # writes tests/golden/ref_<config>.npz
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(HERE, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_builder  # noqa: E402
from sco_py_b200 import workloads as W  # noqa: E402

COUNTS = {"qcqp": 6, "point_robot": 4, "arm": 4}
QP_DUMPS = 6  # QPs recorded for problem 0 of each config (first, then evenly spaced)


class _Recorder(object):
    """Wraps the osqp shim's OSQP so every setup()/solve() pair is captured."""

    def __init__(self, osqp_module):
        self.mod = osqp_module
        self.orig = osqp_module.OSQP
        self.log = []
        rec = self

        class Recording(self.orig):
            def setup(self, P=None, q=None, A=None, l=None, u=None, **kw):
                self._rec = dict(P=sp.csc_matrix(P).toarray(), q=np.array(q, dtype=float),
                                 A=sp.csc_matrix(A).toarray(), l=np.array(l, dtype=float),
                                 u=np.array(u, dtype=float),
                                 settings={k: v for k, v in kw.items()})
                return super().setup(P=P, q=q, A=A, l=l, u=u, **kw)

            def solve(self):
                res = super().solve()
                r = self._rec
                r.update(x=res.x.copy(), status=res.info.status_val, iters=res.info.iter)
                rec.log.append(r)
                return res

        self.cls = Recording

    def __enter__(self):
        self.mod.OSQP = self.cls
        return self

    def __exit__(self, *a):
        self.mod.OSQP = self.orig


def main():
    if not ref_builder.reference_available():
        raise SystemExit("reference not found at %s" % ref_builder.REFERENCE_ROOT)
    ref = ref_builder.import_reference()
    import osqp as osqp_shim
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, count in COUNTS.items():
        st, params, x0 = W.GENERATORS[name](count)
        xs, oks, vios = [], [], []
        arrays = {}
        for i in range(count):
            t0 = time.time()
            with _Recorder(osqp_shim) as rec:
                r = ref_builder.solve_with_reference(ref, st, params[i], x0[i], solver=W.SOLVER_SETTINGS)
            xs.append(r["x"])
            oks.append(r["success"])
            vios.append(r["max_vio"])
            print("%s[%d]: success=%s max_vio=%.3e qps=%d (%.1fs)" % (name, i, r["success"], r["max_vio"],
                                                                      len(rec.log), time.time() - t0))
            if i == 0:
                pick = sorted(set([0, 1] + list(np.linspace(2, len(rec.log) - 1, QP_DUMPS - 2).astype(int))))
                arrays["qp_index"] = np.array(pick)
                for k, qi in enumerate(pick):
                    q = rec.log[qi]
                    A = sp.csr_matrix(q["A"])
                    P = sp.csr_matrix(q["P"])
                    for key, M in (("A", A), ("P", P)):
                        arrays["qp%d_%s_data" % (k, key)] = M.data
                        arrays["qp%d_%s_indices" % (k, key)] = M.indices.astype(np.int32)
                        arrays["qp%d_%s_indptr" % (k, key)] = M.indptr.astype(np.int32)
                        arrays["qp%d_%s_shape" % (k, key)] = np.array(M.shape)
                    for key in ("q", "l", "u", "x"):
                        arrays["qp%d_%s" % (k, key)] = q[key]
                    arrays["qp%d_status" % k] = np.array([q["status"], q["iters"]])
                    s = q["settings"]
                    arrays["qp%d_settings" % k] = np.array([s["rho"], s["sigma"], s["eps_abs"], s["eps_rel"],
                                                           float(s["max_iter"]), float(s["adaptive_rho"])])
        arrays.update(x=np.array(xs), success=np.array(oks), max_vio=np.array(vios),
                      count=np.array(count))
        path = os.path.join(out_dir, "ref_%s.npz" % name)
        np.savez_compressed(path, **arrays)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
