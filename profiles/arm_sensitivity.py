"""Why the arm configuration (C3) cannot be held to 1e-4 on x between two correct implementations.

The arm's only nonlinear constraint is the flange position of a 7-link chain whose Jacobian the reference
obtains by finite differences (expr.py:61-69: numdifftools.Jacobian, step ~5e-7).  Two libm's that differ in the
last bit of sin / cos change f by ~1e-16, the finite-difference Jacobian by ~1e-16 / 5e-7 = 2e-10, and the SQP --
which stops on thresholds (min_approx_improve, min_trust_region_size, solver.py:200-204,246-249) in a valley of
the smoothness objective that is flat along the null space of the three position rows -- may then stop one
iteration earlier or later.  This script measures that sensitivity ON THE CPU ORACLE ALONE: the port against
the same port whose FK value is perturbed by one unit in the last place (deterministically, keyed on x).
No GPU is involved: whatever spread shows up here is a property of the algorithm, not of the CUDA path.

    python profiles/arm_sensitivity.py [count] > profiles/r2_arm_sensitivity.json
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "shims")):
    sys.path.insert(0, p)

import families as fam  # noqa: E402
import sqp_port  # noqa: E402
from sco_py_b200 import workloads as W  # noqa: E402

_orig = fam.fk7_f
ULP = float(os.environ.get("ARM_SENS_ULPS", "1.0")) * 2.0 ** -53


def fk7_perturbed(x):
    f = _orig(x)
    seed = int.from_bytes(hashlib.blake2b(np.ascontiguousarray(x[-7:]).tobytes(), digest_size=8).digest(), "little")
    sign = np.random.default_rng(seed).choice([-1.0, 0.0, 1.0], size=f.shape)
    return f * (1.0 + ULP * sign)


def solve(i, perturbed):
    fam.fk7_f = fk7_perturbed if perturbed else _orig
    st, params, x0 = W.gen_arm(1, first=i)
    r = sqp_port.solve(st, params[0], x0[0], solver=W.SOLVER_SETTINGS)
    fam.fk7_f = _orig
    return r


def job(i):
    a, b = solve(i, False), solve(i, True)
    rel = float(np.abs(a["x"] - b["x"]).max() / max(1.0, np.abs(a["x"]).max()))
    return dict(index=i, rel_dx=rel, sqp_iters=[a["stats"]["sqp_iters"], b["stats"]["sqp_iters"]],
                qp_solves=[a["stats"]["qp_solves"], b["stats"]["qp_solves"]], success=[a["success"], b["success"]],
                dobj_rel=abs(a["objective"] - b["objective"]) / max(1.0, abs(a["objective"])),
                dvio=abs(a["max_vio"] - b["max_vio"]))


if __name__ == "__main__":
    import multiprocessing as mp
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        rows = sorted(pool.imap_unordered(job, range(count)), key=lambda r: r["index"])
    rel = np.array([r["rel_dx"] for r in rows])
    out = dict(what="oracle port vs the same port with the FK value perturbed by 1 ulp (CPU only)", count=count,
               perturbation_ulps=ULP / 2.0 ** -53,
               rel_dx=dict(p50=float(np.percentile(rel, 50)), p90=float(np.percentile(rel, 90)), max=float(rel.max())),
               above_1e_4=int((rel > 1e-4).sum()), above_1e_3=int((rel > 1e-3).sum()),
               different_sqp_iteration_count=int(sum(r["sqp_iters"][0] != r["sqp_iters"][1] for r in rows)),
               verdicts_differ=int(sum(r["success"][0] != r["success"][1] for r in rows)),
               dobj_rel_max=max(r["dobj_rel"] for r in rows), dvio_max=max(r["dvio"] for r in rows), rows=rows)
    print(json.dumps(out, indent=1))
