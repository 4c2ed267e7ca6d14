"""clock64() breakdown of one penalty QP per problem (needs a library built with
SCO_NVCC_FLAGS=-DSCO_TIMING python -m sco_py_b200.build --force): loop / check / setup cycles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sco_py_b200 import workloads as W
from sco_py_b200.engine import Engine, make_settings
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
st, params, x0 = W.gen_batch("qcqp", B)
eng = Engine(st)
s = make_settings(solver=W.SOLVER_SETTINGS)
f, J, b, _ = eng.convexify(params, x0)
for force in (0, 1):
    s.force_generic = force
    for _ in range(2):
        xq, status, iters = eng.qp_solve(params, s, J=J, b=b, lbx=x0 - 1.0, ubx=x0 + 1.0, pi=np.full(B, 1.0), kdup=np.full(B, 1, np.int32))
    torch.cuda.synchronize()
    xq = xq.cpu().numpy(); it = iters.cpu().numpy()
    loop, chk, setup = xq[:, 0], xq[:, 1], xq[:, 2]
    print("force_generic=%d B=%d iters mean %.0f | cycles/iter (loop incl. checks) %.0f | check cycles per check %.0f | checks share %.1f%% | setup cycles %.0f" % (
        force, B, it.mean(), (loop / it).mean(), (chk / np.maximum(it // 25, 1)).mean(), 100 * chk.sum() / loop.sum(), setup.mean()))
    if not force:
        nchk = np.maximum(it // 25, 1)
        print("   test stages (cycles per test): publish+barrier %.0f | mat-vecs+residuals %.0f | warp reductions %.0f | exchange %.0f | decision %.0f" % tuple(
            (xq[:, 3 + k] / nchk).mean() for k in range(5)))
