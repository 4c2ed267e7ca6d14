"""Experiment (needs SCO_NVCC_FLAGS="-DSCO_TIMING -DSCO_SKIP_PROBE" build): fraction of failing
termination tests that a lane-local lower bound of the primal residual would decide without mat-vecs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sco_py_b200 import workloads as W
from sco_py_b200.engine import Engine, make_settings
B = 2048
st, params, x0 = W.gen_batch("qcqp", B)
eng = Engine(st)
s = make_settings(solver=W.SOLVER_SETTINGS)
f, J, b, _ = eng.convexify(params, x0)
for pi, kd, delta in ((1.0, 1, 1.0), (10.0, 3, 0.1), (1e3, 8, 0.01), (1e5, 20, 1e-3)):
    xq, status, iters = eng.qp_solve(params, s, J=J, b=b, lbx=x0 - delta, ubx=x0 + delta, pi=np.full(B, pi), kdup=np.full(B, kd, np.int32))
    xq = xq.cpu().numpy()
    tests, fail, decided, wrong, dual = xq[:, 3].sum(), xq[:, 4].sum(), xq[:, 5].sum(), xq[:, 6].sum(), xq[:, 7].sum()
    print("pi %g kd %d delta %g: iters mean %.0f | tests %d failing %d decided by the bound %d (%.1f%%) | wrong %d | dual bound alone %.1f%%" % (
        pi, kd, delta, iters.float().mean().item(), tests, fail, decided, 100.0 * decided / max(fail, 1), wrong, 100.0 * dual / max(fail, 1)))
