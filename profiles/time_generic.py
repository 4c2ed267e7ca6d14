"""clock64() phase breakdown of the GENERIC shared-memory ADMM loop (sco_qp.cuh generic_loop) on one
penalty QP per problem; needs a -DSCO_TIMING build.   python profiles/time_generic.py [config] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sco_py_b200 import workloads as W
from sco_py_b200.engine import Engine, make_settings
name = sys.argv[1] if len(sys.argv) > 1 else "arm"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
st, params, x0 = W.gen_batch(name, B)
eng = Engine(st)
s = make_settings(solver=W.SOLVER_SETTINGS)
s.force_generic = 1
f, J, b, _ = eng.convexify(params, x0)
for _ in range(2):
    xq, status, iters = eng.qp_solve(params, s, J=J, b=b, lbx=x0 - 1.0, ubx=x0 + 1.0, pi=np.full(B, 1.0), kdup=np.full(B, 1, np.int32))
torch.cuda.synchronize()
xq = xq.cpu().numpy(); it = iters.cpu().numpy().astype(float)
loop, setup = xq[:, 0], xq[:, 2]
ph = [xq[:, 3 + k] for k in range(5)]
print("%s B=%d team %d ctas/sm %d: iters mean %.0f | cycles/iter %.0f | setup %.0f | per iteration: P1 %.0f P2 %.0f P3 %.0f P4 %.0f tests+rest %.0f" % (
    name, B, eng.team, eng.occupancy, it.mean(), (loop / it).mean(), setup.mean(), *[(p / it).mean() for p in ph]))
