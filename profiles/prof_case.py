"""One small invocation of the hot path for ncu (profiles/README.md lists the command lines).
    python profiles/prof_case.py [config] [batch] [launches]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sco_py_b200 import workloads as W  # noqa: E402
from sco_py_b200.engine import Engine, make_settings  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "qcqp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 2
max_sqp = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # > 0: cap SQP iterations (short kernels for ncu)
max_admm = int(sys.argv[5]) if len(sys.argv) > 5 else 0  # > 0: cap ADMM iterations per QP
st, params, x0 = W.gen_batch(name, B)
eng = Engine(st)
s = make_settings(solver=W.SOLVER_SETTINGS)
if max_sqp:
    s.max_sqp_iters = max_sqp
if max_admm:
    s.osqp_max_iter = max_admm
p, x = eng._dev(params), eng._dev(x0)
for _ in range(launches):
    out = eng.solve_batch(p, x, s)
torch.cuda.synchronize()
stats = out["stats"].cpu().numpy()
print(name, B, "converged", int((out["verdict"] == 1).sum()), "mean admm iters", stats[:, 2].mean(),
      "team", eng.team, "smem", eng.smem_bytes, "ctas/sm", eng.occupancy)
