"""Per-problem solve statistics of one batch (sqp_iters, qp_solves, admm_iters, last_status, verdict)
saved to gpurun_out/ for offline analysis of the iteration-count tail.
    python profiles/dump_stats.py [config] [batch] [out.npz]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sco_py_b200 import workloads as W  # noqa: E402
from sco_py_b200.engine import Engine, make_settings  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "qcqp"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "stats_%s_%d.npz" % (name, B))
st, params, x0 = W.gen_batch(name, B)
eng = Engine(st)
s = make_settings(solver=W.SOLVER_SETTINGS)
p, x = eng._dev(params), eng._dev(x0)
torch.cuda.synchronize()
t0 = time.time()
out = eng.solve_batch(p, x, s)
torch.cuda.synchronize()
wall = time.time() - t0
stats = out["stats"].cpu().numpy()
verdict = out["verdict"].cpu().numpy()
os.makedirs(os.path.dirname(out_path), exist_ok=True)
np.savez_compressed(out_path, stats=stats, verdict=verdict, max_vio=out["max_vio"].cpu().numpy(), wall=wall)
it = stats[:, 2].astype(np.float64)
print(name, B, "wall %.3f s" % wall, "converged", int((verdict == 1).sum()),
      "admm iters mean %.0f median %.0f p99 %.0f max %.0f" % (it.mean(), np.median(it), np.percentile(it, 99), it.max()))
