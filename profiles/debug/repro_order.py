"""Order-dependence / sanitizer repro: python profiles/debug/repro_order.py <config> <B> <steps...>
steps: s = full solve, c = closest-point QP, q = penalty QP stage, v = convexify; prints the stats of the solves."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sco_py_b200 import workloads as W
from sco_py_b200.engine import Engine, make_settings
name, B, steps = sys.argv[1], int(sys.argv[2]), sys.argv[3]
st, params, x0 = W.gen_batch(name, B)
eng = Engine(st)
s = make_settings(solver=W.SOLVER_SETTINGS)
if len(sys.argv) > 4:
    s.max_iter = int(sys.argv[4])
last = None
for ch in steps:
    if ch == "s":
        out = eng.solve_batch(params, x0, s)
        torch.cuda.synchronize()
        stats = out["stats"].cpu().numpy(); v = out["verdict"].cpu().numpy()
        print("solve: verdicts", np.bincount(v + 3, minlength=5).tolist(), "stats[0:3]", stats[:3].tolist(), "sum admm", int(stats[:, 2].sum()))
        if last is not None:
            print("   same as previous solve:", bool((stats == last[0]).all()), float(np.abs(out["x"].cpu().numpy() - last[1]).max()))
        last = (stats, out["x"].cpu().numpy())
    elif ch == "c":
        xq, status, iters = eng.qp_solve(params, s, xref=x0, use_penalty=False, closest_point=True)
        torch.cuda.synchronize(); print("closest: iters", iters[:4].tolist())
    elif ch == "v":
        f, J, b, _ = eng.convexify(params, x0); torch.cuda.synchronize(); print("convexify")
    elif ch == "q":
        f, J, b, _ = eng.convexify(params, x0)
        xq, status, iters = eng.qp_solve(params, s, J=J, b=b, lbx=x0 - 1.0, ubx=x0 + 1.0, pi=np.full(B, 1.0), kdup=np.full(B, 1, np.int32))
        torch.cuda.synchronize(); print("qp stage: iters", iters[:4].tolist(), "status", status[:4].tolist())
