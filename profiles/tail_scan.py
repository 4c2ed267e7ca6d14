"""Heavy tail of the C4 workload across the eight 65,536-problem shards the 8-GPU bench solves:
per shard the wall time of one launch and the largest per-problem ADMM iteration counts.
    python profiles/tail_scan.py [shards] [batch]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sco_py_b200 import workloads as W  # noqa: E402
from sco_py_b200.engine import Engine, make_settings  # noqa: E402

shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
eng = None
rows = []
for r in range(shards):
    st, params, x0 = W.gen_batch("qcqp", B, first=r * B)
    if eng is None:
        eng = Engine(st)
        s = make_settings(solver=W.SOLVER_SETTINGS)
    p, x = eng._dev(params), eng._dev(x0)
    torch.cuda.synchronize()
    t0 = time.time()
    out = eng.solve_batch(p, x, s)
    torch.cuda.synchronize()
    wall = time.time() - t0
    it = out["stats"][:, 2].cpu().numpy().astype(np.int64)
    top = np.argsort(-it)[:3]
    rows.append(dict(shard=r, wall_s=wall, total_iters=int(it.sum()), top=[(int(r * B + i), int(it[i])) for i in top]))
    print(json.dumps(rows[-1]))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tail_scan.json"), "w"), indent=1)
