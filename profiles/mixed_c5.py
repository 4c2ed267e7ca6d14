"""BASELINE.json configs[4] in miniature: an i.i.d. mixture of problem shapes (SURVEY.md section 8d, C5:
50 % QCQP-shaped n in {10,20,30} / m in {15,30,45}, 30 % point-robot-shaped T in {20,40} / K in {1,3},
20 % arm-shaped T in {10,20}; problem i drawn from default_rng(5000 + i)), bucketed by structure
signature and solved bucket by bucket on concurrent streams (sco_py_b200/buckets.py).
    python profiles/mixed_c5.py [problems]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sco_py_b200 import buckets  # noqa: E402
from sco_py_b200 import workloads as W  # noqa: E402
from sco_py_b200.engine import make_settings  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
t0 = time.time()
items = []
for i in range(N):
    rng = np.random.default_rng(5000 + i)
    u = rng.uniform()
    if u < 0.5:
        k = int(rng.integers(3))
        st, p, x = W.gen_qcqp(1, n=(10, 20, 30)[k], m=(15, 30, 45)[k], seed_base=5000, first=i)
    elif u < 0.8:
        st, p, x = W.gen_point_robot(1, T=(20, 40)[int(rng.integers(2))], K=(1, 3)[int(rng.integers(2))], seed_base=5000, first=i)
    else:
        st, p, x = W.gen_arm(1, T=(10, 20)[int(rng.integers(2))], seed_base=5000, first=i)
    items.append((st, p[0], x[0]))
gen_s = time.time() - t0
settings = make_settings(solver=W.SOLVER_SETTINGS)
engines = {}
buckets.solve_mixed(items[:64], settings, engines=engines)  # warm-up: engines, allocations
torch.cuda.synchronize()
t0 = time.time()
xs, verdict, vio, stats, report = buckets.solve_mixed(items, settings, engines=engines)
wall = time.time() - t0
rows = sorted(report.values(), key=lambda r: (-r["problems"]))
print("%d problems in %d buckets: %.2f s (%.0f problems/s incl. host staging), %d converged, generation %.1f s" % (
    N, len(rows), wall, N / wall, int((verdict == 1).sum()), gen_s))
for r in rows:
    print("  n=%-4d m_nl=%-4d m_lin=%-4d team=%-4d problems=%-5d converged=%d" % (
        r["n"], r["m_nl"], r["m_lin"], r.get("team", 0), r["problems"], r.get("converged", 0)))
json.dump(dict(problems=N, wall_s=wall, converged=int((verdict == 1).sum()), buckets=rows),
          open(os.path.join(ROOT, "gpurun_out", "mixed_c5.json"), "w"), indent=1)
