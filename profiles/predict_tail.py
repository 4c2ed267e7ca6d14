"""Does the ADMM iteration count of the FIRST penalty QP predict the total work of a problem?
(experiment behind the longest-first scheduling of k_solve)
    python profiles/predict_tail.py [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sco_py_b200 import workloads as W  # noqa: E402
from sco_py_b200.engine import Engine, make_settings  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
st, params, x0 = W.gen_batch("qcqp", B)
eng = Engine(st)
s = make_settings(solver=W.SOLVER_SETTINGS)
p, x = eng._dev(params), eng._dev(x0)
f, J, b, _ = eng.convexify(p, x)
feats = {}
for name, pi, kd, delta in (("pi1", 1.0, 1, 1.0), ("pi10", 10.0, 2, 0.1), ("pi1e3", 1e3, 3, 1.0)):
    xq, status, iters = eng.qp_solve(p, s, J=J, b=b, lbx=x - delta, ubx=x + delta, pi=torch.full((B,), pi, dtype=torch.float64),
                                     kdup=torch.full((B,), kd, dtype=torch.int32))
    feats[name] = iters.cpu().numpy().astype(float)
vio = torch.clamp(f - eng._dev(params[:, st.blocks[0].val.off:st.blocks[0].val.off + st.m_nl]), min=0).sum(dim=1).cpu().numpy()
feats["vio0"] = vio
out = eng.solve_batch(p, x, s)
torch.cuda.synchronize()
tot = out["stats"].cpu().numpy()[:, 2].astype(float)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "predict_tail.npz"), tot=tot, **feats)
order = np.argsort(-tot)
for name, v in feats.items():
    r = np.argsort(np.argsort(-v))  # rank of each problem by the feature (0 = largest)
    top = order[:32]
    print("%-6s spearman %.3f | rank percentile of the 32 longest problems: median %.1f%% worst %.1f%%" % (
        name, np.corrcoef(np.argsort(np.argsort(v)), np.argsort(np.argsort(tot)))[0, 1],
        100 * np.median(r[top]) / B, 100 * r[top].max() / B))
