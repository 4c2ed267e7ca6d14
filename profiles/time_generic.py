"""clock64() phase breakdown of the generic ADMM loops (sco_qp.cuh: generic_loop = strided shared-memory loop,
fast_loop = one thread per entity) on one penalty QP per problem; needs the cycle-counting build:
    SCO_BUILD_TAG=timing SCO_NVCC_FLAGS=-DSCO_TIMING python -m sco_py_b200.build
    SCO_B200_LIB=sco_py_b200/libsco_b200_timing.so python profiles/time_generic.py [config] [batch]"""
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
f, J, b, _ = eng.convexify(params, x0)
res = {}
for force in (1, 0):
    s = make_settings(solver=W.SOLVER_SETTINGS)
    s.force_generic = force
    for _ in range(2):
        xq, status, iters = eng.qp_solve(params, s, J=J, b=b, lbx=x0 - 1.0, ubx=x0 + 1.0, pi=np.full(B, 1.0), kdup=np.full(B, 1, np.int32))
    torch.cuda.synchronize()
    xq = xq.cpu().numpy(); it = iters.cpu().numpy().astype(float)
    res[force] = (status.cpu().numpy(), it)
    loop, setup, scale, chk = xq[:, 0], xq[:, 2], xq[:, 8], xq[:, 1]
    ph = [xq[:, 3 + k] for k in range(5)]
    print("%s B=%d team %d ctas/sm %d %s: iters mean %.0f | cycles/iter %.0f | setup %.0f (load + Ruiz %.0f, S + inverse %.0f) | per iteration: P1 %.0f P2 %.0f P3 %.0f P4 %.0f tests+rest %.0f (check() alone %.0f per test)" % (
        name, B, eng.team, eng.occupancy, "generic_loop" if force else "fast_loop", it.mean(), (loop / it).mean(), setup.mean(), scale.mean(),
        (setup - scale).mean(), *[(p / it).mean() for p in ph], (chk / np.maximum(it // 25, 1)).mean()))
print("same status:", bool((res[0][0] == res[1][0]).all()), " same iteration counts:", bool((res[0][1] == res[1][1]).all()))
