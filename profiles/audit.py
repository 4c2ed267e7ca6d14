"""Parity audit on a larger sample than the unit tests: the device solve of problems [0, N) of a
config against the CPU oracle port (one process per core) -- verdict agreement and the error
distributions SURVEY.md section 8(d) asks for.  TEST INFRASTRUCTURE (imports oracle/).
    python profiles/audit.py qcqp 256 [point_robot 64 arm 64 ...]
"""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np  # noqa: E402

from sco_py_b200 import workloads as W  # noqa: E402


def _cpu(job):
    name, i = job
    import sqp_port
    st, p, x = W.GENERATORS[name](1, first=i)
    r = sqp_port.solve(st, p[0], x[0], solver=W.SOLVER_SETTINGS)
    return i, r["x"], bool(r["success"]), r["objective"], r["max_vio"]


def main():
    import torch
    from sco_py_b200.engine import Engine, make_settings
    args = sys.argv[1:] or ["qcqp", "256"]
    report = {}
    for name, N in zip(args[0::2], [int(a) for a in args[1::2]]):
        st, params, x0 = W.gen_batch(name, N)
        eng = Engine(st)
        out = eng.solve_batch(params, x0, make_settings(solver=W.SOLVER_SETTINGS))
        torch.cuda.synchronize()
        x, v = out["x"].cpu().numpy(), out["verdict"].cpu().numpy()
        obj, vio = out["objective"].cpu().numpy(), out["max_vio"].cpu().numpy()
        t0 = time.time()
        with mp.get_context("fork").Pool(os.cpu_count()) as pool:
            res = pool.map(_cpu, [(name, i) for i in range(N)], chunksize=1)
        cpu_s = time.time() - t0
        match = np.array([(v[i] == 1) == ok for i, _, ok, _, _ in res])
        dx = np.array([np.abs(x[i] - xr).max() / max(1.0, np.abs(xr).max()) for i, xr, _, _, _ in res])
        dobj = np.array([abs(obj[i] - o) / max(1.0, abs(o)) for i, _, _, o, _ in res])
        dvio = np.array([abs(vio[i] - mv) for i, _, _, _, mv in res])
        m = match
        report[name] = dict(problems=N, verdict_match=float(match.mean()), converged_device=int((v == 1).sum()),
                            converged_oracle=int(sum(r[2] for r in res)),
                            rel_dx=dict(median=float(np.median(dx[m])), p90=float(np.quantile(dx[m], 0.9)),
                                        p99=float(np.quantile(dx[m], 0.99)), max=float(dx[m].max())),
                            rel_dobj_max=float(dobj[m].max()), abs_dvio_max=float(dvio[m].max()),
                            cpu_seconds=cpu_s, cpu_cores=os.cpu_count())
        print(name, json.dumps(report[name]))
        eng.close()
    with open(os.path.join(ROOT, "gpurun_out", "audit.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
