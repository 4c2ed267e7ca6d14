"""Mixed batches (BASELINE.json configs[4]): problems of different structures are bucketed by
structure signature; every bucket is one shared-structure batch = one kernel launch of its own engine,
buckets run concurrently on separate CUDA streams, results come back in the caller's order.

With torch.distributed initialised, bucket b's problems are split contiguously over the ranks
(shard.shard_range) like a plain batch -- no data-path collective.
"""
import numpy as np

from . import batch, shard


def bucket_by_signature(items):
    """items: list of (structure, params_row, x0_row).  -> {signature: (structure, [indices])}"""
    out = {}
    for i, (st, _, _) in enumerate(items):
        key = batch.signature(st)
        if key not in out:
            out[key] = (st, [])
        out[key][1].append(i)
    return out


def solve_mixed(items, settings, engines=None, device=0, rank=0, world=1):
    """Solves a list of (structure, params_row, x0_row) of arbitrary structures.
    Returns (x list, verdict[B], max_vio[B], stats[B,4], bucket report); entries of problems owned by
    other ranks are left at their initial values / -2."""
    import torch
    from .engine import Engine
    engines = {} if engines is None else engines
    B = len(items)
    xs = [np.asarray(it[2], dtype=float).copy() for it in items]
    verdict = np.full(B, -2, dtype=np.int32)
    vio = np.full(B, np.nan)
    stats = np.zeros((B, 4), dtype=np.int32)
    pending, report = [], {}
    for key, (st, idx) in bucket_by_signature(items).items():
        lo, hi = shard.shard_range(len(idx), rank, world)
        mine = idx[lo:hi]
        report[key] = dict(n=st.n, m_nl=st.m_nl, m_lin=st.m_lin, problems=len(idx), local=len(mine))
        if not mine:
            continue
        if key not in engines:
            engines[key] = Engine(st, device=device)
        params = np.stack([np.asarray(items[i][1], dtype=float) for i in mine])
        x0 = np.stack([np.asarray(items[i][2], dtype=float) for i in mine])
        stream = torch.cuda.Stream(device=engines[key].device)
        stream.wait_stream(torch.cuda.current_stream(engines[key].device))
        out = engines[key].solve_batch(params, x0, settings, stream=stream)
        pending.append((key, mine, out, stream))
    for key, mine, out, stream in pending:
        stream.synchronize()
        x = out["x"].cpu().numpy()
        v = out["verdict"].cpu().numpy()
        verdict[mine] = v
        vio[mine] = out["max_vio"].cpu().numpy()
        stats[mine] = out["stats"].cpu().numpy()
        for k, i in enumerate(mine):
            xs[i] = x[k]
        report[key]["converged"] = int((v == 1).sum())
        report[key]["team"] = engines[key].team
    return xs, verdict, vio, stats, report


def solve_bucketed(buckets, settings, engines=None, device=0):
    """Pre-bucketed form for large mixed batches (what workloads.gen_mixed emits): `buckets` is a list of
    dict(structure, params[B_k, stride], x0[B_k, n]) -- whole arrays per structure, nothing is stacked row by row.
    Every bucket is one launch on its own stream; returns the list of result dicts (host numpy) in bucket order."""
    import torch
    from .engine import Engine
    engines = {} if engines is None else engines
    pending = []
    for bk in buckets:
        key = batch.signature(bk["structure"])
        if key not in engines:
            engines[key] = Engine(bk["structure"], device=device)
        eng = engines[key]
        stream = torch.cuda.Stream(device=eng.device)
        stream.wait_stream(torch.cuda.current_stream(eng.device))
        pending.append((eng.solve_batch(bk["params"], bk["x0"], settings, stream=stream), stream))
    out = []
    for res, stream in pending:
        stream.synchronize()
        out.append({k: v.cpu().numpy() for k, v in res.items()})
    return out
