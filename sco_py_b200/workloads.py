"""Synthetic workloads of BASELINE.json `configs` (generators fixed by SURVEY.md section 8d).

Each generator returns `(structure, params[B, stride], x0[B, n])`, host NumPy,
float64.  Problem i of a config draws from `np.random.default_rng(seed_base + i)`
so any sub-batch is reproducible on its own (the CPU baseline solves a sample of
the very same problems the GPU solves).

  C2  point robot   T=40, d=2  -> n=80,  120 hinge rows (CIRCLE2D), 4 linear eq rows
  C3  7-DOF arm     T=20       -> n=140, 3 abs rows (FK7, FD Jacobian), 287 linear rows
  C4  QCQP          n=20, m=30 -> 30 hinge rows (QUADFORM)

`build_reference_prob` turns one problem of a batch into reference-style objects
(Variable / Prob / BoundExpr ...) using whichever API module namespace is
passed in -- the reference's own `sco_py` or this package's drop-in -- so the
very same builder drives both sides of a parity test.
"""
import numpy as np

from .structure import (CNT_EQ, CNT_LEQ, FAM_CIRCLE2D, FAM_FK7, FAM_QUADFORM, Block, Field,
                        Structure)

SOLVER_SETTINGS = dict(  # tests/sco_osqp/test_solver.py:15-25
    improve_ratio_threshold=0.25, min_trust_region_size=1e-5, min_approx_improve=1e-8,
    max_iter=50, trust_shrink_ratio=0.1, trust_expand_ratio=1.5, cnt_tolerance=1e-4,
    max_merit_coeff_increases=5, merit_coeff_increase_ratio=10.0,
    initial_trust_region_size=1.0, initial_penalty_coeff=1.0)


def smoothness_Q(T, d):
    """Q = 2 D'D for the (T-1)d x Td first-difference matrix D (sum ||p_{t+1}-p_t||^2)."""
    n = T * d
    D = np.zeros(((T - 1) * d, n))
    for t in range(T - 1):
        for k in range(d):
            D[t * d + k, t * d + k] = -1.0
            D[t * d + k, (t + 1) * d + k] = 1.0
    return 2.0 * D.T @ D


def _csr(A):
    rowptr = [0]
    col = []
    val = []
    for r in range(A.shape[0]):
        (idx,) = np.nonzero(A[r])
        col.extend(idx.tolist())
        val.extend(A[r, idx].tolist())
        rowptr.append(len(col))
    return (np.asarray(rowptr, np.int32), np.asarray(col, np.int32), np.asarray(val, np.float64))


# ------------------------------------------------------------------ C4: QCQP
def qcqp_structure(n=20, m=30):
    ntri = n * (n + 1) // 2
    off = 0
    Q = Field(off, False); off += n * n
    q = Field(off, False); off += n
    par = Field(off, False); off += m * ntri + m * n
    val = Field(off, False); off += m
    blk = Block(FAM_QUADFORM, CNT_LEQ, m, par, val, ipar=[n, m, 0, 0, 0, 0, 0, 0], jw=n)
    return Structure(n=n, stride=off, Q=Q, q=q, c=Field(-1, False), blocks=[blk])


def _indices(B, first, indices):
    return np.arange(first, first + B) if indices is None else np.asarray(indices, dtype=np.int64)


def gen_qcqp(B, n=20, m=30, seed_base=4000, first=0, indices=None):
    indices = _indices(B, first, indices)
    B = len(indices)
    st = qcqp_structure(n, m)
    ntri = n * (n + 1) // 2
    params = np.empty((B, st.stride))
    x0 = np.empty((B, n))
    iu = np.triu_indices(n)
    for b in range(B):
        rng = np.random.default_rng(seed_base + int(indices[b]))
        M = rng.standard_normal((n, n))
        Qm = M.T @ M / n + 0.1 * np.eye(n)
        qv = rng.standard_normal(n)
        S = rng.standard_normal((m, n, n)) / np.sqrt(n)
        Pm = 0.5 * (S + np.transpose(S, (0, 2, 1)))
        a = rng.standard_normal((m, n))
        bv = rng.uniform(0.5, 1.5, m)
        x0[b] = rng.standard_normal(n)
        row = params[b]
        row[st.Q.off:st.Q.off + n * n] = Qm.ravel()
        row[st.q.off:st.q.off + n] = qv
        po = st.blocks[0].par.off
        row[po:po + m * ntri] = Pm[:, iu[0], iu[1]].ravel()
        row[po + m * ntri:po + m * ntri + m * n] = a.ravel()
        row[st.blocks[0].val.off:st.blocks[0].val.off + m] = bv
    return st, params, x0


# ------------------------------------------------------------------ C2: point robot
def point_robot_structure(T=40, K=3):
    n = 2 * T
    shared = smoothness_Q(T, 2).ravel()
    A = np.zeros((4, n))
    A[0, 0] = A[1, 1] = 1.0
    A[2, n - 2] = A[3, n - 1] = 1.0
    rp, ci, cv = _csr(A)
    off = 0
    lin_l = Field(off, False); off += 4
    lin_u = Field(off, False); off += 4
    par = Field(off, False); off += 3 * K
    blk = Block(FAM_CIRCLE2D, CNT_LEQ, T * K, par, Field(-1, False),
                ipar=[T, K, 0, 0, 0, 0, 0, 0], jw=2)
    return Structure(n=n, stride=off, Q=Field(0, True), q=Field(-1, False), c=Field(-1, False),
                     m_lin=4, lin_rowptr=rp, lin_col=ci, lin_val=cv, lin_l=lin_l, lin_u=lin_u,
                     blocks=[blk], shared=shared)


def gen_point_robot(B, T=40, K=3, seed_base=2000, first=0, indices=None):
    indices = _indices(B, first, indices)
    B = len(indices)
    st = point_robot_structure(T, K)
    n = st.n
    params = np.zeros((B, st.stride))
    x0 = np.empty((B, n))
    for b in range(B):
        rng = np.random.default_rng(seed_base + int(indices[b]))
        goal = np.array([10.0, 0.0]) + rng.normal(0.0, 0.5, 2)
        cx = rng.uniform(2.0, 8.0, K)
        cy = rng.normal(0.0, 0.3, K)
        R = rng.uniform(0.5, 1.0, K) + 0.1
        line = np.linspace(0.0, 1.0, T)[:, None] * goal[None, :]
        x0[b] = (line + rng.normal(0.0, 0.05, (T, 2))).ravel()
        row = params[b]
        pin = np.array([0.0, 0.0, goal[0], goal[1]])
        row[st.lin_l.off:st.lin_l.off + 4] = pin
        row[st.lin_u.off:st.lin_u.off + 4] = pin
        po = st.blocks[0].par.off
        row[po:po + 2 * K] = np.stack([cx, cy], axis=1).ravel()
        row[po + 2 * K:po + 3 * K] = R
    return st, params, x0


# ------------------------------------------------------------------ C3: 7-DOF arm
JOINT_LIMIT = 2.9


def arm_structure(T=20):
    from .families_host import fk7_table
    n = 7 * T
    Qs = np.concatenate([smoothness_Q(T, 7).ravel(), fk7_table()])
    # joint limits: x <= 2.9 and -x <= 2.9 (two LEqExpr(AffExpr(+-I))), then the start pin
    A = np.vstack([np.eye(n), -np.eye(n), np.eye(7, n)])
    rp, ci, cv = _csr(A)
    m_lin = 2 * n + 7
    # shared block: Q, then the shared halves of lin_l / lin_u are per problem (start pin differs)
    off = 0
    lin_l = Field(off, False); off += m_lin
    lin_u = Field(off, False); off += m_lin
    val = Field(off, False); off += 3
    blk = Block(FAM_FK7, CNT_EQ, 3, Field(n * n, True), val, ipar=[T, 0, 0, 0, 0, 0, 0, 0], jw=7)
    return Structure(n=n, stride=off, Q=Field(0, True), q=Field(-1, False), c=Field(-1, False),
                     m_lin=m_lin, lin_rowptr=rp, lin_col=ci, lin_val=cv, lin_l=lin_l, lin_u=lin_u,
                     blocks=[blk], shared=Qs)


def gen_arm(B, T=20, seed_base=3000, first=0, fk=None, indices=None):
    if fk is None:
        from .families_host import fk7_pos as fk
    indices = _indices(B, first, indices)
    B = len(indices)
    st = arm_structure(T)
    n = st.n
    params = np.zeros((B, st.stride))
    x0 = np.empty((B, n))
    for b in range(B):
        rng = np.random.default_rng(seed_base + int(indices[b]))
        q0 = rng.uniform(-0.5, 0.5, 7)
        x0[b] = np.tile(q0, T)
        target = fk(q0 + rng.uniform(-0.6, 0.6, 7))
        row = params[b]
        lo = np.concatenate([np.full(2 * n, -np.inf), q0])
        hi = np.concatenate([np.full(2 * n, JOINT_LIMIT), q0])
        row[st.lin_l.off:st.lin_l.off + st.m_lin] = lo
        row[st.lin_u.off:st.lin_u.off + st.m_lin] = hi
        row[st.blocks[0].val.off:st.blocks[0].val.off + 3] = target
    return st, params, x0


GENERATORS = {"qcqp": gen_qcqp, "point_robot": gen_point_robot, "arm": gen_arm}


def _gen_chunk(args):
    name, first, count, kw = args
    _, params, x0 = GENERATORS[name](count, first=first, **kw)
    return first, params, x0


# ------------------------------------------------------------------ C5: mixed shapes (BASELINE.json configs[4])
# i.i.d. mixture, problem i drawn from default_rng(5000 + i): 50 % C4-shaped (n, m) in
# {(10,15), (20,30), (30,45)}, 30 % C2-shaped T in {20,40} x K in {1,3}, 20 % C3-shaped T in {10,20}.
MIXED_SEED = 5000
MIXED_BUCKETS = ([("qcqp", dict(n=n, m=m)) for n, m in ((10, 15), (20, 30), (30, 45))] +
                 [("point_robot", dict(T=T, K=K)) for T in (20, 40) for K in (1, 3)] +
                 [("arm", dict(T=T)) for T in (10, 20)])


def mixed_bucket_of(i):
    """Bucket (index into MIXED_BUCKETS) of problem i of the mixed workload."""
    rng = np.random.default_rng(MIXED_SEED + int(i))
    u = rng.uniform()
    if u < 0.5:
        return int(rng.integers(3))
    if u < 0.8:
        return 3 + 2 * int(rng.integers(2)) + int(rng.integers(2))
    return 7 + int(rng.integers(2))


def _bucket_chunk(args):
    lo, hi = args
    return lo, np.fromiter((mixed_bucket_of(i) for i in range(lo, hi)), dtype=np.int8, count=hi - lo)


def _gen_indexed_chunk(args):
    name, kw, indices, pos = args
    _, params, x0 = GENERATORS[name](len(indices), seed_base=MIXED_SEED, indices=indices, **kw)
    return pos, params, x0


def gen_mixed(B, first=0, workers=None, pinned=False):
    """Problems [first, first + B) of the mixed workload, bucketed by structure.
    -> list of dict(bucket, name, kw, structure, indices (global problem numbers), params, x0), non-empty buckets only.
    The bucket of every problem and the problems of every bucket are computed by a process pool in index chunks --
    nothing is stacked row by row in Python."""
    import multiprocessing as mp
    import os
    workers = workers or min(os.cpu_count() or 1, 32)
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    mapper = pool.imap_unordered if pool else map
    step = max(1024, (B + 8 * workers - 1) // (8 * workers))
    bucket = np.empty(B, dtype=np.int8)
    for lo, arr in mapper(_bucket_chunk, [(first + s, min(first + s + step, first + B)) for s in range(0, B, step)]):
        bucket[lo - first:lo - first + len(arr)] = arr
    out = []
    for bi, (name, kw) in enumerate(MIXED_BUCKETS):
        idx = first + np.nonzero(bucket == bi)[0]
        if idx.size == 0:
            continue
        st = GENERATORS[name](1, **kw)[0]
        if pinned:
            import torch
            params = torch.empty((idx.size, st.stride), dtype=torch.float64, pin_memory=True).numpy()
            x0 = torch.empty((idx.size, st.n), dtype=torch.float64, pin_memory=True).numpy()
        else:
            params, x0 = np.empty((idx.size, st.stride)), np.empty((idx.size, st.n))
        chunk = max(64, (idx.size + 4 * workers - 1) // (4 * workers))
        jobs = [(name, kw, idx[s:s + chunk], s) for s in range(0, idx.size, chunk)]
        for pos, p, x in mapper(_gen_indexed_chunk, jobs):
            params[pos:pos + p.shape[0]] = p
            x0[pos:pos + p.shape[0]] = x
        out.append(dict(bucket=bi, name=name, kw=kw, structure=st, indices=idx, params=params, x0=x0))
    if pool:
        pool.close()
        pool.join()
    return out


def gen_batch(name, B, first=0, workers=None, out_params=None, out_x0=None, **kw):
    """Problems [first, first+B) of a config, generated by a process pool (per-problem seeds make
    chunks independent).  `out_params` / `out_x0` may be preallocated (e.g. pinned) arrays."""
    import multiprocessing as mp
    import os
    st = GENERATORS[name](1, first=first, **kw)[0]
    params = out_params if out_params is not None else np.empty((B, st.stride))
    x0 = out_x0 if out_x0 is not None else np.empty((B, st.n))
    workers = workers or min(os.cpu_count() or 1, 32)
    chunk = max(64, (B + 4 * workers - 1) // (4 * workers))
    jobs = [(name, first + s, min(chunk, B - s), kw) for s in range(0, B, chunk)]
    if workers <= 1 or len(jobs) == 1:
        results = map(_gen_chunk, jobs)
        pool = None
    else:
        pool = mp.get_context("fork").Pool(workers)
        results = pool.imap_unordered(_gen_chunk, jobs)
    for f, p, x in results:
        params[f - first:f - first + p.shape[0]] = p
        x0[f - first:f - first + p.shape[0]] = x
    if pool is not None:
        pool.close()
        pool.join()
    return st, params, x0
