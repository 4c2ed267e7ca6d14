"""Mixed-structure batches (BASELINE.json configs[4]): bucketing on CPU, the bucketed solve on the GPU."""
import numpy as np
import pytest

import sqp_port
from sco_py_b200 import buckets
from sco_py_b200 import workloads as W

SHAPES = ((10, 15, 3), (20, 30, 3), (8, 6, 2), (24, 32, 2), (30, 45, 2))


def _items():
    items = []
    for n, m, count in SHAPES:
        st, p, x = W.gen_qcqp(count, n=n, m=m, seed_base=5000)
        items += [(st, p[i], x[i]) for i in range(count)]
    st, p, x = W.gen_point_robot(2, T=20, K=1, seed_base=5000)
    items += [(st, p[i], x[i]) for i in range(2)]
    rng = np.random.default_rng(0)
    order = rng.permutation(len(items))
    return [items[i] for i in order]


def test_bucketing_groups_equal_structures():
    items = _items()
    b = buckets.bucket_by_signature(items)
    assert len(b) == len(SHAPES) + 1
    assert sorted(i for _, idx in b.values() for i in idx) == list(range(len(items)))
    for st, idx in b.values():
        assert all(items[i][0].n == st.n and items[i][0].m_nl == st.m_nl for i in idx)


@pytest.mark.gpu
def test_mixed_batch_matches_the_oracle_per_problem():
    from sco_py_b200.engine import make_settings
    items = _items()
    xs, verdict, vio, stats, report = buckets.solve_mixed(items, make_settings(solver=W.SOLVER_SETTINGS))
    assert sum(r["problems"] for r in report.values()) == len(items)
    teams = {r["team"] for r in report.values()}
    assert 64 in teams  # dense two-warp kinds (n <= 32, m <= 32) ...
    assert len(teams) > 1  # ... and generic teams for the larger / structured ones
    for i, (st, row, x0) in enumerate(items):
        ref = sqp_port.solve(st, row, x0, solver=W.SOLVER_SETTINGS)
        assert (verdict[i] == 1) == ref["success"], (i, st.n, st.m_nl)
        assert np.abs(xs[i] - ref["x"]).max() <= 1e-4 * max(1.0, np.abs(ref["x"]).max()), (i, st.n, st.m_nl)
        assert abs(vio[i] - ref["max_vio"]) <= 1e-5


def test_mixed_generator_buckets_every_problem_once():
    """workloads.gen_mixed (BASELINE.json configs[4]): vectorised bucketing, reproducible per problem."""
    bks = W.gen_mixed(600, first=1000, workers=2)
    idx = np.concatenate([b["indices"] for b in bks])
    assert sorted(idx.tolist()) == list(range(1000, 1600))
    for b in bks:
        assert all(W.mixed_bucket_of(i) == b["bucket"] for i in b["indices"][:5])
        st, p, x = W.GENERATORS[b["name"]](1, seed_base=W.MIXED_SEED, indices=[int(b["indices"][0])], **b["kw"])
        assert np.array_equal(p[0], b["params"][0]) and np.array_equal(x[0], b["x0"][0])
    again = W.gen_mixed(600, first=1000, workers=1)
    assert all(np.array_equal(a["params"], b["params"]) for a, b in zip(bks, again))


@pytest.mark.gpu
def test_bucketed_solve_matches_per_bucket_solves():
    from sco_py_b200.engine import Engine, make_settings
    bks = W.gen_mixed(96, workers=2)
    s = make_settings(solver=W.SOLVER_SETTINGS)
    outs = buckets.solve_bucketed(bks, s)
    for b, o in zip(bks, outs):
        ref = Engine(b["structure"]).solve_batch_host(b["params"], b["x0"], s)
        assert np.array_equal(ref["verdict"], o["verdict"]) and np.array_equal(ref["x"], o["x"])
