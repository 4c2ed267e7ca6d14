"""On-disk batch format (sco_py_b200/batchfile.py): round trip of every structure kind on CPU; on the GPU a
file written with recorded results solves to the same results."""
import numpy as np
import pytest

import api_builder
from sco_py_b200 import batch, batchfile
from sco_py_b200 import workloads as W


def _same_structure(a, b):
    assert batchfile.structure_to_json(a) == batchfile.structure_to_json(b)
    for name in ("lin_rowptr", "lin_col", "lin_val", "shared", "group_overlap"):
        x, y = getattr(a, name), getattr(b, name)
        assert (x is None) == (y is None)
        if x is not None:
            assert np.array_equal(np.asarray(x), np.asarray(y))
    assert batch.signature(a) == batch.signature(b)


@pytest.mark.parametrize("name", ["qcqp", "point_robot", "arm"])
def test_round_trip_of_the_benchmark_structures(tmp_path, name):
    st, params, x0 = W.GENERATORS[name](3)
    path = str(tmp_path / ("%s.npz" % name))
    batchfile.save_batch(path, st, params, x0, settings=dict(solver=W.SOLVER_SETTINGS),
                         expected=dict(verdict=np.array([1, 0, 1], np.int32)))
    st2, p2, x2, meta = batchfile.load_batch(path)
    _same_structure(st, st2)
    assert np.array_equal(params, p2) and np.array_equal(x0, x2)
    assert meta["settings"]["solver"] == W.SOLVER_SETTINGS and meta["expected"]["verdict"].tolist() == [1, 0, 1]


def test_round_trip_with_groups_affine_objective_and_bounds(tmp_path):
    groups = [(slice(0, 2), ["a", "b"]), (slice(2, 4), ["b"]), (slice(4, 6), [])]
    probs = [api_builder.build_qcqp_variant(i, aff=True, bounds=True, groups=groups)[0] for i in range(2)]
    st, params, x0, _ = batch.compile_batch(probs)
    path = str(tmp_path / "v.npz")
    batchfile.save_batch(path, st, params, x0)
    st2, p2, x2, meta = batchfile.load_batch(path)
    _same_structure(st, st2)
    assert meta["settings"] is None and meta["expected"] == {}


def test_damaged_file_is_refused(tmp_path):
    st, params, x0 = W.gen_qcqp(2, n=4, m=3)
    path = str(tmp_path / "q.npz")
    batchfile.save_batch(path, st, params, x0)
    z = dict(np.load(path))
    z["signature"] = np.array("0" * 64)
    np.savez_compressed(path, **z)
    with pytest.raises(ValueError, match="signature"):
        batchfile.load_batch(path)
    with pytest.raises(ValueError):
        batchfile.save_batch(path, st, params[:, :-1], x0)


@pytest.mark.gpu
def test_recorded_results_reproduce(tmp_path):
    from sco_py_b200.engine import Engine, make_settings
    st, params, x0 = W.gen_qcqp(32, n=8, m=6)
    eng = Engine(st)
    out = eng.solve_batch_host(params, x0, make_settings(solver=W.SOLVER_SETTINGS))
    path = str(tmp_path / "r.npz")
    batchfile.save_batch(path, st, params, x0, settings=dict(solver=W.SOLVER_SETTINGS),
                         expected=dict(verdict=out["verdict"], x=out["x"], max_vio=out["max_vio"]))
    st2, p2, x2, meta = batchfile.load_batch(path)
    eng2 = Engine(st2)
    again = eng2.solve_batch_host(p2, x2, make_settings(solver=meta["settings"]["solver"]))
    assert np.array_equal(again["verdict"], meta["expected"]["verdict"]) and np.array_equal(again["x"], meta["expected"]["x"])
