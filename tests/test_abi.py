"""The C-ABI library loads and exports every symbol include/sco_b200.h declares (no GPU needed,
no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sco_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sco_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("sco_create", "sco_destroy", "sco_solve_batch", "sco_solve_batch_host", "sco_convexify",
                 "sco_qp_solve", "sco_merit", "sco_last_error", "sco_default_settings", "sco_query"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from sco_py_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), "libsco_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == _declared()


def test_default_settings_mirror_the_reference_defaults():
    # sco_py/sco_osqp/solver.py:17-28 and osqp_utils.py:10-15
    from sco_py_b200 import _lib
    lib = _lib.load()
    s = _lib.CSettings()
    lib.sco_default_settings(ctypes.byref(s))
    assert (s.improve_ratio_threshold, s.min_trust_region_size, s.min_approx_improve) == (0.25, 1e-4, 1e-8)
    assert (s.trust_shrink_ratio, s.trust_expand_ratio, s.cnt_tolerance) == (0.1, 1.5, 1e-4)
    assert (s.max_merit_coeff_increases, s.merit_coeff_increase_ratio) == (1, 10.0)
    assert (s.initial_trust_region_size, s.initial_penalty_coeff) == (1.0, 1e3)
    assert (s.osqp_eps_abs, s.osqp_eps_rel, s.osqp_rho, s.osqp_sigma) == (1e-6, 1e-9, 0.1, 5e-10)
    assert (s.osqp_max_iter, s.osqp_adaptive_rho, s.osqp_alpha) == (100000, 0, 1.6)
    assert (s.compound_penalty, s.freeze_sparsity, s.duplicate_rows) == (1, 1, 1)


def test_engine_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sco_py_b200 import workloads as W
    from sco_py_b200.engine import Engine
    st, _, _ = W.gen_qcqp(1)
    with pytest.raises(RuntimeError):
        Engine(st)
    # and the C ABI itself refuses too (no CPU fallback behind the boundary)
    from sco_py_b200 import _lib
    lib = _lib.load()
    cs = _lib.CStructure()
    cs.n, cs.n_groups = 2, 1
    h = ctypes.c_void_p()
    rc = lib.sco_create(ctypes.byref(cs), 0, ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in lib.sco_last_error()


def test_c_example_compiles_links_and_refuses_to_run_without_a_gpu():
    """examples/solve_qcqp.c: a C host needs nothing but include/sco_b200.h and the shared library; on a
    machine without a CUDA device sco_create fails with a message (no CPU fallback)."""
    import shutil
    import subprocess
    from sco_py_b200 import _lib, build
    build.build()
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = os.path.join(ROOT, "examples", "solve_qcqp")
    cmd = ["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "solve_qcqp.c"), "-o", exe,
           _lib.LIB_PATH, "-lm", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)]
    subprocess.check_call(cmd)
    import torch
    if not torch.cuda.is_available():
        p = subprocess.run([exe, "4"], capture_output=True, text=True)
        assert p.returncode == 1 and "no CUDA device" in p.stderr


@pytest.mark.gpu
def test_c_example_runs():
    import subprocess
    from sco_py_b200 import _lib
    exe = os.path.join(ROOT, "examples", "solve_qcqp")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "solve_qcqp.c"),
                           "-o", exe, _lib.LIB_PATH, "-lm", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)])
    p = subprocess.run([exe, "64"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "converged" in p.stdout
