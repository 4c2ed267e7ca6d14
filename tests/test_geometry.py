"""The device expression library (sco_py_b200/geometry.py) and forward-mode derivatives of the sym VM: programs against
plain NumPy formulas and finite differences (CPU), device Jacobians against the host's and the oracle's (GPU)."""
import math

import numpy as np
import pytest

import families  # oracle/families.py
from sco_py_b200 import families_host, geometry as G, sym
from sco_py_b200.expr import SymExpr


def _np_box(p, c, h):
    q = np.abs(np.asarray(p) - c) - h
    return np.linalg.norm(np.maximum(q, 0.0)) + min(q.max(), 0.0)


def _np_capsule(p, a, b, r):
    pa, ba = np.asarray(p) - a, np.asarray(b) - a
    h = min(max(pa @ ba / (ba @ ba), 0.0), 1.0)
    return np.linalg.norm(pa - ba * h) - r


def _fd(f, x, h=1e-6):
    return np.stack([(f(x + h * np.eye(x.size)[j]) - f(x - h * np.eye(x.size)[j])) / (2 * h) for j in range(x.size)], axis=1)


def test_signed_distances_match_numpy():
    x = sym.variables(3)
    c, h = np.array([0.5, -0.2, 0.1]), np.array([0.4, 0.3, 0.2])
    a, b = np.array([-1.0, 0.0, 0.3]), np.array([1.0, 0.5, 0.3])
    rows = [G.box_sdf(x, c, h), G.capsule_sdf(x, a, b, 0.15), G.sphere_sdf(x, c, 0.25),
            G.sphere_pair_clearance(x, 0.1, a, 0.2), G.box_sdf(x[:2], c[:2], h[:2])]
    e = SymExpr(rows, 3, analytic=True)
    rng = np.random.default_rng(1)
    for _ in range(40):
        p = rng.uniform(-1.2, 1.2, 3)
        ref = np.array([_np_box(p, c, h), _np_capsule(p, a, b, 0.15), np.linalg.norm(p - c) - 0.25,
                        np.linalg.norm(p - a) - 0.3, _np_box(p[:2], c[:2], h[:2])])
        assert np.allclose(e.eval(p.reshape(3, 1))[:, 0], ref, rtol=1e-13, atol=1e-13)
        J = e.grad(p.reshape(3, 1))
        Jfd = _fd(lambda v: e.eval(v.reshape(3, 1))[:, 0], p)
        assert np.allclose(J, Jfd, atol=2e-6)  # kinks aside (measure zero), the derivative is exact
        # the oracle's independent interpreter / differentiator agree
        assert np.array_equal(families.vm_f(p.reshape(3, 1), e.program, e.m)[:, 0], e.eval(p.reshape(3, 1))[:, 0])
        assert np.allclose(families.vm_grad(p.reshape(3, 1), e.program, e.m), J, rtol=1e-14, atol=1e-14)


def test_dh_chain_reproduces_the_closed_fk_family_and_shares_subexpressions():
    q = sym.variables(7)
    frames = G.dh_chain(q, families_host.FK7_A, families_host.FK7_D, families_host.FK7_ALPHA)
    flange = G.tool_position(frames[-1], (0.0, 0.0, families_host.FK7_FLANGE))
    R = frames[-1][0]
    e = SymExpr(flange + [R[0][0], R[1][2], R[2][1]], 7, analytic=True)
    assert e.n_instr < 2500  # a tree without temporaries needs > 10^5 instructions for seven links
    rng = np.random.default_rng(2)
    for _ in range(10):
        v = rng.uniform(-2.5, 2.5, 7)
        out = e.eval(v.reshape(7, 1))[:, 0]
        assert np.allclose(out[:3], families_host.fk7_pos(v), rtol=1e-13, atol=1e-14)
        J = e.grad(v.reshape(7, 1))
        Jfd = _fd(lambda u: e.eval(u.reshape(7, 1))[:, 0], v)
        assert np.allclose(J, Jfd, atol=1e-6)
    # orientation error vanishes at the desired orientation and is ~ axis * angle nearby
    v = rng.uniform(-1, 1, 7)
    Rd = [[float(sym.eval_program(sym.compile_rows([R[r][c]])[0], 1, v)[0]) for c in range(3)] for r in range(3)]
    err = SymExpr(G.rotation_error(R, Rd), 7)
    assert np.allclose(err.eval(v.reshape(7, 1)), 0.0, atol=1e-14)


def test_too_many_temporaries_is_an_error():
    x = sym.variables(2)
    terms = [sym.sin(x[0] * (k + 1)) for k in range(sym.MAX_SLOTS + 2)]
    row = sum((t * t for t in terms), sym.Sym.wrap(0.0)) + sum((t * 2.0 for t in terms), sym.Sym.wrap(0.0))
    with pytest.raises(ValueError, match="temporaries"):
        sym.compile_rows([row])


@pytest.mark.gpu
def test_device_jacobians_of_analytic_programs():
    """sco_convexify on a VM block: analytic=True gives the program's exact derivative (host forward mode to 1e-12),
    analytic=False the finite-difference one (host shim to 1e-7)."""
    import api_builder  # noqa: F401
    from sco_py_b200 import batch, expr as E
    from sco_py_b200.engine import Engine
    from sco_py_b200.sco_b200.osqp_utils import OSQPVar
    from sco_py_b200.sco_b200.prob import Prob
    from sco_py_b200.sco_b200.variable import Variable
    n = 6
    x = sym.variables(n)
    c, h = np.array([0.5, -0.2]), np.array([0.4, 0.3])
    rows = [0.05 - G.box_sdf((x[2 * t], x[2 * t + 1]), c, h) for t in range(3)] + \
           [0.1 - G.capsule_sdf((x[0], x[1], x[2]), (-1.0, 0.0, 0.3), (1.0, 0.5, 0.3), 0.15)]
    rng = np.random.default_rng(3)
    pts = rng.uniform(-1.5, 1.5, (8, n))
    for analytic in (True, False):
        probs = []
        for b in range(8):
            prob = Prob()
            ov = np.empty((n, 1), dtype=object)
            for j in range(n):
                ov[j, 0] = OSQPVar("x%05d" % j)
                prob.add_osqp_var(ov[j, 0])
            var = Variable(ov, pts[b].reshape(n, 1))
            prob.add_var(var)
            prob.add_obj_expr(E.BoundExpr(E.QuadExpr(2 * np.eye(n), np.zeros((1, n)), np.zeros((1, 1))), var))
            prob.add_cnt_expr(E.BoundExpr(E.LEqExpr(SymExpr(rows, n, analytic=analytic), np.zeros((4, 1))), var))
            probs.append(prob)
        st, params, x0, _ = batch.compile_batch(probs)
        assert st.blocks[0].ipar[3] == int(analytic)
        eng = Engine(st)
        f, J, bb, _ = [t.cpu().numpy() for t in eng.convexify(params, x0)]
        e = SymExpr(rows, n, analytic=True)
        for b in range(8):
            assert np.allclose(f[b], e.eval(pts[b].reshape(n, 1))[:, 0], rtol=1e-12, atol=1e-12)
            Jh = e.grad(pts[b].reshape(n, 1))
            assert np.allclose(J[b].reshape(4, n), Jh, atol=1e-12 if analytic else 1e-6)
        eng.close()
