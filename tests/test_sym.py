"""The expression language behind SymExpr (sco_py_b200/sym.py): compilation to stack programs, the
host interpreter, and the independent interpreter of the oracle on the same programs."""
import math

import numpy as np
import pytest

import families  # oracle/families.py
from sco_py_b200 import sym
from sco_py_b200.expr import SymExpr


def test_programs_evaluate_like_python():
    x = sym.variables(3)
    rows = [(x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2,
            sym.log(1 + x[0] ** 2) - x[1] / (2 + sym.cos(x[2])),
            -sym.sqrt(x[0] * x[0] + 4) + sym.exp(0.1 * x[2]) * sym.sin(x[1]) + 3,
            sym.Sym.wrap(-1e5)]
    prog, n_instr = sym.compile_rows(rows)
    assert prog.size == len(rows) + 2 * n_instr
    rng = np.random.default_rng(0)
    for _ in range(5):
        v = rng.normal(size=3)
        ref = [(v[1] - v[0] ** 2) ** 2 + (1 - v[0]) ** 2,
               math.log(1 + v[0] ** 2) - v[1] / (2 + math.cos(v[2])),
               -math.sqrt(v[0] * v[0] + 4) + math.exp(0.1 * v[2]) * math.sin(v[1]) + 3,
               -1e5]
        got = sym.eval_program(prog, len(rows), v)
        assert np.allclose(got, ref, rtol=1e-14, atol=1e-14)
        # the oracle's own interpreter agrees bit for bit (same operation order)
        assert np.array_equal(families.vm_f(v.reshape(3, 1), prog, len(rows))[:, 0], got)


def test_integer_powers_are_repeated_products():
    x = sym.variables(1)
    prog, _ = sym.compile_rows([x[0] ** 0, x[0] ** 1, x[0] ** 5])
    v = np.array([1.7])
    assert np.array_equal(sym.eval_program(prog, 3, v), [1.0, 1.7, 1.7 * 1.7 * 1.7 * 1.7 * 1.7])
    with pytest.raises(ValueError):
        x[0] ** 0.5
    with pytest.raises(ValueError):
        x[0] ** -1


def test_stack_depth_is_checked():
    x = sym.variables(1)
    chain = x[0]
    for _ in range(40):  # left-nested chains stay shallow
        chain = chain * x[0] + 1.0
    assert chain.depth() == 2
    sym.compile_rows([chain])
    nested = x[0]
    for _ in range(sym.MAX_STACK + 1):  # right-nested ones need one more slot per level
        nested = sym.Sym(sym.ADD, (sym.Sym.wrap(1.0), nested))
    assert nested.depth() > sym.MAX_STACK
    with pytest.raises(ValueError):
        sym.compile_rows([nested])


def test_symexpr_is_an_expr_without_analytic_derivatives():
    x = sym.variables(2)
    e = SymExpr([x[0] ** 2 * x[1], x[0] + 3 * x[1]], 2)
    pt = np.array([[1.5], [-2.0]])
    assert np.allclose(e.eval(pt)[:, 0], [1.5 ** 2 * -2.0, 1.5 - 6.0])
    J = e.grad(pt)
    assert np.allclose(J, [[2 * 1.5 * -2.0, 1.5 ** 2], [1.0, 3.0]], atol=1e-7)
    scalar = SymExpr([x[0] ** 4 + x[1] ** 4], 2)
    H = scalar.hess(pt)
    assert np.allclose(H, np.diag([12 * 1.5 ** 2, 12 * 4.0]), atol=1e-4)
    q = scalar.convexify(pt, degree=2)
    assert abs(q.eval(pt)[0, 0] - scalar.eval(pt)[0, 0]) <= 1e-6
    assert (e.family, e.m, e.jw, e.ipar[:3]) == (4, 2, 2, [2, 2, e.n_instr])
