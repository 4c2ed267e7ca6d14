"""A tiny expression language for functions the device has no closed family for.

The reference takes arbitrary Python callables (`Expr(f)`, sco_py/expr.py:22-25); a kernel cannot run
Python, so a function is written once with the operators below and compiled to a stack program the
VM family of sco_families.cuh interprets on the device (and `eval_program` interprets on the host,
for the `Expr` behaviour of the object):

    x = sym.variables(2)
    f = (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2            # Rosenbrock, tests/sco_osqp/test_solver.py:99
    h = 10 * (x[1] - x[0] ** 2)
    SymExpr([f], n=2), SymExpr([h], n=2)                       # sco_py_b200.expr

Program encoding (doubles, so it travels inside the parameter block like any other family's
parameters): `m` row offsets (instruction index where row r starts), then instructions as
(opcode, operand) pairs.  The operand is a variable index (PUSH_X), the constant itself (PUSH_C) or
the exponent (POWI); every row ends with END.  Numbers may differ between the problems of a batch,
the opcodes may not.
"""
import math

import numpy as np

END, PUSH_X, PUSH_C, ADD, SUB, MUL, DIV, NEG, POWI, SQRT, LOG, EXP, SIN, COS = range(14)
MAX_STACK = 16


class Sym(object):
    __array_priority__ = 1000

    def __init__(self, op, args=(), value=None):
        self.op, self.args, self.value = op, args, value

    @staticmethod
    def wrap(v):
        return v if isinstance(v, Sym) else Sym(PUSH_C, value=float(v))

    def _bin(self, op, other, swap=False):
        a, b = (Sym.wrap(other), self) if swap else (self, Sym.wrap(other))
        return Sym(op, (a, b))

    def __add__(self, o): return self._bin(ADD, o)
    def __radd__(self, o): return self._bin(ADD, o, True)
    def __sub__(self, o): return self._bin(SUB, o)
    def __rsub__(self, o): return self._bin(SUB, o, True)
    def __mul__(self, o): return self._bin(MUL, o)
    def __rmul__(self, o): return self._bin(MUL, o, True)
    def __truediv__(self, o): return self._bin(DIV, o)
    def __rtruediv__(self, o): return self._bin(DIV, o, True)
    def __neg__(self): return Sym(NEG, (self,))

    def __pow__(self, p):
        if int(p) != p or p < 0 or p > 64:
            raise ValueError("only small non-negative integer powers are supported")
        return Sym(POWI, (self,), value=float(int(p)))

    def emit(self, out):
        for a in self.args:
            a.emit(out)
        out.append((self.op, self.value if self.value is not None else 0.0))

    def depth(self):
        """Stack slots needed to evaluate this node."""
        if not self.args:
            return 1
        d = [a.depth() for a in self.args]
        return max(d[0], d[1] + 1) if len(d) == 2 else d[0]


def variables(n):
    return [Sym(PUSH_X, value=float(j)) for j in range(n)]


def _fn(op):
    return lambda a: Sym(op, (Sym.wrap(a),))


sqrt, log, exp, sin, cos = _fn(SQRT), _fn(LOG), _fn(EXP), _fn(SIN), _fn(COS)


def compile_rows(rows):
    """-> (program as float64 array, number of instructions)."""
    code, offsets = [], []
    for r in rows:
        r = Sym.wrap(r)
        if r.depth() > MAX_STACK:
            raise ValueError("expression needs more than %d stack slots" % MAX_STACK)
        offsets.append(len(code))
        r.emit(code)
        code.append((END, 0.0))
    prog = np.empty(len(rows) + 2 * len(code))
    prog[:len(rows)] = offsets
    prog[len(rows):] = np.asarray(code, dtype=float).ravel()
    return prog, len(code)


def eval_program(prog, m, x):
    """Host interpreter: f(x) of the m rows, x flat.  Same operation order as the device VM."""
    x = np.asarray(x, dtype=float).ravel()
    ins = prog[m:].reshape(-1, 2)
    out = np.empty(m)
    for r in range(m):
        pc = int(prog[r])
        stk = []
        while True:
            op, arg = int(ins[pc, 0]), ins[pc, 1]
            pc += 1
            if op == END:
                break
            if op == PUSH_X:
                stk.append(x[int(arg)])
            elif op == PUSH_C:
                stk.append(arg)
            elif op in (ADD, SUB, MUL, DIV):
                b = stk.pop()
                a = stk.pop()
                stk.append(a + b if op == ADD else a - b if op == SUB else a * b if op == MUL else a / b)
            elif op == NEG:
                stk.append(-stk.pop())
            elif op == POWI:
                a = stk.pop()
                v = 1.0
                for _ in range(int(arg)):
                    v = v * a
                stk.append(v)
            else:
                a = stk.pop()
                stk.append({SQRT: math.sqrt, LOG: math.log, EXP: math.exp, SIN: math.sin, COS: math.cos}[op](a))
        out[r] = stk.pop()
    return out
