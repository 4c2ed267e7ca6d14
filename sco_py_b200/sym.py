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

Derivatives: by default a SymExpr is a black box like `Expr(f)` without grad (finite differences, expr.py:61-69); with
`SymExpr(rows, n, analytic=True)` it is the counterpart of `Expr(f, grad)` (expr.py:86-88): the VM differentiates the
program itself in forward mode -- exact, and one pass per variable instead of four function evaluations.
"""
import math

import numpy as np

END, PUSH_X, PUSH_C, ADD, SUB, MUL, DIV, NEG, POWI, SQRT, LOG, EXP, SIN, COS, ABS, MIN, MAX, TEE, LOAD = range(19)
MAX_STACK = 16
MAX_SLOTS = 128  # temporaries of a program: TEE k copies the top of the stack into slot k, LOAD k pushes it


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
    def __abs__(self): return Sym(ABS, (self,))

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


sqrt, log, exp, sin, cos, fabs = _fn(SQRT), _fn(LOG), _fn(EXP), _fn(SIN), _fn(COS), _fn(ABS)


def minimum(a, b):
    return Sym(MIN, (Sym.wrap(a), Sym.wrap(b)))


def maximum(a, b):
    return Sym(MAX, (Sym.wrap(a), Sym.wrap(b)))


def _shared_nodes(rows):
    """Interior nodes referenced more than once by the rows (expressions are DAGs as soon as a Python variable holding
    a sub-expression is used twice: a rotation matrix entry, a squared distance, ...)."""
    count, order = {}, []
    stack = [Sym.wrap(r) for r in rows]
    while stack:
        node = stack.pop()
        k = id(node)
        count[k] = count.get(k, 0) + 1
        if count[k] == 1:
            order.append(node)
            stack.extend(node.args)
    return {id(n): count[id(n)] for n in order if n.args and count[id(n)] > 1}


def compile_rows(rows):
    """-> (program as float64 array, number of instructions).
    A sub-expression used more than once is evaluated once: the first use is followed by TEE k (copy into temporary k),
    later uses are LOAD k.  Temporaries are per row (a row is evaluated on its own)."""
    code, offsets = [], []
    for r in rows:
        r = Sym.wrap(r)
        shared = _shared_nodes([r])   # id -> number of references
        slot, left, free = {}, {}, list(range(MAX_SLOTS - 1, -1, -1))

        def emit(node):
            k = id(node)
            if k in slot:
                code.append((LOAD, float(slot[k])))
                left[k] -= 1
                if left[k] == 0:  # last use: the temporary can hold something else from here on
                    free.append(slot.pop(k))
                return 1
            depth = 1
            if node.args:
                d = [emit(a) for a in node.args]
                depth = max(d[0], d[1] + 1) if len(d) == 2 else d[0]
            code.append((node.op, node.value if node.value is not None else 0.0))
            if k in shared:
                if not free:
                    raise ValueError("expression needs more than %d temporaries at a time" % MAX_SLOTS)
                slot[k] = free.pop()
                left[k] = shared[k] - 1
                code.append((TEE, float(slot[k])))
            return depth

        offsets.append(len(code))
        import sys
        limit = sys.getrecursionlimit()
        sys.setrecursionlimit(max(limit, 20000))
        try:
            need = emit(r)
        finally:
            sys.setrecursionlimit(limit)
        if need > MAX_STACK:
            raise ValueError("expression needs more than %d stack slots" % MAX_STACK)
        code.append((END, 0.0))
    prog = np.empty(len(rows) + 2 * len(code))
    prog[:len(rows)] = offsets
    prog[len(rows):] = np.asarray(code, dtype=float).ravel()
    return prog, len(code)


def eval_program(prog, m, x):
    """Host interpreter: f(x) of the m rows, x flat.  Same operation order as the device VM."""
    return eval_program_dual(prog, m, x, None)[0]


def eval_program_dual(prog, m, x, wrt):
    """Host interpreter with forward-mode derivatives: -> (f [m], df/dx_wrt [m]); wrt = None skips the derivative.
    The rules are those of the device VM (sco_families.cuh: vm_eval_dual)."""
    x = np.asarray(x, dtype=float).ravel()
    ins = prog[m:].reshape(-1, 2)
    out = np.empty(m)
    dout = np.zeros(m)
    for r in range(m):
        pc = int(prog[r])
        stk, slots = [], {}
        while True:
            op, arg = int(ins[pc, 0]), ins[pc, 1]
            pc += 1
            if op == END:
                break
            if op == PUSH_X:
                stk.append((x[int(arg)], 1.0 if wrt is not None and int(arg) == wrt else 0.0))
            elif op == PUSH_C:
                stk.append((arg, 0.0))
            elif op == LOAD:
                stk.append(slots[int(arg)])
            elif op == TEE:
                slots[int(arg)] = stk[-1]
            elif op in (ADD, SUB, MUL, DIV, MIN, MAX):
                b, db = stk.pop()
                a, da = stk.pop()
                if op == ADD:
                    stk.append((a + b, da + db))
                elif op == SUB:
                    stk.append((a - b, da - db))
                elif op == MUL:
                    stk.append((a * b, da * b + a * db))
                elif op == DIV:
                    q = a / b
                    stk.append((q, (da - q * db) / b))
                elif op == MIN:
                    stk.append((a, da) if a <= b else (b, db))
                else:
                    stk.append((a, da) if a >= b else (b, db))
            elif op == NEG:
                a, da = stk.pop()
                stk.append((-a, -da))
            elif op == ABS:
                a, da = stk.pop()
                stk.append((abs(a), da if a >= 0.0 else -da))
            elif op == POWI:
                a, da = stk.pop()
                k = int(arg)
                v, vm1 = 1.0, 1.0   # a^k and a^(k-1) by repeated products
                for i in range(k):
                    vm1 = v
                    v = v * a
                stk.append((v, k * vm1 * da if k > 0 else 0.0))
            else:
                a, da = stk.pop()
                if op == SQRT:
                    v = math.sqrt(a)
                    stk.append((v, da / (2.0 * v) if da != 0.0 else 0.0))  # sqrt of a constant 0 (inside a box) is flat
                elif op == LOG:
                    stk.append((math.log(a), da / a))
                elif op == EXP:
                    v = math.exp(a)
                    stk.append((v, v * da))
                elif op == SIN:
                    stk.append((math.sin(a), math.cos(a) * da))
                else:
                    stk.append((math.cos(a), -math.sin(a) * da))
        out[r], dout[r] = stk.pop()
    return out, dout


def jacobian(prog, m, n, x):
    """Exact Jacobian (m, n) by forward-mode differentiation, one pass per variable."""
    return np.stack([eval_program_dual(prog, m, x, j)[1] for j in range(n)], axis=1)
