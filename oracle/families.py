"""oracle/families.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatements of the closed nonlinear-constraint families the CUDA engine
evaluates on device (sco_py_b200/csrc/sco_families.cuh).  In the reference these
are user-supplied black-box callables handed to `Expr(f, grad)`
(sco_py/expr.py:22-25); SURVEY.md section 8(d) fixes the concrete functions of
the benchmark configurations.  Every function takes x of shape (n, 1) and
returns f of shape (m, 1) / J of shape (m, n), the contract of expr.py:34-41
and expr.py:78-100.
"""
import numpy as np

# ---------------------------------------------------------------- QUADFORM (C4)


def tri_index(n):
    """Row-major packed upper-triangle index pairs (r <= c)."""
    r, c = np.triu_indices(n)
    return r, c


def unpack_sym(packed, n):
    """packed (..., n(n+1)/2) -> symmetric (..., n, n)."""
    r, c = tri_index(n)
    out = np.zeros(packed.shape[:-1] + (n, n))
    out[..., r, c] = packed
    out[..., c, r] = packed
    return out


def pack_sym(S):
    n = S.shape[-1]
    r, c = tri_index(n)
    return S[..., r, c]


def quadform_f(x, Pm, a):
    """f_j = 0.5 x'P_j x + a_j'x ;  Pm (m, n, n) symmetric, a (m, n)."""
    xv = x[:, 0]
    Px = Pm @ xv  # (m, n)
    return (0.5 * Px @ xv + a @ xv)[:, None]


def quadform_grad(x, Pm, a):
    return Pm @ x[:, 0] + a


# ---------------------------------------------------------------- CIRCLE2D (C2)


def circle_f(x, centers, radii, T):
    """Row t*K+k: R_k - ||p_t - c_k||, p_t = x[2t:2t+2]."""
    p = x[:, 0].reshape(T, 2)
    d = p[:, None, :] - centers[None, :, :]  # (T, K, 2)
    dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
    return (radii[None, :] - dist).reshape(-1, 1)


def circle_grad(x, centers, radii, T):
    K = centers.shape[0]
    p = x[:, 0].reshape(T, 2)
    d = p[:, None, :] - centers[None, :, :]
    dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
    J = np.zeros((T * K, 2 * T))
    for t in range(T):
        for k in range(K):
            J[t * K + k, 2 * t] = -d[t, k, 0] / dist[t, k]
            J[t * K + k, 2 * t + 1] = -d[t, k, 1] / dist[t, k]
    return J


# ---------------------------------------------------------------- FK7 (C3)
# Modified-DH (Craig) chain, Franka-like constants; position of the flange.
FK7_A = np.array([0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088])
FK7_D = np.array([0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.0])
FK7_ALPHA = np.array([0.0, -np.pi / 2, np.pi / 2, np.pi / 2, -np.pi / 2, np.pi / 2, np.pi / 2])
FK7_FLANGE = 0.107


_FK7_CA = [float(v) for v in np.cos(FK7_ALPHA)]
_FK7_SA = [float(v) for v in np.sin(FK7_ALPHA)]


def fk7_pos(qj):
    """qj (7,) -> (3,) flange position.  T_i = Rx(alpha_i) Tx(a_i) Rz(q_i) Tz(d_i).

    Scalar arithmetic in a FIXED order -- every product rounded on its own, three-term sums left to right --
    so that the CUDA kernel (sco_families.cuh: fk7_pos with __dmul_rn / __dadd_rn) performs the very same
    operations and the two sides differ only through sin / cos.  (A BLAS-backed `R @ v` sums in an order that
    depends on the library build; profiles/arm_sensitivity.py shows that differences of ~10 ulp in f are enough
    to make the SQP of an occasional arm problem stop one iteration earlier or later.)"""
    import math
    R = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    p = [0.0, 0.0, 0.0]
    for i in range(7):
        ca, sa = _FK7_CA[i], _FK7_SA[i]
        qi = float(qj[i])
        st, ct = math.sin(qi), math.cos(qi)
        Ri = [[ct, -st, 0.0], [st * ca, ct * ca, -sa], [st * sa, ct * sa, ca]]
        pi = [float(FK7_A[i]), -sa * float(FK7_D[i]), ca * float(FK7_D[i])]
        p = [p[r] + ((R[r][0] * pi[0] + R[r][1] * pi[1]) + R[r][2] * pi[2]) for r in range(3)]
        R = [[(R[r][0] * Ri[0][c] + R[r][1] * Ri[1][c]) + R[r][2] * Ri[2][c] for c in range(3)] for r in range(3)]
    return np.array([p[r] + ((R[r][0] * 0.0 + R[r][1] * 0.0) + R[r][2] * FK7_FLANGE) for r in range(3)])


def fk7_f(x):
    """Rows 0..2: flange position for the joints of the LAST time-step (x[-7:])."""
    return fk7_pos(x[-7:, 0]).reshape(3, 1)


# ---------------------------------------------------------------- VM (stack programs, C1 toy problems)
# Independent restatement of the interpreter of sco_py_b200/sym.py (the product's host code): program =
# m row offsets, then (opcode, operand) pairs; opcodes END 0, PUSH_X 1, PUSH_C 2, ADD 3, SUB 4, MUL 5,
# DIV 6, NEG 7, POWI 8, SQRT 9, LOG 10, EXP 11, SIN 12, COS 13, ABS 14, MIN 15, MAX 16, TEE 17, LOAD 18 (temporaries).


def _vm_run(xv, prog, m, wrt):
    """Values and d/dx_wrt (wrt = -1: values only) of the m rows; dual numbers as (value, derivative) tuples."""
    import math
    ins = np.asarray(prog[m:], dtype=float).reshape(-1, 2)
    val, der = np.zeros(m), np.zeros(m)
    for r in range(m):
        pc, stack, tmp = int(prog[r]), [], {}
        while int(ins[pc, 0]) != 0:
            op, arg = int(ins[pc, 0]), float(ins[pc, 1])
            pc += 1
            if op == 1:
                stack.append((float(xv[int(arg)]), 1.0 if int(arg) == wrt else 0.0))
            elif op == 2:
                stack.append((arg, 0.0))
            elif op in (3, 4, 5, 6, 15, 16):
                (b, db), (a, da) = stack.pop(), stack.pop()
                if op == 3:
                    stack.append((a + b, da + db))
                elif op == 4:
                    stack.append((a - b, da - db))
                elif op == 5:
                    stack.append((a * b, da * b + a * db))
                elif op == 6:
                    q = a / b
                    stack.append((q, (da - q * db) / b))
                elif op == 15:
                    stack.append((a, da) if a <= b else (b, db))
                else:
                    stack.append((a, da) if a >= b else (b, db))
            elif op == 7:
                a, da = stack.pop()
                stack.append((-a, -da))
            elif op == 8:
                a, da = stack.pop()
                v, vm1 = 1.0, 1.0
                for _ in range(int(arg)):
                    vm1 = v
                    v *= a
                stack.append((v, int(arg) * vm1 * da if int(arg) > 0 else 0.0))
            elif op == 14:
                a, da = stack.pop()
                stack.append((abs(a), da if a >= 0.0 else -da))
            elif op == 17:
                tmp[int(arg)] = stack[-1]
            elif op == 18:
                stack.append(tmp[int(arg)])
            else:
                a, da = stack.pop()
                if op == 9:
                    v = math.sqrt(a)
                    stack.append((v, da / (2.0 * v) if da != 0.0 else 0.0))
                elif op == 10:
                    stack.append((math.log(a), da / a))
                elif op == 11:
                    v = math.exp(a)
                    stack.append((v, v * da))
                elif op == 12:
                    stack.append((math.sin(a), math.cos(a) * da))
                else:
                    stack.append((math.cos(a), -math.sin(a) * da))
        val[r], der[r] = stack.pop()
    return val, der


def vm_f(x, prog, m):
    return _vm_run(np.asarray(x, dtype=float).ravel(), prog, m, -1)[0].reshape(m, 1)


def vm_grad(x, prog, m):
    """Exact Jacobian (m, n) of a program: forward mode, one pass per variable -- what a user-supplied `grad` of an
    `Expr(f, grad)` returns (expr.py:86-88)."""
    xv = np.asarray(x, dtype=float).ravel()
    return np.stack([_vm_run(xv, prog, m, j)[1] for j in range(xv.size)], axis=1)
