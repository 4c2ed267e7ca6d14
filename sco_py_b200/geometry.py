"""Device expression library for motion-planning constraints.

The reference's callers (OpenTAMP, README.md:4) hand `Expr` black-box Python callables for collision avoidance and
kinematics; a kernel cannot call Python, so here the same functions are written once in the expression language of
`sco_py_b200.sym` and become rows of a `SymExpr` -- evaluated on the device by the VM family, differentiated there in
forward mode (`SymExpr(rows, n, analytic=True)`) or by finite differences like a gradient-less `Expr`.

Every function takes and returns `sym.Sym` expressions (or plain numbers for constants); points are sequences of
2 or 3 expressions.  Sub-expressions that are used more than once (a squared distance, the entries of a rotation
matrix along a kinematic chain) are evaluated once: `sym.compile_rows` turns shared nodes into temporaries.

    signed distances      box_sdf, capsule_sdf, sphere_sdf           (negative inside)
    collision pairs       sphere_pair_clearance, point_capsule_clearance
    kinematics            dh_chain (modified DH, Craig): pose of every link; flange position / orientation
    orientation           rotation_error(R, R_des): 0.5 sum_k R[:, k] x R_des[:, k]  (zero iff aligned, small-angle = axis*angle)

Typical use (an obstacle-avoidance LEq row per way-point and obstacle, an end-effector pose Eq block):
    x = sym.variables(n)
    rows = [margin - box_sdf((x[2*t], x[2*t+1]), centre, half) for t in range(T)]
    prob.add_cnt_expr(BoundExpr(LEqExpr(SymExpr(rows, n, analytic=True), np.zeros((T, 1))), var))
"""
import math

from . import sym


def _sub(p, q):
    return [a - b for a, b in zip(p, q)]


def dot(p, q):
    out = None
    for a, b in zip(p, q):
        t = a * b
        out = t if out is None else out + t
    return out


def norm(p):
    return sym.sqrt(dot(p, p))


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def sphere_sdf(p, centre, radius):
    """Signed distance of point p to a ball / disc."""
    return norm(_sub(p, centre)) - radius


def box_sdf(p, centre, half):
    """Signed distance of point p to an axis-aligned box (2-D or 3-D): centre, half extents."""
    q = [abs(a - c) - h for a, c, h in zip(p, centre, half)]
    outside = norm([sym.maximum(v, 0.0) for v in q])
    inner = q[0]
    for v in q[1:]:
        inner = sym.maximum(inner, v)
    return outside + sym.minimum(inner, 0.0)


def capsule_sdf(p, a, b, radius):
    """Signed distance of point p to the capsule with axis a-b and the given radius."""
    pa, ba = _sub(p, a), _sub(b, a)
    h = sym.minimum(sym.maximum(dot(pa, ba) / dot(ba, ba), 0.0), 1.0)
    return norm([u - v * h for u, v in zip(pa, ba)]) - radius


def sphere_pair_clearance(p1, r1, p2, r2):
    """Clearance of two balls (collision pair): > 0 apart, < 0 penetrating."""
    return norm(_sub(p1, p2)) - (r1 + r2)


def point_capsule_clearance(p, r, a, b, radius):
    """Clearance of a ball around p and a capsule (a link of another body)."""
    return capsule_sdf(p, a, b, radius) - r


def dh_chain(q, a, d, alpha):
    """Modified-DH chain (Craig): T_i = Rx(alpha_i) Tx(a_i) Rz(q_i) Tz(d_i).  q: joint expressions; a, d, alpha: numbers.
    -> list of (R, p) per link: R 3x3 rows of expressions, p the origin of frame i in the base frame.  The operation
    order is that of the closed FK7 family (sco_families.cuh: fk7_pos)."""
    R = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    p = [0.0, 0.0, 0.0]
    frames = []
    for qi, ai, di, al in zip(q, a, d, alpha):
        ca, sa = math.cos(al), math.sin(al)
        st, ct = sym.sin(qi), sym.cos(qi)
        Ri = [[ct, -st, 0.0], [st * ca, ct * ca, -sa], [st * sa, ct * sa, ca]]
        off = [ai, -sa * di, ca * di]
        p = [p[r] + ((R[r][0] * off[0] + R[r][1] * off[1]) + R[r][2] * off[2]) for r in range(3)]
        R = [[(R[r][0] * Ri[0][c] + R[r][1] * Ri[1][c]) + R[r][2] * Ri[2][c] for c in range(3)] for r in range(3)]
        frames.append((R, p))
    return frames


def tool_position(frame, offset):
    """Position of a point fixed in a link frame (e.g. the flange: offset (0, 0, 0.107))."""
    R, p = frame
    return [p[r] + ((R[r][0] * offset[0] + R[r][1] * offset[1]) + R[r][2] * offset[2]) for r in range(3)]


def rotation_error(R, R_des):
    """0.5 * sum_k R[:, k] x R_des[:, k]: the usual orientation error vector (zero iff the frames coincide; for small
    errors the rotation axis times the angle).  R: rows of expressions, R_des: rows of numbers."""
    e = [0.0, 0.0, 0.0]
    for k in range(3):
        c = cross([R[0][k], R[1][k], R[2][k]], [R_des[0][k], R_des[1][k], R_des[2][k]])
        e = [e[i] + c[i] for i in range(3)]
    return [0.5 * v for v in e]
