"""Host-side constants of the device constraint families (workload generation only).

The solve path never evaluates these on the host; they exist so that synthetic
workloads can place reachable targets (workloads.gen_arm) and so the shared
parameter tables handed to the kernels (cos/sin of the DH twist angles) are the
same numbers on both sides.
"""
import numpy as np

FK7_A = np.array([0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088])
FK7_D = np.array([0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.0])
FK7_ALPHA = np.array([0.0, -np.pi / 2, np.pi / 2, np.pi / 2, -np.pi / 2, np.pi / 2, np.pi / 2])
FK7_FLANGE = 0.107


FD_BASE_STEP = np.finfo(float).eps ** (1.0 / 2.5)  # numdifftools default for first derivatives


def fk7_table():
    """30 doubles handed to the kernels: a[7], d[7], cos(alpha)[7], sin(alpha)[7], flange, FD base step."""
    return np.concatenate([FK7_A, FK7_D, np.cos(FK7_ALPHA), np.sin(FK7_ALPHA),
                           [FK7_FLANGE, FD_BASE_STEP]])


def fk7_pos(qj):
    """Flange position of the chain for joint angles qj (7,).  Plain scalar arithmetic, one rounding per
    operation, in the order the kernel uses (sco_families.cuh: fk7_pos), so host and device agree up to
    sin / cos."""
    import math
    ca_ = [float(v) for v in np.cos(FK7_ALPHA)]
    sa_ = [float(v) for v in np.sin(FK7_ALPHA)]
    R = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    p = [0.0, 0.0, 0.0]
    for i in range(7):
        ca, sa = ca_[i], sa_[i]
        th = float(qj[i])
        sn, cs = math.sin(th), math.cos(th)
        Ri = [[cs, -sn, 0.0], [sn * ca, cs * ca, -sa], [sn * sa, cs * sa, ca]]
        off = [float(FK7_A[i]), -sa * float(FK7_D[i]), ca * float(FK7_D[i])]
        p = [p[r] + ((R[r][0] * off[0] + R[r][1] * off[1]) + R[r][2] * off[2]) for r in range(3)]
        R = [[(R[r][0] * Ri[0][c] + R[r][1] * Ri[1][c]) + R[r][2] * Ri[2][c] for c in range(3)] for r in range(3)]
    return np.array([p[r] + ((R[r][0] * 0.0 + R[r][1] * 0.0) + R[r][2] * FK7_FLANGE) for r in range(3)])
