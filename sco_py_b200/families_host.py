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
    R = np.eye(3)
    p = np.zeros(3)
    ca_, sa_ = np.cos(FK7_ALPHA), np.sin(FK7_ALPHA)
    for i in range(7):
        ca, sa = ca_[i], sa_[i]
        ct, st = np.cos(qj[i]), np.sin(qj[i])
        Ri = np.array([[ct, -st, 0.0], [st * ca, ct * ca, -sa], [st * sa, ct * sa, ca]])
        p = p + R @ np.array([FK7_A[i], -sa * FK7_D[i], ca * FK7_D[i]])
        R = R @ Ri
    return p + R @ np.array([0.0, 0.0, FK7_FLANGE])
