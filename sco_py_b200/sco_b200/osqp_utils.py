"""Scalar QP variable of the backend.  Keeps the reference's name and fields (OSQPVar,
sco_py/sco_osqp/osqp_utils.py:17-51) so that problem-building code ports unchanged: `var_name`
fixes the column order of the QP (stable sort by name, osqp_utils.py:136-143), lb / ub carry the
trust region, `val` the last QP solution."""
import numpy as np

DEFAULT_MAX_ITER = int(1e05)   # osqp_utils.py:10-15
DEFAULT_SIGMA = 5e-10
DEFAULT_RHO = 1e-01
DEFAULT_ADAPTIVE_RHO = False
DEFAULT_EPS_ABS = 1e-06
DEFAULT_EPS_REL = 1e-09


class OSQPVar(object):
    def __init__(self, var_name, lb=-np.inf, ub=np.inf, val=None):
        self.var_name = var_name
        self._lower_bound = lb
        self._upper_bound = ub
        self.val = val

    def __lt__(self, other):
        return self.var_name < other.var_name

    def __repr__(self):
        return "OSQPVar with name %s" % self.var_name

    def get_lower_bound(self):
        return self._lower_bound

    def set_lower_bound(self, lb_val):
        assert isinstance(lb_val, float) and not np.isnan(lb_val)
        self._lower_bound = lb_val

    def get_upper_bound(self):
        return self._upper_bound

    def set_upper_bound(self, ub_val):
        assert isinstance(ub_val, float) and not np.isnan(ub_val)
        self._upper_bound = ub_val
