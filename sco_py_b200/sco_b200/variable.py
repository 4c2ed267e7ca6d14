"""Variable: an ordered array of scalar QP variables with a current and a saved value.  Same
constructor, methods, copies and errors as sco_py/sco_osqp/variable.py:4-73."""
import numpy as np


class Variable(object):
    def __init__(self, osqp_vars, value=None):
        assert isinstance(osqp_vars, np.ndarray) and len(osqp_vars) > 0
        self._osqp_vars = osqp_vars.copy()
        if value is not None:
            assert isinstance(value, np.ndarray) and osqp_vars.shape == value.shape
            self._value = value.copy()
        else:
            self._value = None
        self._saved_value = None

    def get_osqp_vars(self):
        return self._osqp_vars

    def get_value(self):
        return None if self._value is None else self._value.copy()

    def add_trust_region(self, trust_box_size):
        """Box of half-width `trust_box_size` around the SAVED value, written into the scalar
        variables' bounds (variable.py:37-45; overwrites user bounds, quirk C-11)."""
        assert self._saved_value is not None
        for index, ov in np.ndenumerate(self._osqp_vars):
            ov.set_lower_bound(float(self._saved_value[index] - trust_box_size))
            ov.set_upper_bound(float(self._saved_value[index] + trust_box_size))

    def update(self):
        value = np.zeros(self._osqp_vars.shape)
        for index, ov in np.ndenumerate(self._osqp_vars):
            if ov.val is None:
                raise ValueError("The variable %s does not have a legitimate value" % ov.var_name)
            value[index] = ov.val
        self._value = value

    def save(self):
        assert not np.any(np.isnan(self._value))
        self._saved_value = self._value.copy()

    def restore(self):
        self._value = self._saved_value.copy()
