"""Variable of the B200 backend.

API contract taken from the reference's OSQP backend (sco_py/sco_osqp/variable.py:4-73): a Variable
groups scalar QP variables (`OSQPVar`) in a fixed array shape, carries their current numeric value
and one saved snapshot, and turns a trust-region radius into bounds on the scalars.  Arrays cross the
boundary by copy in both directions, as there (variable.py:19,23,33).

The solve itself never touches these objects: `batch.compile_batch` reads the initial values out of
them once and `batch.scatter_solution` writes the device result back (into `_value` and into every
scalar's `.val`).
"""
import numpy as np


def _as_object_array(scalars):
    if not isinstance(scalars, np.ndarray):
        raise AssertionError("osqp_vars must be a numpy array of OSQPVar objects")
    if scalars.size == 0:
        raise AssertionError("a Variable needs at least one scalar variable")
    return np.array(scalars, dtype=object, copy=True)


class Variable(object):
    __slots__ = ("_osqp_vars", "_value", "_saved_value")

    def __init__(self, osqp_vars, value=None):
        self._osqp_vars = _as_object_array(osqp_vars)
        self._saved_value = None
        self._value = None
        if value is not None:
            if not isinstance(value, np.ndarray) or value.shape != self._osqp_vars.shape:
                raise AssertionError("value must be a numpy array shaped like osqp_vars")
            self._value = np.array(value, dtype=float, copy=True)

    # -- accessors -----------------------------------------------------------------
    def get_osqp_vars(self):
        return self._osqp_vars

    def get_value(self):
        return None if self._value is None else np.array(self._value, copy=True)

    # -- snapshot ------------------------------------------------------------------
    def save(self):
        """Remember the current value (the centre of the next trust region)."""
        if np.isnan(self._value).any():
            raise AssertionError("cannot save a value that contains NaN")
        self._saved_value = np.array(self._value, copy=True)

    def restore(self):
        self._value = np.array(self._saved_value, copy=True)

    # -- QP side -------------------------------------------------------------------
    def add_trust_region(self, trust_box_size):
        """Bounds saved_value -/+ radius on every scalar (this replaces whatever bounds the scalars
        had: quirk C-11 of SURVEY.md, variable.py:43-45)."""
        if self._saved_value is None:
            raise AssertionError("save() must be called before add_trust_region()")
        centre = self._saved_value.ravel()
        for k, scalar in enumerate(self._osqp_vars.ravel()):
            scalar.set_lower_bound(float(centre[k] - trust_box_size))
            scalar.set_upper_bound(float(centre[k] + trust_box_size))

    def update(self):
        """Pull the last QP solution out of the scalars' `.val`; a scalar that was never solved for is
        an error (ValueError, as variable.py:57-59)."""
        flat = self._osqp_vars.ravel()
        missing = [s.var_name for s in flat if s.val is None]
        if missing:
            raise ValueError("The variable %s does not have a legitimate value" % missing[0])
        self._value = np.array([s.val for s in flat], dtype=float).reshape(self._osqp_vars.shape)
