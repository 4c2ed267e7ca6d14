"""sco_py_b200 -- B200-native batched penalty-SQP behind the sco_py API surface."""
__version__ = "0.1.0"
