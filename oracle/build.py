"""oracle/build.py -- TEST INFRASTRUCTURE. Compiles the CPU oracle's C core.

`python oracle/build.py` (or __graft_entry__.build()) produces
oracle/libsco_oracle.so from oracle/osqp_core.c with plain gcc.  The .so is
git-ignored (built artefact) but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsco_oracle.so")
SRCS = [os.path.join(HERE, "osqp_core.c")]


def build(force=False):
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SRCS)):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-o", LIB] + SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
