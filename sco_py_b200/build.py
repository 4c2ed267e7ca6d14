"""Builds sco_py_b200/libsco_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libsco_b200.so")
CSRC = os.path.join(PKG, "csrc")
SRCS = [os.path.join(CSRC, "sco_abi.cu")]
DEPS = SRCS + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [
    os.path.join(ROOT, "include", "sco_b200.h")]


def build(force=False, verbose=False):
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in DEPS)):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += SRCS + ["-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
