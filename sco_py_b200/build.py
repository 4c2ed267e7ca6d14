"""Builds sco_py_b200/libsco_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

The kernels are instantiated once per team size (threads per problem) from
csrc/sco_team.cu; the translation units are compiled in parallel and linked with
the host side (csrc/sco_abi.cu) into one shared library.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
# SCO_BUILD_TAG=<tag> builds a side-by-side variant (libsco_b200_<tag>.so, objects in build_<tag>/), e.g. the
# cycle-counting diagnostic build:  SCO_BUILD_TAG=timing SCO_NVCC_FLAGS=-DSCO_TIMING python -m sco_py_b200.build
# and load it with SCO_B200_LIB=.../libsco_b200_timing.so
TAG = os.environ.get("SCO_BUILD_TAG", "")
LIB = os.path.join(PKG, "libsco_b200%s.so" % ("_" + TAG if TAG else ""))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build" + ("_" + TAG if TAG else ""))
TEAMS = (32, 64, 128, 256, 512)
DENSE_KINDS = (1, 2, 3, 4)  # size table of the dense ADMM loop, see sco_create
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "sco_b200.h")]


def _units():
    """(object, source, extra flags)"""
    units = [(os.path.join(OBJ, "sco_abi.o"), os.path.join(CSRC, "sco_abi.cu"), []),
             (os.path.join(OBJ, "sco_probe.o"), os.path.join(CSRC, "sco_probe.cu"), [])]
    for t in TEAMS:
        units.append((os.path.join(OBJ, "sco_team_%d.o" % t), os.path.join(CSRC, "sco_team.cu"),
                      ["-DSCO_TEAM=%d" % t]))
    for dk in DENSE_KINDS:
        units.append((os.path.join(OBJ, "sco_dense_%d.o" % dk), os.path.join(CSRC, "sco_team.cu"),
                      ["-DSCO_TEAM=64", "-DSCO_DK=%d" % dk]))
    return [u for u in units if os.path.exists(u[1])]


def build(force=False, verbose=False):
    deps = _deps()
    # extra flags (e.g. -DSCO_TIMING, whose kernels overwrite result slots with cycle counts) are part of what a
    # build is: objects compiled under other flags are never reused
    flags = os.environ.get("SCO_NVCC_FLAGS", "").strip()
    stamp = os.path.join(OBJ, "flags.stamp")
    try:
        same_flags = open(stamp).read() == flags
    except OSError:
        same_flags = not os.path.exists(OBJ) or not os.listdir(OBJ)
    if not same_flags:
        force = True
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in deps)):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    with open(stamp, "w") as f:
        f.write(flags)
    base = [nvcc] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                            "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        base += ["-Xptxas", "-v"]
    base += flags.split()  # e.g. -DSCO_TIMING for the clock64 experiment

    def compile_one(u):
        obj, src, extra = u
        p = subprocess.run(base + extra + ["-c", src, "-o", obj], capture_output=True, text=True)
        return u, p

    units = _units()
    headers = [d for d in deps if not d.endswith(".cu")]

    def stale(u):
        obj, src, _ = u
        return (force or not os.path.exists(obj)
                or any(os.path.getmtime(obj) < os.path.getmtime(d) for d in headers + [src]))

    todo = [u for u in units if stale(u)]
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 4))) as ex:
        results = list(ex.map(compile_one, todo))
    for (obj, src, extra), p in results:
        if verbose or p.returncode:
            sys.stderr.write("== %s %s\n%s%s" % (os.path.basename(src), " ".join(extra), p.stdout, p.stderr))
        if p.returncode:
            raise RuntimeError("nvcc failed on %s %s" % (src, extra))
    subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", LIB] + [u[0] for u in units])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
