"""Extracts the SASS of the plain ADMM iteration loop of dense_qp_solve<20,30> (the headline kernel's inner loop) with
instruction counts.  No GPU needed:  python profiles/sass_loop.py > profiles/r2_sass_dense_loop.txt"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
obj = os.path.join(tmp, "d3.o")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "sco_py_b200", "csrc"), "-DSCO_TEAM=64", "-DSCO_DK=3",
                "-c", os.path.join(ROOT, "sco_py_b200", "csrc", "sco_team.cu"), "-o", obj], check=True, capture_output=True)
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-sf", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith("$_Z7k_solve") and "dense_qp_solveILi20ELi30" in l and l.rstrip().endswith(":")][0]
end = [i for i in range(start + 1, len(dis)) if ".type" in dis[i] and "@function" in dis[i]][0]
ins = []
for l in dis[start:end]:
    m = re.search(r"/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    elif re.match(r"\s*\.L_x_\d+:", l):
        ins.append((None, l.strip()))
labels = {t[:-1]: i for i, (a, t) in enumerate(ins) if a is None}


def mnemonic(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]


cands = []
for k, (a, t) in enumerate(ins):
    m = re.search(r"BRA.*`\((\.L_x_\d+)\)", t)
    if a is not None and m and m.group(1) in labels and labels[m.group(1)] < k:
        body = [x for x in ins[labels[m.group(1)]:k + 1] if x[0] is not None]
        cnt = Counter(mnemonic(x[1]) for x in body)
        cands.append((len(body), labels[m.group(1)], k, cnt))
# the plain-iteration loop: smallest backward-branch body that holds the 50-term dot product and exactly one barrier
loop = min(c for c in cands if 45 <= c[3].get("DFMA", 0) <= 70 and c[3].get("BAR.SYNC.DEFER_BLOCKING", 0) == 1)
n, lo, hi, cnt = loop
pref = lambda p: sum(v for k_, v in cnt.items() if k_.startswith(p))
print("SASS of the plain ADMM iteration loop of dense_qp_solve<20,30> inside k_solve<64,3> (C4: n = 20, m = 30)")
print("nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo; cuobjdump -xelf + nvdisasm; profiles/sass_loop.py")
print("loop body: %d instructions per trip = one ADMM iteration; both warps run it (warp 0 = penalty rows, warp 1 = variables)," % n)
print("the row and the variable update sequences are both in the body, a warp executes its own")
print()
print("instruction mix: " + ", ".join("%s %d" % kv for kv in sorted(cnt.items(), key=lambda kv: -kv[1])))
print("FP64-pipe instructions %d (DFMA %d, DADD %d, DMUL %d, DSETP %d) | shared loads %d (LDS.128 %d) | shared stores %d | barriers %d | "
      "LOCAL loads %d, LOCAL stores %d" % (pref("DFMA") + pref("DADD") + pref("DMUL") + pref("DSETP") + pref("DMNMX"), pref("DFMA"),
                                           pref("DADD"), pref("DMUL"), pref("DSETP"), pref("LDS"), pref("LDS.128"), pref("STS"),
                                           pref("BAR"), pref("LDL"), pref("STL")))
print("tensor-core / TMA mnemonics (UTCMMA, UTMALDG, UBLKCP ...): %d -- an FP64 mat-vec has no tcgen05 path (DESIGN.md section 4)"
      % sum(v for k_, v in cnt.items() if k_.startswith(("UTC", "UTMA", "UBLK"))))
print()
for a, t in ins[lo:hi + 1]:
    print(("        /*%04x*/  %s" % (a, t)) if a is not None else t)
