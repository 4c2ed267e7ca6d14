"""Per-source-line stall-sample summary of an ncu report (source page, cuda+sass correlation).
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python profiles/ncu_lines.py src.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
fname, hdr, out = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "-"):
        d = dict(zip(hdr[4:], r[4:]))
        st = sorted(((k[6:], int(v)) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0),
                    key=lambda t: -t[1])
        out.append((int(d["# Samples"]), int(d["Instructions Executed"]), fname, r[0], r[1].strip()[:90], st[:4]))
tot = sum(o[0] for o in out)
print("total samples", tot)
agg = {}
for o in out:
    agg[o[2]] = agg.get(o[2], 0) + o[0]
print("by file:", {k: "%.1f%%" % (100.0 * v / tot) for k, v in agg.items()})
out.sort(key=lambda t: -t[0])
for o in out[:top]:
    print("%6d %5.1f%% inst %9d %s:%s | %s | %s" % (o[0], 100.0 * o[0] / tot, o[1], o[2], o[3], o[4], o[5]))
