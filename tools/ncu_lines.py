"""Aggregate the warp-stall samples of an .ncu-rep by CUDA source line.
usage: python tools/ncu_lines.py <rep> [top]   (ncu -i ... --page source --print-source cuda,sass --csv)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = collections.Counter()
reasons = collections.defaultdict(collections.Counter)
src = {}
fname, hdr = "", None
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and len(r) > i_s:
        try:
            n = int(r[i_s])
        except ValueError:
            continue
        if r[0]:
            key = (fname, int(r[0]))
            src[key] = r[1].strip()[:90]
        if n:
            agg[key] += n
            for i in stall:
                if r[i] not in ("", "0"):
                    reasons[key][hdr[i][6:]] += int(r[i])
tot = sum(agg.values())
print("samples", tot)
for key, n in agg.most_common(top):
    print(f"{n:6d} {100 * n / tot:5.1f}%  {key[0]}:{key[1]:<5d} {src.get(key, ''):90s} {reasons[key].most_common(3)}")
