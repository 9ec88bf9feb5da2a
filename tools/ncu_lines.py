"""Aggregates `ncu --page source --print-source cuda,sass --csv` per CUDA source line: samples and instructions."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = None
data = []
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > ii and r[0] != "":
        try: data.append((int(r[si]), int(r[ii]), int(r[0]), r[1]))
        except ValueError: pass
tot = sum(d[0] for d in data) or 1; toti = sum(d[1] for d in data) or 1
print("total samples", tot, "warp instructions", toti)
top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for d in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"{d[0]:6d} {d[0]/tot:6.1%} inst={d[1]:8d} {d[1]/toti:6.1%} L{d[2]:<4d} {d[3].strip()[:105]}")
