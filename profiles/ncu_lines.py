#!/usr/bin/env python
"""Summarise an ncu report by CUDA source line: stall samples and executed instructions.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv; python ncu_lines.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > 7 and r[0].isdigit():
        try:
            out.append((int(r[6] or 0), int(r[7] or 0), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print(f"total samples {tot}, warp instructions {toti}")
for s, i, f, l, src in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{100*s/tot:5.1f}% samp {100*i/toti:5.1f}% inst  {f}:{l}: {src}")
