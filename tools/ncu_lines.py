#!/usr/bin/env python
"""Aggregate an ncu report's source page per CUDA source line: tools/ncu_lines.py rep.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0].isdigit():
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)))
        except Exception: pass
ts = sum(a[3] for a in agg); ti = sum(a[4] for a in agg); tt = sum(a[5] for a in agg)
print("samples", ts, "warp-inst", ti, "thread-inst", tt, "avg lanes %.2f" % (tt / max(ti, 1)))
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    print(f"{a[0]:15s} L{a[1]:4d} inst {a[4]/ti*100:5.1f}% samp {a[3]/ts*100:5.1f}% lanes {a[5]/max(a[4],1):5.1f} | {a[2]}")
