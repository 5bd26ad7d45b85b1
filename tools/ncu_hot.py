"""Hottest source lines per kernel of an ncu report captured with --import-source on.
  python tools/ncu_hot.py report.ncu-rep [kernel substring] [n lines]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = kern = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        kern = r[1][:48]
        agg.setdefault(kern, [])
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0].isdigit() and kern:
        try:
            agg[kern].append((cur, int(r[0]), r[1].strip()[:100], int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)))
        except (ValueError, IndexError):
            pass
for k, a in agg.items():
    if want not in k:
        continue
    ti = sum(x[4] for x in a)
    ts = sum(x[3] for x in a)
    tt = sum(x[5] for x in a)
    print("==", k, "warp-inst", ti, "avg lanes %.1f" % (tt / max(ti, 1)))
    for x in sorted(a, key=lambda x: -x[4])[:top]:
        print("  %-15s L%4d inst %5.1f%% samp %5.1f%% lanes %5.1f | %s" %
              (x[0], x[1], x[4] / ti * 100, x[3] / max(ts, 1) * 100, x[5] / max(x[4], 1), x[2]))
