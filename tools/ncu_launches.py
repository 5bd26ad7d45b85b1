"""Per-kernel durations of the LAST render in an ncu launch list (csv from --metrics gpu__time_duration.sum).
  python tools/ncu_launches.py gpurun_out/l_a.csv"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    u = r[mu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    seq.append((re.sub(r"\(.*", "", r[kn])[:44], ms))
idx = [i for i, (k, _) in enumerate(seq) if "raster_init" in k]
part = seq[idx[-2]:idx[-1]] if len(idx) >= 2 else seq
tot = sum(ms for _, ms in part)
for k, ms in part:
    print("  %-46s %.4f ms" % (k, ms))
print("  %-46s %.4f ms" % ("total", tot))
