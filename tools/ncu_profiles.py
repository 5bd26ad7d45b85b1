#!/usr/bin/env python
"""Turn one round of ncu output into the tracked files under profiles/:
  tools/ncu_profiles.py <launch-list.csv> <full-capture.ncu-rep> <tag>
writes profiles/<tag>_launches_bench.csv (copy), profiles/<tag>_launches_bench_summary.txt (share of device
time per kernel), profiles/<tag>_render_pipeline_bench.txt (per-kernel metrics + hottest source lines per
kernel) and profiles/render_traffic.json (dram bytes per render, read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1:4]
prof = os.path.join(ROOT, "profiles")

# ---- launch list: share of device time per kernel
rows = list(csv.reader(open(launches, errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    us = v / 1e3 if r[mu] in ("ns", "nsecond") else v * 1e3 if r[mu] in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("fb::", "")
    tot[name] += us
    cnt[name] += 1
shutil.copy(launches, os.path.join(prof, tag + "_launches_bench.csv"))
total = sum(tot.values())
with open(os.path.join(prof, tag + "_launches_bench_summary.txt"), "w") as f:
    f.write("share of device time per kernel over `python bench.py --steps 2 --warmup 3` "
            "(ncu launch list, cold-cache serialised; includes the end-to-end leg's chunked renders)\n")
    f.write("total %.1f us\n" % total)
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write("%6.2f%%  n=%4d  %12.1f us  %s\n" % (100 * v / total, cnt[k], v, k))

# ---- full capture: per-kernel metrics, traffic, hot lines
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
summ = summ.split("\n== hottest source lines")[0]
hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "", "14"], capture_output=True, text=True).stdout
with open(os.path.join(prof, tag + "_render_pipeline_bench.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on, one launch of every kernel of one render of the bench\n"
            "workload (200 000 triangles, 92 views, 512 x 512), driver profiles/prof_render.py\n\n")
    f.write(summ)
    f.write("\n== hottest source lines per kernel (share of warp instructions, share of stall samples, mean active lanes)\n")
    f.write(hot)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hd, un = rr[0], rr[1]
ik, ird, iwr = hd.index("Kernel Name"), hd.index("dram__bytes_read.sum"), hd.index("dram__bytes_write.sum")


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


per = {}
for r in rr[2:]:
    per[re.sub(r"\(.*", "", r[ik]).replace("void ", "")] = int(to_bytes(r[ird], un[ird]) + to_bytes(r[iwr], un[iwr]))
json.dump({"dram_bytes_per_render": sum(per.values()), "per_kernel": per,
           "source": "profiles/%s_render_pipeline_bench.txt (ncu --set full, one launch of each pipeline kernel, "
                     "bench workload 92 views x 512^2)" % tag},
          open(os.path.join(prof, "render_traffic.json"), "w"), indent=1)
print(open(os.path.join(prof, tag + "_launches_bench_summary.txt")).read()[:1500])
print(json.dumps(per, indent=1), sum(per.values()))
