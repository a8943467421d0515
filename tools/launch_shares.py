"""Per-kernel share of the step from an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file launches.csv python bench.py ...
  python tools/launch_shares.py launches.csv "<command>" > profiles/rNN_ncu_launch_shares.txt"""
import csv
import re
import sys
from collections import defaultdict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"^void |sed::|\(.*$", "", r[ki])
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    acc[name][0] += 1
    acc[name][1] += v
total = sum(v[1] for v in acc.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none: %s" % cmd)
print("total %.1f us over %d launches (cold-cache, serialised: compare shares)" % (total, sum(v[0] for v in acc.values())))
conv = 0.0
for name, (n, s) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    if s / total < 0.004:
        continue
    print("%-62s n=%4d sum=%9.1f us  share=%5.1f%%  avg=%8.1f us" % (name, n, s, 100 * s / total, s / n))
    if name.startswith(("conv_block1_tc_kernel", "conv_umma2_kernel")):
        conv += s
print("tensor-core conv stack (conv_block1_tc_kernel + conv_umma2_kernel) share: %.1f%%" % (100 * conv / total))
