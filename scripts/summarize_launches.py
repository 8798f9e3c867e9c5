"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count / time / share of one step.
usage: python scripts/summarize_launches.py launches.csv [marker-kernel-substring]"""
import collections, csv, sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "traj_viou_warp"
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)]
idx = [i for i, r in enumerate(rows) if marker in r[0]]
seg = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows
seg = [r for r in seg if "spin_kernel" not in r[0]]      # the head-start spin kernel of bench.py's instrumented step is not part of the step
agg = collections.OrderedDict()
for k, v in seg:
    k = k.split("(")[0][:72]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
print("one step between the last two '%s' launches: %d launches, %.3f ms of kernel time" % (marker, len(seg), tot / 1e6))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-74s n=%5d %10.3f ms %5.1f%%" % (k, n, v / 1e6, 100 * v / tot))
