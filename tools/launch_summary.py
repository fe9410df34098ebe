"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel.
    python tools/launch_summary.py launches.csv [first_launch_index [top_n [end_index]]]
    (default: second half = the timed step)"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
data = rows[1:]
start = int(sys.argv[2]) if len(sys.argv) > 2 else len(data) // 2
end = int(sys.argv[4]) if len(sys.argv) > 4 else len(data)
sel = data[start:end]
agg, cnt = collections.Counter(), collections.Counter()
for r in sel:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v
    name = r[ki].split("(")[0][-70:]
    agg[name] += v
    cnt[name] += 1
tot = sum(agg.values())
print(f"{len(sel)} launches, {tot:.3f} ms summed kernel time (cold-cache, serialised under ncu)")
for k, v in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
    print(f"{v:8.3f} ms {100 * v / tot:5.1f}%  x{cnt[k]:3d}  {k}")
