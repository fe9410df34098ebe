"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [first] [count]"""
import collections, csv, sys
rows, hdr = [], None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        rows.append(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[first:first + count]
agg = collections.OrderedDict()
for r in rows:
    name = r[ki].split("(")[0][-60:]
    ms = float(r[vi].replace(",", "")) / 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.3f} ms summed kernel time (cold-cache, serialised under ncu)")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms:8.3f} ms  {100 * ms / tot:5.1f}%  x{n:4d}  {name}")
