"""Summarise an .ncu-rep: headline metrics + top stall sites.  python tools/ncu_summary.py file.ncu-rep"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed.sum.per_cycle_elapsed", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for i, name in enumerate(h):
    if name in want or any(name == w + s for w in want for s in ("", ".per_second")):
        print(f"{name:75s} {v[i]:>16s} {u[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
si, srci, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[2:] if len(r) > si and r[si].strip().isdigit()]
tot = sum(int(r[si]) for r in data)
agg = collections.Counter()
for r in data:
    for i in stall_cols:
        if r[i].strip().isdigit():
            agg[h[i]] += int(r[i])
print("stalls:", {k: v for k, v in agg.most_common(7)}, "instrs", len(data))
for r in sorted(data, key=lambda r: -int(r[si]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    st = {h[i]: int(r[i]) for i in stall_cols if r[i].strip().isdigit() and int(r[i]) > 0}
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}% ie={r[ie]:>8} {r[srci][:64]:64s} {top}")
