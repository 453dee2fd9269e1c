"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv`).

    python profiles/summarize_launches.py gpurun_out/launches.csv "header comment" [skip_first_n_launches] > profiles/x.csv

Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes.
GB/s = (dram read + write) / duration of the same launch."""
import collections
import csv
import sys

path = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0

per_id = collections.OrderedDict()
with open(path) as f:
    rows = [r for r in f if r.startswith('"')]
for r in csv.DictReader(rows):
    d = per_id.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))

agg = collections.OrderedDict()
for i, d in per_id.items():
    if i < skip:
        continue
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    a[2] += d.get("dram__bytes_read.sum", 0.0) / 1e6
    a[3] += d.get("dram__bytes_write.sum", 0.0) / 1e6
    a[4] += d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
total = sum(a[1] for a in agg.values())
print(f"# {note}")
print("# per-launch times are cold-cache and serialised: compare SHARES; GB/s = (dram read+write)/duration")
print(f"# launches counted from ID {skip}; total {total:.1f} us")
print("launches,us_per_launch,share_pct,dram_read_MB,dram_write_MB,dram_GBs,ncu_dram_pct_of_peak,kernel")
for name, (n, us, rd, wr, pct) in agg.items():
    short = name.replace("admmtv::", "").replace("void ", "")
    gbs = (rd + wr) / us * 1e3 if us > 0 else 0.0
    print(f"{n},{us / n:.1f},{100 * us / total:.1f},{rd / n:.1f},{wr / n:.1f},{gbs:.0f},{pct / n:.1f},\"{short[:90]}\"")
