"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list per kernel."""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
per = OrderedDict()
for r in rd:
    name = r["Kernel Name"].split("(")[0]
    metric, val = r["Metric Name"], r["Metric Value"].replace(",", "")
    try:
        v = float(val)
    except ValueError:
        continue
    unit = r["Metric Unit"]
    k = per.setdefault(name, {"launch_ids": set(), "time_us": 0.0, "rd": 0.0, "wr": 0.0})
    k["launch_ids"].add(r["ID"])
    if metric == "gpu__time_duration.sum":
        k["time_us"] += v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("u") else v * 1e3)
    elif metric == "dram__bytes_read.sum":
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        k["rd"] += v * mult
    elif metric == "dram__bytes_write.sum":
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        k["wr"] += v * mult
tot = sum(k["time_us"] for k in per.values())
print(f"{'kernel':60s} {'launches':>8s} {'time_ms':>9s} {'share':>6s} {'avg_us':>8s} {'dramRd_MB/l':>11s} {'dramWr_MB/l':>11s}")
for name, k in sorted(per.items(), key=lambda kv: -kv[1]["time_us"]):
    n = len(k["launch_ids"])
    print(f"{name[:60]:60s} {n:8d} {k['time_us'] / 1e3:9.3f} {100 * k['time_us'] / tot:5.1f}% {k['time_us'] / n:8.1f} "
          f"{k['rd'] / n / 1e6:11.2f} {k['wr'] / n / 1e6:11.2f}")
print(f"total {tot / 1e3:.3f} ms over {sum(len(k['launch_ids']) for k in per.values())} launches")
