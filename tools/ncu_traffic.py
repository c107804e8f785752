"""profiles/<round>_traffic.json from an ncu launch list with DRAM byte counters: DRAM read + write bytes per launch of
the conv kernels (conv1d_tc_kernel + trunk_kernel), averaged over their launches -- what bench.py copies into
``roofline.traffic``.
    python tools/ncu_traffic.py gpurun_out/launches.csv profiles/r2_traffic.json"""
import csv
import json
import sys
from collections import OrderedDict

path, out = sys.argv[1], sys.argv[2]
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
per = OrderedDict()
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in csv.DictReader(lines):
    name = r["Kernel Name"].split("(")[0]
    if "conv1d_tc_kernel" not in name and "trunk_kernel" not in name:
        continue
    key = "conv1d_tc_kernel" if "conv1d_tc" in name else name.replace("void trunk::", "")
    k = per.setdefault(key, {"ids": set(), "read": 0.0, "write": 0.0})
    k["ids"].add(r["ID"])
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if r["Metric Name"] == "dram__bytes_read.sum":
        k["read"] += v * MULT.get(r["Metric Unit"], 1)
    elif r["Metric Name"] == "dram__bytes_write.sum":
        k["write"] += v * MULT.get(r["Metric Unit"], 1)
n = sum(len(k["ids"]) for k in per.values())
tot = sum(k["read"] + k["write"] for k in per.values())
doc = {"kernel": "ou::tc::conv1d_tc_kernel + ou::trunk::trunk_kernel",
       "source": f"{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                 "--clock-control none over `OU_PIPELINE=0 bench.py --steps 1 --warmup 1 --diffusion-steps 4` "
                 "(full-batch launches, the same op stream bench.py's roofline pass times); dram read+write averaged per launch",
       "per_kernel_mb_per_launch": {key: {"read": round(k["read"] / len(k["ids"]) / 1e6, 2),
                                          "write": round(k["write"] / len(k["ids"]) / 1e6, 2),
                                          "launches": len(k["ids"])} for key, k in per.items()},
       "launches": n, "traffic_bytes_per_launch": int(tot / max(n, 1))}
with open(out, "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps(doc)[:400])
