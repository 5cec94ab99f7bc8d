"""Summarise an ncu CSV (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) per kernel name:
launches, device time, DRAM bytes moved, achieved DRAM GB/s and its fraction of the measured HBM copy bandwidth
(MEASURED_PEAKS.json hbm_gbs).  Per-launch ncu durations are cold-cache and serialised: the GB/s column is what one launch
achieves in isolation; the in-step time shares come from tools/kernel_times.py (CUPTI).
usage: python tools/summarize_mem_kernels.py gpurun_out/x.csv > profiles/x.md"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
header = None
for r in rd:
    if header is None:
        if "Kernel Name" in r and "Metric Name" in r:
            header = {n: i for i, n in enumerate(r)}
        continue
    if len(r) < len(header):
        continue
    rows.append(r)
per = defaultdict(lambda: defaultdict(float))
launch = {}
for r in rows:
    kid = r[header["ID"]]
    name = re.sub(r"\(.*", "", r[header["Kernel Name"]]).replace("b200::", "").replace("void ", "")
    metric, unit, val = r[header["Metric Name"]], r[header["Metric Unit"]], float(r[header["Metric Value"]].replace(",", ""))
    if metric == "gpu__time_duration.sum":
        val *= {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "nsecond": 1e-9, "ms": 1e-3, "msecond": 1e-3, "second": 1.0, "s": 1.0}[unit]
    else:
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1.0}[unit]
    launch.setdefault(kid, name)
    per[kid][metric] += val
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for kid, m in per.items():
    a = agg[launch[kid]]
    a[0] += 1
    a[1] += m["gpu__time_duration.sum"]
    a[2] += m["dram__bytes_read.sum"]
    a[3] += m["dram__bytes_write.sum"]
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | time (ms, ncu cold-cache) | DRAM read (MB) | DRAM write (MB) | achieved GB/s | of measured HBM peak %.0f GB/s |" % peak)
print("|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / a[1] / 1e9 if a[1] > 0 else 0.0
    print("| `%s` | %d | %.3f | %.1f | %.1f | %.0f | %.2f |" % (name[:70], a[0], a[1] * 1e3, a[2] / 1e6, a[3] / 1e6, gbs, gbs / peak))
print("\ntotal time of the listed launches: %.2f ms" % (tot * 1e3))
