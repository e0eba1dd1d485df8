"""DRAM bytes per launch of a step-kernel capture -> profiles/r2_traffic.json (read by bench.py for roofline.traffic).

    ncu -i X.ncu-rep --page raw --csv > raw.csv
    python profiles/tools/ncu_traffic.py raw.csv B4096_T20 "gpurun_out/X.ncu-rep (...)"  [json path]
"""
import csv
import json
import os
import sys

raw, key, source = sys.argv[1], sys.argv[2], sys.argv[3]
path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(__file__), "..", "r2_traffic.json")
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]


def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v * scale


rd, wr, t_us = get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum")
try:
    data = json.load(open(path))
except Exception:
    data = {}
data[key] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
             "kernel": vals[hdr.index("Kernel Name")], "ncu_time_us": t_us, "source": source}
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
print(key, data[key])
