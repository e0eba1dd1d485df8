"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv`
launch list: last launch of every kernel (earlier ones are warm-up), units normalised.
usage: python profiles/tools/ncu_launch_summary.py gpurun_out/r2_all_kernels.csv"""
import collections
import csv
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0,
        "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per = collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    val = float(d["Metric Value"].replace(",", "")) * UNIT.get(d["Metric Unit"], 1.0)
    per.setdefault(name, collections.OrderedDict()).setdefault(d["ID"], {})[d["Metric Name"]] = val
print("| kernel | launches | time of last launch (us) | DRAM read (MB) | DRAM written (MB) | DRAM GB/s | registers | grid |")
print("|---|---|---|---|---|---|---|---|")
for name, launches in per.items():
    last = list(launches.values())[-1]
    t = last.get("gpu__time_duration.sum", float("nan"))
    rd, wr = last.get("dram__bytes_read.sum", 0.0), last.get("dram__bytes_write.sum", 0.0)
    print(f"| `{name}` | {len(launches)} | {t:.1f} | {rd / 1e6:.3f} | {wr / 1e6:.3f} | {(rd + wr) / t / 1e3:.1f} | "
          f"{last.get('launch__registers_per_thread', 0):.0f} | {last.get('launch__grid_size', 0):.0f} |")
