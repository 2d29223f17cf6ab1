"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) into one row
per kernel: launches, mean / min duration, DRAM bytes per launch, DRAM GB/s of the LONGEST launch and its fraction of the HBM peak.
usage: python tools/ncu_durations.py gpurun_out/launches.csv profiles/r02_ncu_kernel_durations.csv [peak GB/s]"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
peak = float(sys.argv[3]) if len(sys.argv) > 3 else None
if peak is None:
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
hdr = rows[0]
iN, iM, iU, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
mult = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launch = defaultdict(dict)
for r in rows[1:]:
    v = float(r[iV].replace(",", "")) * mult.get(r[iU], 1.0)
    launch[r[iID]]["name"] = re.sub(r"\(.*", "", r[iN]).replace("void ", "").replace("rio::<unnamed>::", "").replace("(anonymous namespace)::", "").strip()
    launch[r[iID]][r[iM]] = v
by = defaultdict(list)
for l in launch.values():
    if "gpu__time_duration.sum" in l:
        by[l["name"]].append(l)
with open(dst, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tools/bench_kernels.py; one row per kernel; the bandwidth columns describe the LONGEST launch (the 100M-object / 134M-slot one); peak %.1f GB/s\n" % peak)
    f.write("kernel,launches,mean_ms,longest_ms,dram_read_MB,dram_write_MB,dram_GBps,frac_of_peak\n")
    for name, ls in sorted(by.items(), key=lambda kv: -max(x["gpu__time_duration.sum"] for x in kv[1])):
        big = max(ls, key=lambda x: x["gpu__time_duration.sum"])
        ms = big["gpu__time_duration.sum"]
        rd, wr = big.get("dram__bytes_read.sum", 0.0), big.get("dram__bytes_write.sum", 0.0)
        gbs = (rd + wr) / (ms * 1e-3) / 1e9 if ms > 0 else 0
        f.write("%s,%d,%.4f,%.4f,%.1f,%.1f,%.0f,%.3f\n" % (name.replace(",", ";"), len(ls), sum(x["gpu__time_duration.sum"] for x in ls) / len(ls), ms, rd / 1e6, wr / 1e6, gbs, gbs / peak))
print(open(dst).read())
