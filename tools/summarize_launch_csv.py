#!/usr/bin/env python
"""tools/summarize_launch_csv.py LAUNCHES.csv — per-launch table from `ncu --csv --log-file` (long form: one row per
launch and metric): kernel, duration, DRAM bytes, L2 hit rate; then totals per kernel name."""
import csv
import sys
from collections import OrderedDict, defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi, ui, ii = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = OrderedDict()
for r in rows:
    if r is hdr or len(r) <= vi or not r[ii].isdigit():
        continue
    d = launches.setdefault(r[ii], {"kernel": r[ki].split("(")[0][:60]})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if u == "ns":
        v /= 1e6
    elif u == "us":
        v /= 1e3
    elif u == "s":
        v *= 1e3
    elif u == "Kbyte":
        v *= 1e3
    elif u == "Mbyte":
        v *= 1e6
    elif u == "Gbyte":
        v *= 1e9
    d[r[mi]] = v
tot = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
print(f"{'id':>4} {'ms':>9} {'dram_rd_MB':>11} {'dram_wr_MB':>11} {'l2hit%':>7}  kernel")
for i, d in launches.items():
    ms = d.get("gpu__time_duration.sum", 0.0)
    rd, wr = d.get("dram__bytes_read.sum", 0.0) / 1e6, d.get("dram__bytes_write.sum", 0.0) / 1e6
    print(f"{i:>4} {ms:9.3f} {rd:11.1f} {wr:11.1f} {d.get('lts__t_sector_hit_rate.pct', float('nan')):7.1f}  {d['kernel']}")
    t = tot[d["kernel"]]
    t[0] += 1
    t[1] += ms
    t[2] += rd
    t[3] += wr
print("\n# per kernel: launches, total ms, mean ms, mean DRAM read MB, mean DRAM write MB")
for k, t in tot.items():
    print(f"{t[0]:4d} {t[1]:9.3f} {t[1] / t[0]:9.3f} {t[2] / t[0]:11.1f} {t[3] / t[0]:11.1f}  {k}")
