#!/usr/bin/env python
"""tools/summarize_ncu.py REPORT.ncu-rep [OUT.txt] — key metrics of every captured launch (read with `ncu -i … --page raw --csv`)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_sectors.sum",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
out = [f"# ncu -i {rep.split('/')[-1]} --page raw --csv   (captured with --set full --clock-control none --import-source on)"]
for r in rows[2:]:
    out.append(f"## {r[h.index('Kernel Name')][:110]}")
    for w in WANT:
        if w in h:
            out.append(f"{w} = {r[h.index(w)]} {units[h.index(w)]}")
    out.append("")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
