#!/bin/bash
mkdir -p gpurun_out
timeout 200 tools/bin/probe_fetch > gpurun_out/probe_fetch.txt 2>&1; cat gpurun_out/probe_fetch.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'^(g_|s_)' --csv --log-file gpurun_out/probe_fetch_ncu.csv tools/bin/probe_fetch > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/probe_fetch_ncu.csv')) if len(r)>8]
h=rows[0]; d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[h.index('ID')], r[h.index('Kernel Name')].split('(')[0]),{})[r[h.index('Metric Name')]]=float(r[h.index('Metric Value')].replace(',',''))
seen={}
for (i,k),m in d.items():
    seen[k]=seen.get(k,0)+1
    if seen[k]!=2 and not k.startswith('s_'): continue
    print(f"{k:20s} id={i:3s} read {m['dram__bytes_read.sum']/2**26:7.1f} B/q  write {m['dram__bytes_write.sum']/2**26:6.1f} B/q  {m['gpu__time_duration.sum']/1e6:7.3f} ms")
PY
