#!/bin/bash
# ESDF 512^3 check: parity tests, event timings, per-kernel time + DRAM bytes + instructions (ncu, caches left alone).
python -m pytest tests/test_gpu_esdf.py tests/test_gpu_golden_ref.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
FUELGPU_ESDF_BAND=32 python tools/esdf512.py V1 4
FUELGPU_ESDF_BAND=32 python tools/esdf512.py V0 4
python tools/esdf512.py V1 4
# second update of a run (steady state): every kernel of it, caches and clocks left alone
FUELGPU_ESDF_BAND=32 \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --cache-control none --clock-control none -k regex:"zpack|envelope|esdf" -s 33 -c 33 --csv --log-file gpurun_out/esdf512_launches.csv python tools/esdf512.py V1 2 > /dev/null 2>&1
python - <<PY
import csv
lines=[l for l in open("gpurun_out/esdf512_launches.csv") if not l.startswith("==")]
per={}
order=[]
for row in csv.DictReader(lines):
    k=(row["ID"],row["Kernel Name"][:58])
    if k not in per: per[k]={}; order.append(k)
    per[k][row["Metric Name"]]=float(row["Metric Value"].replace(",",""))
tot={}
for k in order:
    for m,v in per[k].items(): tot[m]=tot.get(m,0)+v
for k in order[:5]:
    print(k[1], {m.split("__")[1][:12]:v for m,v in per[k].items()})
print("kernels %d  sum time %.1f us  dram read %.1f MB  write %.1f MB  total %.1f MB (algorithmic 671.1 MB)  warp-instr %.1f M"%(
    len(order), tot["gpu__time_duration.sum"]/1e3, tot["dram__bytes_read.sum"]/1e6, tot["dram__bytes_write.sum"]/1e6,
    (tot["dram__bytes_read.sum"]+tot["dram__bytes_write.sum"])/1e6, tot["smsp__inst_executed.sum"]/1e6))
PY
