#!/bin/bash
# ESDF 512^3 check: parity tests, event timings, per-kernel time + DRAM bytes + instructions (ncu, caches left alone).
python -m pytest tests/test_gpu_esdf.py tests/test_gpu_golden_ref.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
python tools/esdf512.py V1 4
python tools/esdf512.py V0 4
FUELGPU_ESDF_BAND=32 python tools/esdf512.py V1 4
FUELGPU_ESDF_BAND=32 \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --cache-control none --clock-control none -k regex:"zpack|envelope|esdf" -c 5 --csv --log-file gpurun_out/esdf512_launches.csv python tools/esdf512.py V1 1 > /dev/null 2>&1
python - <<PY
import csv
lines=[l for l in open("gpurun_out/esdf512_launches.csv") if not l.startswith("==")]
cur=None
for row in csv.DictReader(lines):
    if row["ID"]!=cur:
        cur=row["ID"]; print()
        print(row["Kernel Name"][:60], end=" | ")
    print(row["Metric Name"].split("__")[1][:14], row["Metric Value"], end=" | ")
print()
PY
