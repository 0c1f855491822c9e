#!/bin/bash
# ESDF 512^3 check: parity tests, event timings, per-kernel time + DRAM bytes (ncu).
python -m pytest tests/test_gpu_esdf.py -m gpu -x -q 2>&1 | tail -3
python tools/esdf512.py V1 4
python tools/esdf512.py V0 4
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"zsweep|envelope|esdf" -c 6 --csv --log-file gpurun_out/esdf512_launches.csv python tools/esdf512.py V1 1 > /dev/null 2>&1
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
