#!/bin/bash
# Round-2 evidence run (one B200, under gpurun): bench lines, launch lists, ncu --set full of the ESDF kernels.
# Numbers printed under ncu are never bench values; the bench JSONs come from separate runs.
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench.err
# every launch of two bench steps with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02_launches_office.csv \
    python bench.py --steps 4 --warmup 3 --no-esdf512 > /dev/null 2>&1
# one whole 512^3 ESDF update in steady state: every kernel, DRAM bytes, instructions (caches left alone)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --cache-control none --clock-control none -k regex:"zpack|envelope" -s 33 -c 33 --csv --log-file gpurun_out/r02_esdf512_launches.csv \
    python tools/esdf512.py V1 2 > /dev/null 2>&1
# ncu --set full of the three ESDF kernels (zpack, zy tile, x tile) of the second update
ncu --set full --import-source on --cache-control none --clock-control none -k regex:"zpack|envelope" -s 33 -c 3 \
    -o gpurun_out/r02_esdf512_full python tools/esdf512.py V1 2 > /dev/null 2>&1
# 512^3 frontier search: every kernel
# (one search = 6 launches of the sweep + small-path attempt, 1 recompaction, 20 of union-find/claims/scans, then
#  8 per split level enqueued in batches of 12 levels = 119 launches after 1 of the upload: the window is the third search)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none \
    -s 239 -c 119 --csv --log-file gpurun_out/r02_frontier512_launches.csv python tools/frontier512.py > /dev/null 2>&1
# the long-line case (BASELINE config 4 geometry, one GPU): 2-CTA cluster tiles
python tools/esdf_long.py 1024 1024 256 4 > gpurun_out/r02_esdf_long.log 2>&1
FUELGPU_ESDF_CLUSTER=0 python tools/esdf_long.py 1024 1024 256 4 >> gpurun_out/r02_esdf_long.log 2>&1
ls -la gpurun_out
