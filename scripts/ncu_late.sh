#!/bin/bash
# launch list of a LATE frame of the first pass (attention context ~200 of 250): skip ~100 frames of 177 launches
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${1:-18500} -c 400 --csv --log-file gpurun_out/launches_late.csv \
    python bench.py --steps 1 --warmup 3 --no-lm > gpurun_out/ncu_bench_late.log 2>&1
tail -1 gpurun_out/ncu_bench_late.log | cut -c1-120
