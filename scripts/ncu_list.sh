#!/bin/bash
# ncu launch list of a short bench pass (3 frames per step): per-launch device time, cold-cache + serialised.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${1:-1200} -c ${2:-400} --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --frames 3 --no-lm > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
