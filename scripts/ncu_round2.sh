#!/bin/bash
# (1) --set full of the first eight tcgen05 .ts GEMM launches of a steady-state codec frame, (2) launch list of LM decode steps
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_ts -s 93 -c 8 -o gpurun_out/prof_codec_gemm_ts -f \
    python bench.py --steps 1 --warmup 3 --frames 3 --no-lm > gpurun_out/ncu_full_ts.log 2>&1
tail -1 gpurun_out/ncu_full_ts.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/lm_launches.csv \
    python scripts/lm_step_driver.py 2 32 > gpurun_out/ncu_lm.log 2>&1
tail -2 gpurun_out/ncu_lm.log
ls -la gpurun_out/*.ncu-rep
