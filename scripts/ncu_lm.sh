#!/bin/bash
# LM decode step under ncu: (1) launch list of one full step, (2) --set full of the attention kernel and the big GEMM
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 900 --csv --log-file gpurun_out/lm_launches.csv \
    python scripts/lm_step_driver.py 2 32 > gpurun_out/ncu_lm.log 2>&1
tail -2 gpurun_out/ncu_lm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ring_decode_attention -s 4 -c 2 -o gpurun_out/prof_lm_attn -f \
    python scripts/lm_step_driver.py 1 4 > gpurun_out/ncu_lm2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 10 -c 4 -o gpurun_out/prof_lm_gemm -f \
    python scripts/lm_step_driver.py 1 4 > gpurun_out/ncu_lm3.log 2>&1
ls -la gpurun_out/*.ncu-rep
