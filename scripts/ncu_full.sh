#!/bin/bash
# --set full captures of the dominant kernels (one GPU; bring the .ncu-rep back in gpurun_out/)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 6 -o gpurun_out/prof_codec_gemm_tc -f \
    python bench.py --steps 1 --warmup 3 --frames 3 --no-lm > gpurun_out/ncu_full1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ring_decode_attention -s 1 -c 2 -o gpurun_out/prof_lm_attn -f \
    python scripts/lm_step_driver.py 1 4 > gpurun_out/ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rvq_level -s 8 -c 2 -o gpurun_out/prof_rvq -f \
    python bench.py --steps 1 --warmup 3 --frames 3 --no-lm > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep
