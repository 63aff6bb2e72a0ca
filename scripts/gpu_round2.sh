#!/bin/bash
# One gpurun call of round 2: GPU tests, smoke, the bench line, an ncu metrics pass (launch list + DRAM bytes) and a
# compute-sanitizer pass over the persistent tcgen05 GEMM.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
STAGE=${1:-all}
if [ "$STAGE" = "all" ] || [ "$STAGE" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -rA 2>&1 | tail -150 > gpurun_out/r2_pytest_gpu.log
  tail -5 gpurun_out/r2_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "bench" ]; then
  timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r2_codec_metrics.csv python bench.py --frames 3 --steps 1 --warmup 3 --no-lm > gpurun_out/r2_ncu_bench.log 2>&1
  python scripts/ncu_traffic.py gpurun_out/r2_codec_metrics.csv gpurun_out/r2_traffic.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; python bench.py --frames 3 --steps 1 --warmup 3 --no-lm" | tee gpurun_out/r2_traffic.txt
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "sanitizer" ]; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gemm_tc_gpu.py -x -q 2>&1 | grep -v "Host Frame" | head -60 > gpurun_out/r2_sanitizer_memcheck.log
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gemm_tc_gpu.py -x -q -k "ts or persistent or conv" 2>&1 | tail -25 > gpurun_out/r2_sanitizer_racecheck.log
  # LM kernels changed in round 2: cp.async-ring attention (incl. the key-split form), skinny GEMM epilogues / in-kernel finalize
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_lm_gpu.py -x -q -k "attention or skinny_gemm or norms" 2>&1 | grep -v "Host Frame" | tail -30 > gpurun_out/r2_sanitizer_memcheck_lm.log
  # (the tests' torch SDPA reference kernel, fmha_cutlassF, reports hazards of its own: count ours separately)
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_lm_gpu.py -x -q -k "rope_append_and_ring or skinny_gemm" > gpurun_out/_race_lm_full.log 2>&1
  { echo "hazard lines naming rstnet:: kernels: $(grep -c 'rstnet::' gpurun_out/_race_lm_full.log)"; echo "hazard lines naming torch's fmha_cutlassF reference kernel: $(grep -c 'fmha_cutlassF' gpurun_out/_race_lm_full.log)"; tail -n 3 gpurun_out/_race_lm_full.log; } > gpurun_out/r2_sanitizer_racecheck_lm.log; rm -f gpurun_out/_race_lm_full.log
  tail -n 4 gpurun_out/r2_sanitizer_memcheck.log; tail -n 4 gpurun_out/r2_sanitizer_racecheck.log
  tail -n 4 gpurun_out/r2_sanitizer_memcheck_lm.log; tail -n 4 gpurun_out/r2_sanitizer_racecheck_lm.log
fi
