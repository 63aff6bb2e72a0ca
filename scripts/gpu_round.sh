#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
if [ "$1" == "ncu" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --frames 3 --no-lm > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
fi
