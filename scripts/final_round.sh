#!/bin/bash
# Round-end validation on one GPU: parity tests (twice: the second pass would expose a flaky race), smoke, both bench arms.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest 1" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== pytest 2" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -2 | tee -a gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench.log; cut -c1-250 gpurun_out/bench.log
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.log; cut -c1-400 gpurun_out/bench_ref.log
