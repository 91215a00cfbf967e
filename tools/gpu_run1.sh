#!/bin/bash
# GPU box script: full GPU suite, bench line, ncu launch list (run under gpurun from the repo root)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -s --durations=15 > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"estep|gram|niw|hmm|mnw" -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/tests_gpu.log
cat gpurun_out/bench.json | head -c 6000
