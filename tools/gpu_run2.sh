#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s --durations=8 > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
timeout 300 python tools/time_sustained.py > gpurun_out/sustained.log 2>&1
timeout 600 python bench.py --cfg5 on --no-secondary --no-cpu-baseline --steps 10 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
echo "bench rc=$?" >> gpurun_out/bench_cfg5.err
timeout 300 python tools/time_models.py > gpurun_out/time_models.log 2>&1
tail -15 gpurun_out/tests_gpu.log; cat gpurun_out/sustained.log; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_cfg5.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step']); print('cfg5',json.dumps(d.get('cfg5'))); print('e2e',d['e2e']['ms_per_step'],'e2e20',d['e2e_iters20']['ms_per_iteration'])
except Exception as e: print('bench parse failed',e)
PY
tail -5 gpurun_out/bench_cfg5.err; cat gpurun_out/time_models.log
