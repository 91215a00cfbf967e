#!/bin/bash
# session check: full GPU suite, bench, predict / expectation-input timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=10 > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -6 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
echo "bench rc=$?"
timeout 300 python tools/time_predict.py > gpurun_out/time_predict.log 2>&1; tail -8 gpurun_out/time_predict.log
timeout 300 python tools/time_given.py > gpurun_out/time_given.log 2>&1; tail -8 gpurun_out/time_given.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print('e2e',d['e2e']['ms_per_step'],'e2e20',d['e2e_iters20']['ms_per_iteration'])
print(json.dumps(d['roofline']['kernels_ms_per_step']), d['roofline']['frac'])
PY
