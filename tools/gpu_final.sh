#!/bin/bash
# end-of-round record: full GPU suite, the driver's bench command (both arms), dev timings of the widened rows
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=10 > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"
timeout 300 python tools/time_predict.py > gpurun_out/time_predict.log 2>&1; tail -3 gpurun_out/time_predict.log
timeout 300 python tools/time_given.py > gpurun_out/time_given.log 2>&1; tail -2 gpurun_out/time_given.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print('e2e',d['e2e']['ms_per_step'],'e2e20',d['e2e_iters20']['ms_per_iteration'])
print(json.dumps(d['roofline']['kernels_ms_per_step']), d['roofline']['frac'])
for k,v in d['secondary'].items(): print(k, v['ms_per_step'], json.dumps(v['kernels_ms_per_step']))
PY
head -c 500 gpurun_out/bench_ref.json
