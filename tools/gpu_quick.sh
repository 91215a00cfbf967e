#!/bin/bash
# quick check after a kernel change: kernel + parity tests, then the bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x ${QUICK_K:+-k "$QUICK_K"} > gpurun_out/tests_quick.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_quick.log
tail -5 gpurun_out/tests_quick.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print('e2e',d['e2e']['ms_per_step'],'e2e20',d['e2e_iters20']['ms_per_iteration'])
print(json.dumps(d['roofline']['kernels_ms_per_step']), d['roofline']['frac'])
for k,v in d['secondary'].items(): print(k, v['ms_per_step'], json.dumps(v['kernels_ms_per_step']))
PY
