#!/bin/bash
# quick check after a kernel change: E-step / Gram kernel tests, then the bench line with its secondary block
timeout 300 python -m pytest tests/test_cuda_kernels.py -m gpu -q -x -k "estep_kernels" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['roofline']['kernels_ms_per_step']), d['clocks']['sm_mhz'])
for k,v in d['secondary'].items(): print(k, round(v['ms_per_step'],3), json.dumps(v['kernels_ms_per_step']))"
