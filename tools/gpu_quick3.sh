#!/bin/bash
timeout 120 python -m pytest tests/test_cuda_kernels.py -m gpu -q -x -k "gram" 2>&1 | tail -4
timeout 120 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "gmm_golden or cfg2" 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['roofline']['kernels_ms_per_step']), d['clocks']['sm_mhz'])"
