timeout 180 python -m pytest tests/test_cuda_kernels.py -m gpu -q -x -k "gram_kernels" 2>&1 | tail -4
timeout 300 python -m pytest tests -m gpu -q -x -k "molt or arhmm or diag or iso or hmm or given" 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d['roofline']['kernels_ms_per_step']))
for k,v in d['secondary'].items(): print(k, v['ms_per_step'], json.dumps(v['kernels_ms_per_step']))"
