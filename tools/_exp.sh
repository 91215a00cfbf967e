timeout 400 python bench.py --steps 10 --warmup 3 2>gpurun_out/b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['roofline']['kernels_ms_per_step']))
for k,v in d['secondary'].items(): print(k, json.dumps({a:b for a,b in v.items() if a in ('ms_per_step','hbm_frac','roofline_frac','kernels_ms_per_step','error','algorithmic_gb_per_step')}))"
tail -3 gpurun_out/b.err
