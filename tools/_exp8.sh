mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8.json 2> gpurun_out/bench_8.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_8.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'replicas_equal',d.get('replicas_bitwise_equal'))
    print('e2e',d['e2e']['ms_per_step'],'e2e20',d['e2e_iters20']['ms_per_iteration'])
    print('cfg5',json.dumps(d.get('cfg5')))
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench_8.err').read()[-3000:])
PY
