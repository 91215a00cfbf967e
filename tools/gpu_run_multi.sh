#!/bin/bash
# usage: bash tools/gpu_run_multi.sh N   (on a box with N GPUs)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
for w in 1 2 4 8; do
  if [ $w -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 29511 tools/time_h2d.py 2>gpurun_out/h2d_$w.err | tail -1 > gpurun_out/h2d_$w.json
    cat gpurun_out/h2d_$w.json
  fi
done
if [ $N -eq 2 ]; then
  timeout 900 python -m pytest tests/test_nccl_sharded.py -m gpu -q -x -s > gpurun_out/tests_nccl.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/tests_nccl.log
  tail -15 gpurun_out/tests_nccl.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$N.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'replicas_equal',d.get('replicas_bitwise_equal'))
    print('e2e',d['e2e']['ms_per_step'],'e2e20',d['e2e_iters20']['ms_per_iteration'])
    print('cfg5',json.dumps(d.get('cfg5')))
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench_$N.err').read()[-3000:])
PY
