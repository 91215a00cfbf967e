#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q -s --durations=10 > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -6 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err
echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gram_umma_kernel" --launch-skip 2 -c 1 -f -o /tmp/ncu/gram python tools/prof_driver.py cfg2 3 > gpurun_out/ncu_gram.log 2>&1
ncu -i /tmp/ncu/gram.ncu-rep --page details > gpurun_out/r02_ncu_gram_details.txt 2>&1
ncu -i /tmp/ncu/gram.ncu-rep --page raw --csv > gpurun_out/r02_ncu_gram_raw.csv 2>&1
ncu -i /tmp/ncu/gram.ncu-rep --page source --csv > gpurun_out/r02_ncu_gram_source.csv 2>&1
grep -E "Duration|SM Frequency|DRAM Throughput" gpurun_out/r02_ncu_gram_details.txt | head
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}))
print('e2e',d['e2e']['ms_per_step'],'e2e20',d['e2e_iters20']['ms_per_iteration'])
print(json.dumps(d['roofline']['kernels_ms_per_step']), d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'])
print(json.dumps(d['secondary'])[:1500])
print(json.dumps(d['cpu_baseline']))
PY
cat gpurun_out/bench_ref.json | head -c 600
