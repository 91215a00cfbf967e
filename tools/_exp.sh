mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "wsum or given or golden or rowterm" > gpurun_out/tests_wsum.log 2>&1; tail -15 gpurun_out/tests_wsum.log
timeout 300 python tools/time_given.py 2>&1 | tail -3
