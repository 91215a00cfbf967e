timeout 200 python -m pytest tests -m gpu -q -x -k "softmax_rows or given or prx" 2>&1 | tail -6
timeout 200 python tools/time_given.py 2>&1 | tail -1
