timeout 120 python -m pytest tests -m gpu -q -x -k "moe_moments or predict" 2>&1 | tail -8
timeout 200 python tools/time_predict.py 2>&1 | tail -3
