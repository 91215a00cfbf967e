timeout 200 python -m pytest tests -m gpu -q -x -k "hmm or arhmm" 2>&1 | tail -2
timeout 200 python tools/time_models.py cfg4 2>&1 | tail -3
