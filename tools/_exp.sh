timeout 600 python -m pytest tests -m gpu -q -x -k "streamed or full_size" 2>&1 | tail -2
timeout 600 python tools/time_stream.py 2>&1 | tail -8 | cut -c1-330
