timeout 900 python -m pytest tests -m gpu -q -x -k "stream" 2>&1 | tail -8
