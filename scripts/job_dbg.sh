set -u
timeout 900 python -m pytest tests -m gpu -q -x -k "hybrid" 2>&1 | tail -4
timeout 600 python scripts/timeline_hybrid.py 2>&1 | tail -5
timeout 600 python scripts/hybrid_times.py 861 2>&1 | tail -1
