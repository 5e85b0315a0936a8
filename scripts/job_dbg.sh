set -u
timeout 900 python -m pytest tests -m gpu -q -x -k "hybrid" 2>&1 | tail -12
echo "== mega hybrid"; timeout 600 python scripts/hybrid_times.py 861 2>&1 | tail -2
echo "== graph path"; ZB_MEGA_HYBRID=0 timeout 600 python scripts/hybrid_times.py 861 2>&1 | tail -2
