set -u
timeout 300 python scripts/tc_check.py --mega --dims tiny --batch 1 2 --frames 30 --oracle 2>&1 | grep -v CUDAEvent | cut -c1-330
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
