set -u
O=gpurun_out
timeout 600 python scripts/tc_check.py --mega --dims full --batch 1 --frames 200 --cond-len 160 --time --timeline 2>&1 | grep -v CUDAEvent | cut -c1-400 | head -20
timeout 300 python scripts/tc_check.py --mega --dims tiny --batch 1 2 --frames 12 --oracle 2>&1 | grep -v CUDAEvent | cut -c1-400
