set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_gputest_d.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_gputest_d.log
timeout 300 python scripts/dac_times.py --batch 64 --frames 861 --reps 3 2>&1 | tail -3
timeout 300 python scripts/dac_times.py --batch 1 --frames 861 --reps 5 2>&1 | tail -3
timeout 600 python scripts/tc_check.py --mega --dims full --batch 1 --frames 200 --cond-len 160 --time --timeline 2>&1 | grep -v CUDAEvent | cut -c1-400 | head -20
for B in 4 64; do timeout 300 python scripts/tc_check.py --dims full --batch $B --frames 200 --cond-len 160 --time --no-compare 2>&1 | grep -v CUDAEvent | cut -c1-300; done
