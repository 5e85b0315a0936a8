set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_gputest_c.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_gputest_c.log
for B in 1 4 64; do
  timeout 300 python scripts/tc_check.py --dims full --batch $B --frames 440 --cond-len 160 --time --timeline 2>&1 | grep -v CUDAEvent > $O/r2_tl_b$B.log; tail -40 $O/r2_tl_b$B.log | head -3
done
