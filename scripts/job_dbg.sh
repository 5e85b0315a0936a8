set -u
timeout 900 python -m pytest tests -m gpu -q -x -k "full_size or batch3 or tcgen05_decode_step or hybrid or gemm" 2>&1 | tail -3
for E in "ZB_TC_TWO_CTA=0" "ZB_TC_TWO_CTA=1"; do
echo "== $E"; env $E python bench.py --batch 64 --frames 200 --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['roofline']['us_per_launch'],1), d['breakdown_ms'])"
done
