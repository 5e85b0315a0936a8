set -u
timeout 900 python -m pytest tests -m gpu -q -x -k "hybrid" 2>&1 | tail -3
python bench.py --variant hybrid --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hyb b1', round(d['value'],2), round(d['roofline']['us_per_launch'],1), d['breakdown_ms'])"
python bench.py --variant hybrid --batch 32 --prefix-frames 258 --frames 300 --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5', round(d['roofline']['us_per_launch'],1), d['breakdown_ms'])"
