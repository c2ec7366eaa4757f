# developer script: marginal GPU cost of every stage at saturation (duplicate the idempotent stage, read the throughput)
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for d in none:0 pyr:1 align:1 klt:1 refine:1 ssd:1 none:0; do
  SVO_DIAG_DUP=$d BENCH_ONLY=device timeout 100 python bench.py --steps 400 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$d', round(d['value']), 'frames/s', round(1e6 / d['value'], 2), 'us/frame')"
done
