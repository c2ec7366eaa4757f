timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for e in X=1 SVO_NO_FORK=1; do
  env $e BENCH_NO_C4=1 timeout 120 python bench.py --steps 400 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e value', round(d['value']), 'e2e', round(d['e2e']['value']), d['single_stream'])"
done
