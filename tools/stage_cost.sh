timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 8 16; do BENCH_C4_CLUSTER=$c timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 cluster $c', d['stress_c4']); print(d['single_stream'])"
done
