python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for cd in "1 24" "1 12" "1 6" "1 3" "2 6" "2 3" "4 3"; do set -- $cd
  SVO_INGEST_CTAS=$1 SVO_INGEST_DEPTH=$2 BENCH_ONLY=host timeout 100 python bench.py --steps 400 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas/image $1 depth $2 e2e', round(d['e2e']), 'p50', round(d['trace']['host_step_ms_p50'],3))"
  SVO_INGEST_CTAS=$1 SVO_INGEST_DEPTH=$2 python tools/quick_time.py C3 40 | tail -2 | head -1
done
