python bench.py --steps 5 --warmup 3 --streams 32 --host-threads 16 --no-cpu-baseline > /dev/null 2>&1
for st in "8 8" "16 16" "32 16"; do set -- $st
  BENCH_ONLY=device timeout 100 python bench.py --steps 200 --warmup 3 --streams $1 --host-threads $2 --align-cluster 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('S=$1 T=$2 value', round(d['value']), 'p50 step ms', round(d['trace']['host_step_ms_p50'],3))"
done
SVO_INGEST_MIX=1 timeout 100 python bench.py --steps 200 --warmup 3 --streams 32 --host-threads 16 --align-cluster 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mix value', round(d['value']), 'e2e', round(d['e2e']['value']))"
SVO_NO_INGEST=1 timeout 100 python bench.py --steps 200 --warmup 3 --streams 32 --host-threads 16 --align-cluster 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dma value', round(d['value']), 'e2e', round(d['e2e']['value']))"
