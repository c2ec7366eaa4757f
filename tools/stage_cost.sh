# developer script: throughput vs alignment cluster size
python bench.py --steps 5 --warmup 3 --streams 32 --host-threads 16 --no-cpu-baseline > /dev/null 2>&1
for c in 8 4 2 1; do
  BENCH_ONLY=device timeout 100 python bench.py --steps 200 --warmup 3 --streams 32 --host-threads 16 --align-cluster $c --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cluster $c', round(d['value']), 'frames/s', round(1e6 / d['value'], 2), 'us/frame')"
done
for c in 8 4 2 1; do SVO_ALIGN_CLUSTER=$c python tools/quick_time.py C3 40 | tail -2; done
timeout 100 python bench.py --steps 200 --warmup 3 --streams 32 --host-threads 16 --align-cluster 2 --no-cpu-baseline > gpurun_out/cl2.json 2>gpurun_out/cl2.err
