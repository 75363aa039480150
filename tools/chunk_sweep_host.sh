# chunk-count sweep of the host-buffer pipeline (MD_PIPE_CHUNKS), default bench workload, e2e figure
for c in 3 4 5 6 7 8; do
  MD_PIPE_CHUNKS=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --lean 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('chunks=$c value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
done
