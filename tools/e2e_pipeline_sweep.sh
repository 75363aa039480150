# e2e through streams.BatchPipeline (two contexts, MD_MEM_HOST_ASYNC) against the single chained context, for several chunk counts of the host pipeline
for c in ${CHUNKS:-7 5 4 3}; do MD_PIPE_CHUNKS=$c python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-secondary --lean ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('CHUNKS=$c value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),d['e2e'].get('api','')[:24],'| one context',round(d['e2e_one_context']['value'],1) if d.get('e2e_one_context') else None)"; done
