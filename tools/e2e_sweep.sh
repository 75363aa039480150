# e2e (host buffers) against points per LK CTA and host-pipeline chunk count
for n in 8 4 1; do for c in 0 4 8; do
  MD_LK_NPTS=$n MD_PIPE_CHUNKS=$c python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-secondary --lean 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('NPTS=$n chunks=$c value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
done; done
