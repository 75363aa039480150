for n in 4 2 8; do MD_LK_NPTS=$n python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-secondary --lean 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('NPTS=$n',round(d['value'],1),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1),[(s['kernel'][:2],round(s['ms'],3)) for s in d['stages']])"; done
