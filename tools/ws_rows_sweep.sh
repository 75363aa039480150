# fused planes + sums kernel: rows per CTA segment (MD_WS_ROWS) and threads per CTA (MD_WS_NT, 0 = per-level choice) on the default bench
for cfg in ${CFGS:-"160 0" "320 0" "240 0" "160 128" "320 256"}; do set -- $cfg; MD_WS_ROWS=$1 MD_WS_NT=$2 python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-secondary --lean --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('ROWS=$1 NT=$2',round(d['value'],1),round(d['ms_per_step'],3),[(s['kernel'][:2],round(s['ms'],3)) for s in d['stages']])"; done
