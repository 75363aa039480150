# k_window_sums variants on the default bench: MD_WS_MODE 2 = fused planes + sums (default), 1 = planes + single-pass sums, 0 = planes + sliding sums
for m in ${MODES:-2 1 0}; do MD_WS_MODE=$m python bench.py --steps 30 --warmup 4 --no-cpu-baseline --no-secondary --lean ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('WS_MODE=$m',round(d['value'],1),round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1) if d.get('e2e') else None,[(s['kernel'][:2],round(s['ms'],3)) for s in d['stages']])"; done
