# chunk-boundary sweep of the device-resident pipeline (MD_PIPE_BOUNDS), default bench workload
for b in "" "2,8" "2,6,16" "1,4,12" "4,16" "2,8,30" "3,12,30" "2,10,30" "2,6,16,30" "4,14,30"; do
  MD_PIPE_BOUNDS="$b" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --lean --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bounds=[$b]',round(d['value'],1),round(d['ms_per_step'],3))"
done
