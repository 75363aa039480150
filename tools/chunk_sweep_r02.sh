for b in "" "16" "11,22" "8,16,24" "4,18" "6,19"; do
  MD_PIPE_BOUNDS="$b" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --lean --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('device bounds=[$b]',round(d['value'],1),round(d['ms_per_step'],3))"
done
for c in 5 6 7 8; do
  MD_PIPE_CHUNKS=$c python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --lean 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('host chunks=$c e2e',round(d['e2e']['value'],1))"
done
