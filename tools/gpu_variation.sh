# is the flat N >= 2 loss of `value` a property of the second GPU?  the same 1-GPU bench on each GPU alone, then both at once
run1() { CUDA_VISIBLE_DEVICES=$1 python bench.py --steps 60 --warmup 6 --no-cpu-baseline --no-e2e --no-secondary --lean --no-pin 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('alone on GPU $1', round(d['value'],1), round(d['ms_per_step'],3), d['clocks'])"; }
run1 0
run1 1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 60 --warmup 6 --no-cpu-baseline --no-e2e --no-secondary --lean 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('N=2', round(d['value'],1), round(d['ms_per_step'],3), d['ms_per_step_by_rank'], d['clocks'])"
nvidia-smi --query-gpu=index,clocks.max.sm,power.limit,temperature.gpu,clocks.sm --format=csv
