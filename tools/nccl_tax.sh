# which part of a 2-rank run costs rank 1 its ~4 % of the device-resident loop?  (bench value, ms per step of each rank)
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 6 --no-cpu-baseline --no-e2e --no-secondary --lean $EXTRA 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$* $EXTRA',round(d['value'],1),[round(x,3) for x in d['ms_per_step_by_rank']], d.get('host_pinning'))"; }
run X=1
EXTRA=--no-pin run X=1
run CUDA_VISIBLE_DEVICES=1,0
run NCCL_P2P_DISABLE=1 NCCL_SHM_DISABLE=1
run MD_BENCH_NO_SAMPLER=1
run CUDA_DEVICE_MAX_CONNECTIONS=32
