# which part of an initialised NCCL communicator costs the device-resident loop its 4 % at N = 2?  (bench value, ms per step)
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 6 --no-cpu-baseline --no-e2e --lean 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$*',round(d['value'],1),round(d['ms_per_step'],3))"; }
run X=1
run NCCL_NVLS_ENABLE=0
run NCCL_P2P_DISABLE=1
run NCCL_CUMEM_ENABLE=0
run NCCL_P2P_DISABLE=1 NCCL_SHM_DISABLE=1
