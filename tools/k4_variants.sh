set -e
cd motion_detection_b200/csrc
for v in 2 4 3; do
  touch k_mask.cu
  make -s NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -DMASK_CTAS_PER_SM=$v" >/dev/null 2>&1
  grep -A2 "k_maskILb1" k_mask.ptxas.log | tail -2 | tr '\n' ' '; echo
  (cd ../..; python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('CTAS=$v',d['value'],[(s['kernel'][:3],round(s['ms'],3)) for s in d['stages']])")
done
