#!/bin/bash
# multi-GPU visit: parity worker, then the bench lines at N GPUs (usage: scripts/gpu_multi.sh N "workloads")
set -u
N=${1:-2}
WL=${2:-"pr bfs sssp cc"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29612 tests/part_worker.py > gpurun_out/part_worker_n$N.log 2>&1; echo "worker rc=$?"; grep -c PART_WORKER_OK gpurun_out/part_worker_n$N.log
grep -A8 Traceback gpurun_out/part_worker_n$N.log | head -30
port=29620
for w in $WL; do
  port=$((port+1))
  $TR --master-port $port bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err; echo "bench $w rc=$?"
  tail -3 gpurun_out/bench_${w}_n$N.err | cut -c1-300
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_${w}_n$N.json').read().strip().splitlines()[-1])
    print('$w', 'N=$N', 'GTEPS', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3), d.get('partition','')[-90:])
except Exception as e: print('no json', e)
"
done
[ -n "${SKIP_NCCL:-}" ] && exit 0
VGLB_PR_EXCHANGE=nccl $TR --master-port 29640 bench.py --gpus $N --workload pr --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pr_nccl_n$N.json 2> gpurun_out/bench_pr_nccl_n$N.err; echo "bench pr nccl rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_pr_nccl_n$N.json').read().strip().splitlines()[-1]); print('pr nccl N=$N GTEPS', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3))"
