#!/bin/bash
# Evidence for BFS / SSSP / CC: ncu launch lists of the bench command and one --set full capture of each dominant kernel.
set -u
mkdir -p gpurun_out
for w in bfs sssp cc; do
  python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$w.log 2>&1 || { echo "plain $w failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bfs_|sssp_|cc_" -c 600 --csv --log-file gpurun_out/launches_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1; echo "launch list $w rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:bfs_bu_kernel -s 4 -c 1 -o gpurun_out/prof_bfs_bu_r1 \
    python bench.py --workload bfs --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bfs.log 2>&1; echo "full bfs rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sssp_relax_flat_kernel -s 130 -c 1 -o gpurun_out/prof_sssp_relax_r1 \
    python bench.py --workload sssp --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_sssp.log 2>&1; echo "full sssp rc=$?"
ncu --set full --clock-control none --import-source on -k regex:cc_hook_kernel -s 3 -c 1 -o gpurun_out/prof_cc_hook_r1 \
    python bench.py --workload cc --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_cc.log 2>&1; echo "full cc rc=$?"
