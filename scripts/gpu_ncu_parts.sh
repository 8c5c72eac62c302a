#!/bin/bash
set -u
for m in 2 1; do
  export VGLB_PR_ONLY=$m
  python scripts/dev_pr_one.py > gpurun_out/plain_only$m.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pr_sweep -s 2 -c 1 -o gpurun_out/prof_pr_only$m -f python scripts/dev_pr_one.py > gpurun_out/ncu_only$m.log 2>&1
  tail -n 1 gpurun_out/ncu_only$m.log
done
