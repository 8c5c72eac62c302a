#!/bin/bash
set -u
export VGLB_PR_VARIANT=7
for m in 124 1 2; do
  export VGLB_PR_TIERMASK=$m
  python scripts/dev_pr_one.py > gpurun_out/plain_m$m.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pr_sweep -s 2 -c 1 -o gpurun_out/prof_pr_mask$m -f python scripts/dev_pr_one.py > gpurun_out/ncu_m$m.log 2>&1
  tail -n 2 gpurun_out/ncu_m$m.log
done
