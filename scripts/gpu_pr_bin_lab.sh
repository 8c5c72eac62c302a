#!/bin/bash
# developer run of dev/pr_bin_lab (column-binned gather lab): a few configurations + optionally one ncu capture of the bin kernel
set -u
out=gpurun_out/r2_pr_bin_lab.txt
: > $out
for cost in 0 0.004 0.008; do
for cfg in "32 49152 32" "32 49152 16"; do
  LAB_RUN_COST=$cost timeout 300 ./dev/pr_bin_lab 24 $cfg 2>&1 | grep -v "^graph\|^flat" >> $out
done
done
cat $out
if [ "${1:-}" = "ncu" ]; then
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:bin_kernel -s 1 -c 1 -o gpurun_out/r2_pr_bin_lab -f ./dev/pr_bin_lab 24 32 49152 32 > gpurun_out/ncu_pr_bin_lab.log 2>&1
  tail -3 gpurun_out/ncu_pr_bin_lab.log
fi
