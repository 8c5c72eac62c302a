#!/bin/bash
set -u
timeout 300 ./dev/pr_lab 24 > gpurun_out/plain_lab.log 2>&1 || exit 1
for s in 2 8 26; do
ncu --set full --clock-control none -k regex:staged -s $s -c 1 -o gpurun_out/prof_lab_$s -f ./dev/pr_lab 24 > gpurun_out/ncu_lab_$s.log 2>&1
done
ls -la gpurun_out
