#!/bin/bash
# One GPU visit: parity tests, the four bench lines, the ncu launch list and one full capture of the PageRank sweep.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
for w in pr bfs sssp cc; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_pr.json 2> gpurun_out/bench_ref_pr.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pr.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pr.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pr_sweep -s 70 -c 1 -o gpurun_out/prof_pr_r1 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -n 2; cat gpurun_out/bench_pr.json
