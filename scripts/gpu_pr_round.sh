#!/bin/bash
# developer round for the PageRank sweep: parity tests, bench with and without the column bins, launch list of one run
set -u
python -m pytest tests -m gpu -x -q -k "pagerank or pr" 2>&1 | tail -3
python bench.py --no-extras --no-cpu-baseline 2> gpurun_out/r2e_bench.err > gpurun_out/r2e_bench.json
python scripts/show_bench.py gpurun_out/r2e_bench.json 2>/dev/null | head -3
VGLB_PR_NO_BINS=1 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no bins', d['value'], d['ms_per_step'], d['e2e'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pr_" -c 600 --csv --log-file gpurun_out/r2_pr_launches.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_pr_launches.csv")) if len(r)>10]
h=rows[0]
out=[]
for r in rows[1:]:
    d=dict(zip(h,r)); out.append((d["Kernel Name"].split("(")[0], float(d["Metric Value"].replace(",",""))))
for k,t in out[-8:]: print(k, t/1000, "us")
P
