run() { python bench.py --workload bfs --steps 16 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],4), d['per_source_gteps'])"; }
run default
for a in 4 8 30 60; do VGLB_BFS_ALPHA=$a run alpha$a; done
for b in 4 9 36 72; do VGLB_BFS_BETA=$b run beta$b; done
python bench.py --workload bfs20 --steps 16 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bfs20', round(d['value'],1), round(d['ms_per_step'],4), d['per_source_gteps'])"
