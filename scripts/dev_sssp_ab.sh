#!/bin/bash
# developer A/B of the SSSP bench line (uniform s24 ef32): the library as built + dev/lib_<name>.so builds (+ knobs given as NAME=VALUE)
run() { python bench.py --workload sssp --steps 16 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), 'GTEPS', round(d['ms_per_step'],3), 'ms', d['per_source_gteps'], 'iters', d.get('extras'))"; }
run default
for n in "$@"; do VGLB_LIB_PATH=$PWD/dev/lib_$n.so run $n; done
