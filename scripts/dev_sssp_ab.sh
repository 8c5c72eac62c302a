#!/bin/bash
# developer A/B of the SSSP bench line (uniform s24 ef32): delta scale knob and dev/lib_<name>.so builds
run() { python bench.py --workload sssp --steps 16 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), 'GTEPS', round(d['ms_per_step'],3), 'ms', d['per_source_gteps'])"; }
run default
for s in 2 3 6 8; do VGLB_SSSP_DELTA_SCALE=$s run delta$s; done
for n in "$@"; do VGLB_LIB_PATH=$PWD/dev/lib_$n.so run $n; done
