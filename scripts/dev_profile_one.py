"""Developer aid for ncu: build one BASELINE graph and run ONE algorithm call on it (no parity gate, no timing loop), e.g.
    ncu --set full -k regex:sssp_relax_flat -c 40 -o gpurun_out/x python scripts/dev_profile_one.py sssp
Workloads: pr | bfs | sssp | cc (bench.py's configurations)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vectorgraphlibrary_b200 as vgl  # noqa: E402
from vectorgraphlibrary_b200 import dist as vdist  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "pr"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kind, scale, ef, _ = bench.WORKLOADS[workload]
if len(sys.argv) > 3:
    scale = int(sys.argv[3])
ctx = vgl.Context(0)
r = vdist.SingleGpuRunner(vgl, ctx, bench.algo_of(workload), kind, scale, ef, bench.PR_ITERS)
for i in range(runs):
    st = r.step(i)
ctx.synchronize()
print(workload, "seconds", st["seconds"], "iterations", st["iterations"], "launches", st["kernel_launches"])
r.close()
