"""Developer check (GPU): A/B timing of dev/lib_<name>.so builds (scripts/dev_build_variant.sh) — PageRank, scale 24."""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import vectorgraphlibrary_b200 as vgl
import oracle as O
ctx = vgl.Context(0)
dsrc, ddst = ctx.generate_edges(vgl.GEN_RMAT, 24, 16)
g = vgl.Graph.from_edges(ctx, 1 << 24, dsrc, ddst, 0)
best = 1e9
times = []
for rep in range(6):
    _, st = g.pagerank(20)
    times.append(round(st.seconds * 50, 4))
    best = min(best, st.seconds)
print(times)
V = 1 << 16
src, dst = O.generate_edges(vgl.GEN_RMAT, 16, 16)
og = O.OracleGraph(V, src, dst)
gs = vgl.Graph.from_edges(ctx, V, src, dst, 0)
ranks, _ = gs.pagerank(20)
err = O.rel_l1(gs.to_original(ranks), og.pagerank_f64(20))
print("%%-10s relL1 %%.2e  %%.4f ms/sweep  %%.1f GTEPS  %%.1f%%%% of 6539 GB/s" %% (os.environ.get("VGLB_NAME"), err, best * 50, 20 * g.E / best / 1e9, st.algorithmic_bytes / best / 65.392e9), flush=True)
''' % ROOT
for name in sys.argv[1:]:
    env = dict(os.environ, VGLB_LIB_PATH=os.path.join(ROOT, "dev", "lib_%s.so" % name), VGLB_NAME=name)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
