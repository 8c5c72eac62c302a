"""Developer check (GPU): PageRank sweep variants (VGLB_PR_VARIANT) — parity at scale 16, timing at scale 24."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import vectorgraphlibrary_b200 as vgl
import oracle as O

variants = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,3,4,7,8,15".split(","))]
ctx = vgl.Context(0)
V = 1 << 16
src, dst = O.generate_edges(vgl.GEN_RMAT, 16, 16)
og = O.OracleGraph(V, src, dst)
r64 = og.pagerank_f64(20)
gs = vgl.Graph.from_edges(ctx, V, src, dst, 0)
dsrc, ddst = ctx.generate_edges(vgl.GEN_RMAT, 24, 16)
g = vgl.Graph.from_edges(ctx, 1 << 24, dsrc, ddst, 0)
dsrc.free(); ddst.free()
for v in variants:
    os.environ["VGLB_PR_VARIANT"] = str(v)
    ranks, _ = gs.pagerank(20)
    err = O.rel_l1(gs.to_original(ranks), r64)
    best = 1e9
    for rep in range(4):
        _, st = g.pagerank(20)
        best = min(best, st.seconds)
    print("variant %2d: relL1 vs f64 %.2e   s24: %.4f ms/sweep  %.1f GTEPS  %.0f GB/s (%.1f%% of 6539)" % (
        v, err, best * 1e3 / 20, 20 * g.E / best / 1e9, st.algorithmic_bytes / best / 1e9, st.algorithmic_bytes / best / 1e9 / 65.392), flush=True)
