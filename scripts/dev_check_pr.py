"""Developer check (GPU): graph build + PageRank parity against the C oracle at small scale, then a timing at scale."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import vectorgraphlibrary_b200 as vgl
import oracle as O

ctx = vgl.Context(0)
for kind, scale, ef in [(vgl.GEN_RMAT, 8, 4), (vgl.GEN_KRONECKER, 14, 16), (vgl.GEN_UNIFORM, 12, 32), (vgl.GEN_RMAT, 18, 16)]:
    V = 1 << scale
    src, dst = O.generate_edges(kind, scale, ef)
    dsrc, ddst = ctx.generate_edges(kind, scale, ef)
    assert np.array_equal(dsrc.to_numpy(), src) and np.array_equal(ddst.to_numpy(), dst), "generator mismatch"
    og = O.OracleGraph(V, src, dst)
    g = vgl.Graph.from_edges(ctx, V, dsrc, ddst, vgl.GRAPH_WITH_INCOMING)
    ptr, adj = g.layout()
    print("scale", scale, "layout", np.array_equal(ptr, og.row_ptr), np.array_equal(adj, og.adj), np.array_equal(g.orig_to_sorted(), og.fwd), g.tiers())
    ranks, st = g.pagerank(20)
    r = g.to_original(ranks)
    r32 = og.pagerank_f32(20, 8)
    r64 = og.pagerank_f64(20)
    print("  pr relL1 vs f32 oracle", O.rel_l1(r, r32), "vs f64", O.rel_l1(r, r64), "oracle vs f64", O.rel_l1(r32, r64), "ms", st.seconds * 1e3, "launches", st.kernel_launches)
    g.free()

for scale in (20, 22, 24):
    V = 1 << scale
    t = time.time()
    dsrc, ddst = ctx.generate_edges(vgl.GEN_RMAT, scale, 16)
    g = vgl.Graph.from_edges(ctx, V, dsrc, ddst, 0)
    ctx.synchronize()
    print("scale", scale, "build s", time.time() - t, "tiers", g.tiers(), "maxdeg", g.info.max_degree)
    for rep in range(3):
        ranks, st = g.pagerank(20)
        print("  pr 20 it: %.3f ms  -> %.1f GTEPS, %.0f GB/s algorithmic" % (st.seconds * 1e3, 20 * g.E / st.seconds / 1e9, st.algorithmic_bytes / st.seconds / 1e9))
    print("  sum", float(ranks.to_numpy().astype(np.float64).sum()))
    g.free(); dsrc.free(); ddst.free()
