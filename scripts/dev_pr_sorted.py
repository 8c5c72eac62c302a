"""dev: PageRank sweep time on the reference edge order vs rows sorted by column (the partitioned builder's order, P = 1)."""
import sys, numpy as np
sys.path.insert(0, ".")
import vectorgraphlibrary_b200 as vgl
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
with vgl.Context(0) as ctx:
    ds, dd = ctx.generate_edges(0, scale, 16)
    g = vgl.Graph.from_edges(ctx, 1 << scale, ds, dd)
    out = ctx.empty(1 << scale, np.float32)
    for i in range(4):
        _, st = g.pagerank(20, 0.85, out)
    print("reference order : %.4f ms/sweep" % (st.seconds * 1e3 / 20), flush=True)
    r0 = g.to_original(out)
    g.free()
    comm = vgl.Comm(ctx, 0, 1)
    g = vgl.Graph.from_edges_partitioned(ctx, comm, 1 << scale, ds, dd)
    for i in range(4):
        _, st = g.pagerank(20, 0.85, out)
    print("column-sorted   : %.4f ms/sweep" % (st.seconds * 1e3 / 20), flush=True)
    r1 = g.to_original(out)
    print("relL1 between the two", float(np.abs(r0.astype(np.float64) - r1).sum() / np.abs(r0).sum()))
