import os, sys, numpy as np
sys.path.insert(0, ".")
import vectorgraphlibrary_b200 as vgl
from vectorgraphlibrary_b200.dist import pick_sources
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 26
with vgl.Context(0) as ctx:
    ds, dd = ctx.generate_edges(1, scale, 16)
    g = vgl.Graph.from_edges(ctx, 1 << scale, ds, dd, vgl.GRAPH_WITH_INCOMING); ds.free(); dd.free()
    ptr, _ = g.layout(); fwd = g.orig_to_sorted()
    srcs = [int(fwd[x]) for x in pick_sources(1 << scale, np.diff(ptr)[fwd], 4, vgl.MASTER_SEED)]
    out = ctx.empty(1 << scale, np.int32)
    for i in range(4):
        if i == 3: os.environ["VGLB_BFS_TRACE"] = "1"
        _, st = g.bfs(srcs[i % 4], True, out)
        print("ms %.3f levels %d bu %d launches %d" % (st.seconds * 1e3, st.iterations, st.bottom_up_levels, st.kernel_launches), flush=True)
    print(g.tiers())
