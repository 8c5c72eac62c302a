import os, sys, numpy as np
sys.path.insert(0, ".")
import vectorgraphlibrary_b200 as vgl
from vectorgraphlibrary_b200.dist import pick_sources
with vgl.Context(0) as ctx:
    ds, dd = ctx.generate_edges(2, 24, 32)
    g = vgl.Graph.from_edges(ctx, 1 << 24, ds, dd); ds.free(); dd.free()
    w = g.synthetic_weights(vgl.MASTER_SEED ^ 0x5555)
    ptr, _ = g.layout(); fwd = g.orig_to_sorted()
    srcs = [int(fwd[x]) for x in pick_sources(1 << 24, np.diff(ptr)[fwd], 4, vgl.MASTER_SEED)]
    out = ctx.empty(1 << 24, np.float32)
    for i in range(2):
        if i == 1: os.environ["VGLB_SSSP_TRACE"] = "1"
        _, st = g.sssp(w, srcs[0], out)
    print("ms", st.seconds * 1e3, g.tiers())
