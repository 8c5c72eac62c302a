"""Developer check (GPU): time of each degree tier of the PageRank sweep alone (VGLB_PR_TIERMASK)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import vectorgraphlibrary_b200 as vgl
ctx = vgl.Context(0)
dsrc, ddst = ctx.generate_edges(vgl.GEN_RMAT, 24, 16)
g = vgl.Graph.from_edges(ctx, 1 << 24, dsrc, ddst, 0)
ptr, _ = g.layout()
tb = [0] + g.tiers()[1]
os.environ["VGLB_PR_VARIANT"] = sys.argv[1] if len(sys.argv) > 1 else "7"
for mask in [0xff, 1, 2, 4, 8, 16, 32, 64, 0]:
    os.environ["VGLB_PR_TIERMASK"] = str(mask)
    best = 1e9
    for rep in range(3):
        _, st = g.pagerank(20)
        best = min(best, st.seconds)
    edges = 0
    for t in range(7):
        if (mask >> t) & 1:
            hi = tb[t + 1] if t < 6 else g.V
            edges += int(ptr[hi] - ptr[tb[t]])
    print("mask %3d: %.4f ms/sweep  edges %d (%.1f%%)  %.1f Gedge/s" % (mask, best * 1e3 / 20, edges, 100.0 * edges / g.E, edges / (best / 20) / 1e9 if best > 0 else 0), flush=True)
