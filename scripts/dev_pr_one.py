"""Developer helper (GPU): build the bench graph and run a few PageRank sweeps (for ncu captures with VGLB_PR_* knobs)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import vectorgraphlibrary_b200 as vgl
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = vgl.Context(0)
dsrc, ddst = ctx.generate_edges(vgl.GEN_RMAT, scale, 16)
g = vgl.Graph.from_edges(ctx, 1 << scale, dsrc, ddst, 0)
_, st = g.pagerank(4)
print("ms/sweep", st.seconds * 1e3 / 4)
