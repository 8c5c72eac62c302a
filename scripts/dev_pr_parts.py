"""Developer check (GPU): time of the heavy region and of the tail tiers of the PageRank sweep alone (VGLB_PR_ONLY)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import vectorgraphlibrary_b200 as vgl
ctx = vgl.Context(0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dsrc, ddst = ctx.generate_edges(vgl.GEN_RMAT, scale, 16)
g = vgl.Graph.from_edges(ctx, 1 << scale, dsrc, ddst, 0)
for only, inter in (("", "1"), ("", "0"), ("1", "0"), ("2", "0"), ("3", "0")):
    os.environ.pop("VGLB_PR_ONLY", None)
    if only:
        os.environ["VGLB_PR_ONLY"] = only
    os.environ["VGLB_PR_INTERLEAVE"] = inter
    best = 1e9
    for rep in range(4):
        _, st = g.pagerank(20)
        best = min(best, st.seconds)
    print("only=%s interleave=%s: %.4f ms/sweep" % (only, inter, best * 1e3 / 20), flush=True)
