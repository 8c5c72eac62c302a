"""dev: the lambda-API algorithms (tests/shim/shim_algorithms) against the fused entry points on the same graph."""
import subprocess, sys, tempfile, numpy as np
sys.path.insert(0, ".")
import vectorgraphlibrary_b200 as vgl
from vectorgraphlibrary_b200.dist import pick_sources
kind, scale, ef = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
V = 1 << scale
with vgl.Context(0) as ctx:
    ds, dd = ctx.generate_edges(kind, scale, ef)
    g = vgl.Graph.from_edges(ctx, V, ds, dd, vgl.GRAPH_WITH_INCOMING)
    ptr, _ = g.layout(); fwd = g.orig_to_sorted()
    s_orig = pick_sources(V, np.diff(ptr)[fwd], 1, vgl.MASTER_SEED)[0]
    w = g.synthetic_weights(77)
    t = {}
    for i in range(3):
        t["bfs"] = g.bfs(int(fwd[s_orig]), False)[1].seconds * 1e3   # top-down only, like the lambda version
        t["bfs_do"] = g.bfs(int(fwd[s_orig]), True)[1].seconds * 1e3
        t["sssp"] = g.sssp(w, int(fwd[s_orig]))[1].seconds * 1e3
        t["cc"] = g.cc()[1].seconds * 1e3
        t["pagerank"] = g.pagerank(20)[1].seconds * 1e3
    print("fused ms:", {k: round(v, 3) for k, v in t.items()}, flush=True)
with tempfile.TemporaryDirectory() as d:
    p = subprocess.run(["tests/shim/shim_algorithms", str(kind), str(scale), str(ef), str(vgl.MASTER_SEED), str(s_orig), "77", "20", d],
                       capture_output=True, text=True)
    print(p.stdout.strip()[-400:], p.stderr[-300:])
