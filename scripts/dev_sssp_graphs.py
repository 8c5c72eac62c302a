"""dev: fused SSSP on several graph families (near/far robustness), parity at small scale, time at scale."""
import sys, numpy as np
sys.path.insert(0, ".")
import vectorgraphlibrary_b200 as vgl, oracle as O
from vectorgraphlibrary_b200.dist import pick_sources
with vgl.Context(0) as ctx:
    for kind, scale, ef in ((0, 15, 16), (1, 15, 16), (2, 15, 32)):
        src, dst = O.generate_edges(kind, scale, ef); V = 1 << scale
        og = O.OracleGraph(V, src, dst); g = vgl.Graph.from_edges(ctx, V, src, dst)
        w = g.synthetic_weights(7); fwd = g.orig_to_sorted()
        for s in O.pick_sources(V, np.bincount(src, minlength=V), 3):
            d, st = g.sssp(w, int(fwd[s]))
            assert np.array_equal(g.to_original(d).view(np.uint32), og.sssp(s, 7)[0].view(np.uint32)), (kind, s)
        g.free()
    print("parity ok", flush=True)
    for kind, scale, ef in ((0, 22, 16), (1, 22, 16), (2, 24, 32), (0, 24, 16)):
        ds, dd = ctx.generate_edges(kind, scale, ef); V = 1 << scale
        g = vgl.Graph.from_edges(ctx, V, ds, dd); ds.free(); dd.free()
        w = g.synthetic_weights(77); ptr, _ = g.layout(); fwd = g.orig_to_sorted()
        srcs = [int(fwd[x]) for x in pick_sources(V, np.diff(ptr)[fwd], 4, vgl.MASTER_SEED)]
        ts = []
        for i in range(6):
            _, st = g.sssp(w, srcs[i % 4])
            if i >= 2: ts.append(st.seconds)
        print(f"kind {kind} scale {scale} ef {ef}: {1e3 * np.mean(ts):.2f} ms rounds {st.iterations} edges/E {st.edges_inspected / g.E:.2f} GTEPS {g.E / np.mean(ts) / 1e9:.1f}", flush=True)
        g.free()
