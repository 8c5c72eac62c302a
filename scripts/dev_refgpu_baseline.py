"""dev (GPU): the reference's OWN CUDA backend (vgl_compute_api/gpu recompiled for sm_100a, oracle/_ref/libvgl_refgpu.so) as a second
stated baseline, beside (a) the reference's unmodified algorithm sources on this repo's backend through the include-path overlay
(libvgl_dropin.so) and (b) the fused entry points of libvgl_b200 — same graph, same sources. Times are the algorithm calls only
(graph import / upload excluded), best of 3. Usage: python scripts/dev_refgpu_baseline.py [scale] [edge factor]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle as O
import vectorgraphlibrary_b200 as vgl

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 16
V = 1 << scale
src, dst = O.generate_edges(0, scale, ef, 0xB200)
E = len(src)
outdeg = np.bincount(src, minlength=V)
sources = [int(s) for s in O.pick_sources(V, outdeg, 3, 0xB200)]
print(f"RMAT scale {scale} ef {ef}: V {V} E {E}; GTEPS = E (x 20 sweeps for PageRank) / time", flush=True)
best = lambda f: min(f() for _ in range(3))
rows = {}
for which in ("refgpu", "dropin"):
    if not O.gpu_ref_available(which):
        print(which, "not built"); continue
    G = O.GpuRefGraph(V, src, dst, which)
    r = {}
    r["pagerank"] = best(lambda: G.pagerank(20)[1])
    r["bfs"] = float(np.mean([best(lambda s=s: G.bfs(s)[1]) for s in sources]))
    r["cc"] = best(lambda: G.cc()[1])
    r["sssp"] = float(np.mean([best(lambda s=s: G.sssp(s, 0x5555, 2)[1]) for s in sources[:1]]))
    G.close()
    rows[which] = r
    print(which, {k: round(v * 1e3, 3) for k, v in r.items()}, "ms", flush=True)
with vgl.Context(0) as ctx:
    G = vgl.Graph.from_edges(ctx, V, src, dst, vgl.GRAPH_WITH_INCOMING)
    fwd = G.orig_to_sorted()
    w = G.synthetic_weights(0x5555)
    r = {}
    r["pagerank"] = best(lambda: G.pagerank(20)[1].seconds)
    r["bfs"] = float(np.mean([best(lambda s=s: G.bfs(int(fwd[s]), direction_optimising=False)[1].seconds) for s in sources]))
    r["bfs_do"] = float(np.mean([best(lambda s=s: G.bfs(int(fwd[s]), direction_optimising=True)[1].seconds) for s in sources]))
    r["cc"] = best(lambda: G.cc()[1].seconds)
    r["sssp"] = best(lambda: G.sssp(w, int(fwd[sources[0]]))[1].seconds)
    rows["fused"] = r
    print("fused", {k: round(v * 1e3, 3) for k, v in r.items()}, "ms", flush=True)
print("\nGTEPS (PageRank: 20 sweeps; BFS top-down as in bfs.hpp; the fused bfs_do line is direction-optimising):")
for name, r in rows.items():
    print(f"  {name:7s}", {k: round((20 if k == 'pagerank' else 1) * E / v / 1e9, 2) for k, v in r.items()})
