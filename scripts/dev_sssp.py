"""dev: SSSP timing on the BASELINE graph for several VGLB_SSSP_DELTA_SCALE values (parity checked at scale 16)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
scales = sys.argv[1:] or ["0", "1", "2", "4", "8"]
code = r'''
import os, sys, numpy as np
sys.path.insert(0, ".")
import vectorgraphlibrary_b200 as vgl, oracle as O
from vectorgraphlibrary_b200.dist import pick_sources
with vgl.Context(0) as ctx:
    # parity, scale 16
    src, dst = O.generate_edges(2, 16, 32)
    og = O.OracleGraph(1 << 16, src, dst)
    g = vgl.Graph.from_edges(ctx, 1 << 16, src, dst)
    w = g.synthetic_weights(7); fwd = g.orig_to_sorted()
    s = O.pick_sources(1 << 16, np.bincount(src, minlength=1 << 16), 1)[0]
    d, st = g.sssp(w, int(fwd[s]))
    ok = np.array_equal(g.to_original(d).view(np.uint32), og.sssp(s, 7)[0].view(np.uint32))
    g.free()
    ds, dd = ctx.generate_edges(2, 24, 32)
    g = vgl.Graph.from_edges(ctx, 1 << 24, ds, dd); ds.free(); dd.free()
    w = g.synthetic_weights(vgl.MASTER_SEED ^ 0x5555)
    ptr, _ = g.layout(); fwd = g.orig_to_sorted()
    srcs = [int(fwd[x]) for x in pick_sources(1 << 24, np.diff(ptr)[fwd], 4, vgl.MASTER_SEED)]
    out = ctx.empty(1 << 24, np.float32)
    ts = []
    for i in range(6):
        _, st = g.sssp(w, srcs[i % 4], out)
        if i >= 2: ts.append(st.seconds)
    print("scale", os.environ.get("VGLB_SSSP_DELTA_SCALE"), "parity", ok, "ms %.2f" % (1e3 * np.mean(ts)), "rounds", st.iterations,
          "launches", st.kernel_launches, "edges/E %.2f" % (st.edges_inspected / g.E), "GTEPS %.1f" % (g.E / np.mean(ts) / 1e9), flush=True)
'''
for sc in scales:
    env = dict(os.environ, VGLB_SSSP_DELTA_SCALE=sc)
    subprocess.run([sys.executable, "-c", code], env=env)
