"""dev: where the end-to-end PageRank step (host CSR -> HBM -> 20 sweeps -> host) spends its time."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import vectorgraphlibrary_b200 as vgl
with vgl.Context(0) as ctx:
    ds, dd = ctx.generate_edges(0, 24, 16)
    g0 = vgl.Graph.from_edges(ctx, 1 << 24, ds, dd); ds.free(); dd.free()
    ptr, adj = g0.layout()
    H = {"ptr": vgl.pinned_array(len(ptr), np.int64), "adj": vgl.pinned_array(len(adj), np.int32), "fwd": vgl.pinned_array(g0.V, np.int32),
         "out": vgl.pinned_array(g0.V, np.float32)}
    H["ptr"][:], H["adj"][:], H["fwd"][:] = ptr, adj, g0.orig_to_sorted()
    g0.free()
    L = vgl.lib()
    vgl._check(L.vglb_set_upload_hint(ctx.h, vgl.HINT_PAGERANK))
    for it in range(4):
        ctx.synchronize(); t0 = time.perf_counter()
        g = vgl.Graph.from_csr(ctx, H["ptr"], H["adj"], H["fwd"]); ctx.synchronize(); t1 = time.perf_counter()
        out = ctx.empty(g.V, np.float32)
        _, st = g.pagerank(20, 0.85, out); ctx.synchronize(); t2 = time.perf_counter()
        vgl._check(L.vglb_memcpy_d2h(ctx.h, H["out"].ctypes.data, out.ptr, H["out"].nbytes)); ctx.synchronize(); t3 = time.perf_counter()
        out.free(); g.free(); ctx.synchronize(); t4 = time.perf_counter()
        print(f"from_csr (H2D {(H['ptr'].nbytes + H['adj'].nbytes + H['fwd'].nbytes) / 1e9:.2f} GB + tiers) {1e3 * (t1 - t0):.1f} ms | pagerank call {1e3 * (t2 - t1):.1f} ms "
              f"(device loop {1e3 * st.seconds:.1f} ms => prepare {1e3 * (t2 - t1 - st.seconds):.1f} ms) | D2H {1e3 * (t3 - t2):.1f} ms | free {1e3 * (t4 - t3):.1f} ms", flush=True)
