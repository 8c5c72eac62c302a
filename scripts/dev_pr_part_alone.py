"""dev: one rank's PageRank sweep of a P-way partitioned graph timed alone on one GPU (detached communicator, no exchange)."""
import os, sys, numpy as np
sys.path.insert(0, ".")
os.environ["VGLB_PR_NO_EXCHANGE"] = "1"
import vectorgraphlibrary_b200 as vgl
with vgl.Context(0) as ctx:
    for P in [int(x) for x in (sys.argv[1:] or ["1", "2", "4", "8"])]:
        scale = 24 + int(np.log2(P))
        comm = vgl.Comm(ctx, 0, P, detached=True)
        g = vgl.Graph.from_generator_partitioned(ctx, comm, 0, scale, 16)
        out = ctx.empty(g.V, np.float32)
        for i in range(3):
            _, st = g.pagerank(20, 0.85, out)
        print(f"P={P} scale={scale} rows={g.V} edges={g.E}: {st.seconds * 1e3 / 20:.4f} ms/sweep (kernel only)", flush=True)
        out.free(); g.free(); comm.close()
