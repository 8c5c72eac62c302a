"""One rank of the multi-GPU parity test (launched by torchrun from tests/test_gpu_partition.py, or by hand:
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/part_worker.py).
Every rank builds its part, runs the four partitioned algorithms and checks the WHOLE result against tests/golden/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OMP_NUM_THREADS", "8")


def main():
    import torch
    import oracle as O
    import vectorgraphlibrary_b200 as vgl
    from vectorgraphlibrary_b200 import dist as vdist
    from test_gpu_partition import _check_algorithms, _edges

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tcomm = vdist.Communicator.from_env(local)
    ctx = vgl.Context(local)
    comm = vgl.Comm(ctx, rank, world, exchange=lambda b: tcomm.broadcast_bytes(b, vgl.UNIQUE_ID_BYTES, 0))
    for name in ("rmat_s8_ef4", "kron_s10_ef16", "ru_s10_ef32", "rmat_s11_ef8"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        src, dst = _edges(O, g)
        V = 1 << int(g["scale"])
        G = vgl.Graph.from_edges_partitioned(ctx, comm, V, src, dst, vgl.GRAPH_WITH_INCOMING)
        _check_algorithms(vgl, ctx, O, comm, g, G)
        ptr, adj = G.layout()
        iptr, iadj = G.layout(incoming=True)
        H = vgl.Graph.from_csr_partitioned(ctx, comm, V, ptr, adj, G.orig_to_sorted(), iptr, iadj)
        assert H.E_global == len(src)
        _check_algorithms(vgl, ctx, O, comm, g, H)
        s2, d2 = O.symmetrize(src, dst)
        G2 = vgl.Graph.from_edges_partitioned(ctx, comm, V, s2, d2)
        lab2, _ = G2.cc()
        assert np.array_equal(G2.to_original(lab2), g["cc_symmetric"])
        for X in (G, H, G2):
            X.free()
    # a bigger case against the C oracle: Kronecker scale 16, generator-built parts
    scale, ef = 16, 16
    V = 1 << scale
    src, dst = O.generate_edges(O.GEN_KRONECKER, scale, ef)
    og = O.OracleGraph(V, src, dst)
    G = vgl.Graph.from_generator_partitioned(ctx, comm, vgl.GEN_KRONECKER, scale, ef, vgl.GRAPH_WITH_INCOMING)
    fwd = G.orig_to_sorted()
    for s in O.pick_sources(V, np.bincount(src, minlength=V), 2):
        lv, st = G.bfs(int(fwd[s]), True)
        assert np.array_equal(G.to_original(lv), og.bfs(s)[0])
        w = G.synthetic_weights(7)
        d, _ = G.sssp(w, int(fwd[s]))
        assert np.array_equal(G.to_original(d).view(np.uint32), og.sssp(s, 7)[0].view(np.uint32))
    lab, _ = G.cc()
    assert np.array_equal(G.to_original(lab), og.cc()[0])
    ranks, _ = G.pagerank(20)
    r = G.to_original(ranks)
    assert O.rel_l1(r, og.pagerank_f64(20)) <= 1e-6
    # the reference-order fp32 oracle drifts from fp64 truth as V grows; same rule as test_gpu_parity
    r32 = og.pagerank_f32(20, 8)
    if O.rel_l1(r32, og.pagerank_f64(20)) <= 5e-7:
        assert O.rel_l1(r, r32) <= 1e-6
    # the peer-store exchange (sweep epilogue writes into the peers' vectors) must give the same ranks, run after run
    G.set_exchange(vgl.EXCHANGE_P2P)
    for _ in range(3):
        ranks2, _ = G.pagerank(20)
        assert O.rel_l1(G.to_original(ranks2), og.pagerank_f64(20)) <= 1e-6
        assert O.rel_l1(G.to_original(ranks2), r) <= 1e-7
    G.set_exchange(vgl.EXCHANGE_NCCL)
    G.free()
    comm.close()
    ctx.close()
    tcomm.close()
    print(f"PART_WORKER_OK rank {rank}/{world}", flush=True)


if __name__ == "__main__":
    main()
