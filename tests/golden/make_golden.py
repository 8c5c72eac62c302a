"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libvgl_ref_pr.so).

Run in the build container only (needs /root/reference to have been compiled by `make -C oracle ref`):

    OMP_NUM_THREADS=8 python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these fixtures are outputs of the reference's
multicore build on our deterministic synthetic graphs (include/vglb_synth.h). Each file records the generator
parameters, the reference's VectCSR layout, and BFS / SSSP / PageRank / CC results in ORIGINAL vertex order.
PageRank goldens record the OpenMP thread count T they were produced with.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
os.environ.setdefault("OMP_NUM_THREADS", "8")
import oracle as O  # noqa: E402

CASES = [
    # name, kind, scale, edge factor
    ("kron_s10_ef16", O.GEN_KRONECKER, 10, 16),
    ("rmat_s11_ef8", O.GEN_RMAT, 11, 8),
    ("ru_s10_ef32", O.GEN_UNIFORM, 10, 32),
    ("rmat_s8_ef4", O.GEN_RMAT, 8, 4),
]
WEIGHT_SEED = 0xB200 ^ 0x5555
PR_ITERS = 20


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, kind, scale, ef in CASES:
        V = 1 << scale
        src, dst = O.generate_edges(kind, scale, ef)
        rg = O.RefGraph(V, src, dst, "pr")
        out_ptr, out_adj, out_fwd, _, _ = rg.layout(0)
        in_ptr, in_adj, in_fwd, _, _ = rg.layout(1)
        outdeg = np.bincount(src, minlength=V)
        sources = np.array(O.pick_sources(V, outdeg, 3), np.int32)
        bfs = np.stack([rg.bfs(int(s), 0)[0] for s in sources])
        bfs_seq = np.stack([rg.bfs(int(s), 1)[0] for s in sources])
        assert np.array_equal(bfs, bfs_seq)
        sssp = np.stack([rg.sssp(int(s), WEIGHT_SEED, 0)[0] for s in sources])
        sssp_aa = np.stack([rg.sssp(int(s), WEIGHT_SEED, 1)[0] for s in sources])
        assert np.array_equal(sssp.view(np.uint32), sssp_aa.view(np.uint32))
        pr, _ = rg.pagerank(PR_ITERS)
        cc_dir, _ = rg.cc()
        s2, d2 = O.symmetrize(src, dst)
        rg2 = O.RefGraph(V, s2, d2, "pr")
        cc_sym, _ = rg2.cc()
        np.savez_compressed(
            os.path.join(here, name + ".npz"),
            kind=kind, scale=scale, edge_factor=ef, seed=O.MASTER_SEED, weight_seed=WEIGHT_SEED,
            edges_checksum=np.array([int(src.astype(np.int64).sum()), int(dst.astype(np.int64).sum()),
                                     int((src.astype(np.int64) * 31 + dst).sum())], np.int64),
            out_ptr=out_ptr, out_adj=out_adj, out_fwd=out_fwd, in_ptr=in_ptr, in_adj=in_adj, in_fwd=in_fwd,
            sources=sources, bfs_levels=bfs, sssp_dist=sssp, pr_ranks=pr, pr_iters=PR_ITERS,
            pr_threads=rg.threads(), cc_directed=cc_dir, cc_symmetric=cc_sym)
        print(name, "V", V, "E", len(src), "sources", sources.tolist(), "pr sum", float(pr.sum()))
        rg.close()
        rg2.close()


if __name__ == "__main__":
    main()
