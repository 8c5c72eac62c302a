"""The column-binned PageRank sweep (csrc/pagerank_bins.cu) against the oracle and against the round-1 warp-task sweep:
the same ranks within the PageRank tolerance on graphs that exercise every part of it — several shared-memory bins and a
non-empty cold bin (more than 32 x 49152 columns), rows longer than one 4096-edge row chunk, graphs without heavy rows, and the
build behind a pinned vglb_graph_from_csr upload (vglb_set_upload_hint)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PR_TOL = 1e-6  # relative L1, north_star


def _ranks(vgl, ctx, V, src, dst, iters, no_bins):
    old = os.environ.pop("VGLB_PR_NO_BINS", None)
    if no_bins:
        os.environ["VGLB_PR_NO_BINS"] = "1"  # read when the graph prepares its sweep
    try:
        G = vgl.Graph.from_edges(ctx, V, src, dst, 0)
        r, st = G.pagerank(iters)
        out = G.to_original(r)
        G.free()
        return out, st
    finally:
        os.environ.pop("VGLB_PR_NO_BINS", None)
        if old is not None:
            os.environ["VGLB_PR_NO_BINS"] = old


@pytest.mark.parametrize("kind,scale,ef", [(0, 21, 16), (1, 18, 16), (2, 16, 32), (0, 12, 4)])
def test_binned_sweep_matches_oracle_and_warp_tasks(vgl, ctx, oracle, kind, scale, ef):
    O = oracle
    V = 1 << scale
    src, dst = O.generate_edges(kind, scale, ef, 0xB1A5 + scale)
    iters = 10
    binned, st = _ranks(vgl, ctx, V, src, dst, iters, no_bins=False)
    tasks, _ = _ranks(vgl, ctx, V, src, dst, iters, no_bins=True)
    ref = O.OracleGraph(V, src, dst).pagerank_f64(iters)
    assert O.rel_l1(binned, ref) <= PR_TOL
    assert O.rel_l1(tasks, ref) <= PR_TOL
    assert O.rel_l1(binned, tasks) <= PR_TOL
    again, _ = _ranks(vgl, ctx, V, src, dst, iters, no_bins=False)
    assert np.array_equal(binned.view(np.uint32), again.view(np.uint32)), "the binned sweep sums in a fixed order"
    assert abs(float(binned.sum()) - 1.0) < 1e-3


def test_bins_built_behind_a_pinned_upload(vgl, ctx, oracle):
    """vglb_set_upload_hint(PAGERANK) + pinned host arrays: vglb_graph_from_csr builds the bins while the rest of the adjacency is
    still in flight; the ranks are those of the graph built from the edge list."""
    O = oracle
    scale, ef, iters = 19, 16, 10  # 8 M edges: the upload goes up in chunks
    V = 1 << scale
    src, dst = O.generate_edges(0, scale, ef, 0xC5B)
    G0 = vgl.Graph.from_edges(ctx, V, src, dst, 0)
    ptr, adj = G0.layout()
    fwd = G0.orig_to_sorted()
    want = G0.to_original(G0.pagerank(iters)[0])
    G0.free()
    h_ptr, h_adj, h_fwd = vgl.pinned_array(V + 1, np.int64), vgl.pinned_array(len(adj), np.int32), vgl.pinned_array(V, np.int32)
    h_ptr[:], h_adj[:], h_fwd[:] = ptr, adj, fwd
    L = vgl.lib()
    try:
        vgl._check(L.vglb_set_upload_hint(ctx.h, vgl.HINT_PAGERANK))
        G = vgl.Graph.from_csr(ctx, h_ptr, h_adj, h_fwd)
        got = G.to_original(G.pagerank(iters)[0])
        G.free()
    finally:
        vgl._check(L.vglb_set_upload_hint(ctx.h, 0))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert O.rel_l1(got, O.OracleGraph(V, src, dst).pagerank_f64(iters)) <= PR_TOL
    with pytest.raises(vgl.VglbError):
        vgl._check(L.vglb_set_upload_hint(ctx.h, 0x40))
