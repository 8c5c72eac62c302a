"""Device-side verification (csrc/verify.cu = vgl_runtime/helpers/verify_results/verify_results.h on the device), the EdgesArray
outgoing -> incoming mirror (VGL_Graph::copy_outgoing_to_incoming_edges) and add_group_of_vertices through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_verify_results_counts_like_the_reference(vgl, ctx):
    rng = np.random.default_rng(11)
    n = 100_003
    a = rng.integers(-5, 1000, n).astype(np.int32)
    b = a.copy()
    wrong = rng.choice(n, 137, replace=False)
    b[wrong] += 1
    da, db = ctx.from_numpy(a), ctx.from_numpy(b)
    assert vgl.verify_results(ctx, da, da) == 0
    assert vgl.verify_results(ctx, da, db) == 137
    x = rng.random(n).astype(np.float32)
    y = x.copy()
    y[:50] += np.float32(1e-3)       # beyond 100 * FLT_EPSILON
    y[50:100] += np.float32(1e-6)    # within (are_same, verify_results.h:9-13)
    expect = int((np.abs(x - y) > np.finfo(np.float32).eps * 100.0).sum())
    assert vgl.verify_results(ctx, ctx.from_numpy(x), ctx.from_numpy(y)) == expect == 50
    diff, rel, err = vgl.verify_ranking_results(ctx, ctx.from_numpy(y), ctx.from_numpy(x))
    assert abs(diff - float(np.abs(x.astype(np.float64) - y).mean())) < 1e-12 and err == 0
    assert abs(rel - float(np.abs(x.astype(np.float64) - y).sum() / np.abs(x.astype(np.float64)).sum())) < 1e-12
    _, _, err = vgl.verify_ranking_results(ctx, ctx.from_numpy(x + np.float32(1.0)), ctx.from_numpy(x))
    assert err == n  # verify_results.h:127-130


def _equal_components_reference(first, second):
    """verify_results.h:198-254 restated: maps keep the last assignment, each mismatch counts once per direction."""
    f_s, s_f = {}, {}
    for f, s in zip(first.tolist(), second.tolist()):
        f_s[f] = s
        s_f[s] = f
    errors = 0
    for f, s in zip(first.tolist(), second.tolist()):
        errors += f_s[f] != s
        errors += s_f[s] != f
    return errors


def test_equal_components_on_device(vgl, ctx, oracle):
    O = oracle
    V = 1 << 12
    src, dst = O.generate_edges(O.GEN_RMAT, 12, 4, 0xC1)
    s2, d2 = O.symmetrize(src, dst)
    G = vgl.Graph.from_edges(ctx, V, s2, d2)
    lab, _ = G.cc()
    labels = lab.to_numpy()
    # the same partition under a relabelling (what the reference's seq_bfs_based produces: component numbers, not minimum ids)
    uniq, relabelled = np.unique(labels, return_inverse=True)
    other = (relabelled + 1).astype(np.int32)
    assert vgl.equal_components(ctx, lab, ctx.from_numpy(other)) == 0 == _equal_components_reference(labels, other)
    broken = other.copy()
    broken[np.flatnonzero(labels == labels[0])[:3]] = int(other.max()) + 1  # split three vertices off the first component
    assert vgl.equal_components(ctx, lab, ctx.from_numpy(broken)) == _equal_components_reference(labels, broken) > 0
    assert vgl.verify_results(ctx, lab, lab) == 0
    G.free()


def test_edges_array_mirror_out_to_in(vgl, ctx, oracle):
    """Every incoming-CSR position must receive the value of the SAME edge: weights are a function of the edge's end points
    (vglb_edge_weight), so the expected incoming segment can be computed independently."""
    O = oracle
    for kind, scale, ef, with_incoming in ((0, 12, 8, True), (1, 11, 16, False)):
        V = 1 << scale
        src, dst = O.generate_edges(kind, scale, ef, 0xD7)
        G = vgl.Graph.from_edges(ctx, V, src, dst, vgl.GRAPH_WITH_INCOMING if with_incoming else 0)
        w_out = G.synthetic_weights(99)
        w_in = G.mirror_out_to_in(w_out)          # (derives the incoming CSR when the graph was built without it)
        G2 = vgl.Graph(ctx, G.h)                  # refreshed info: the incoming CSR exists now
        iptr, iadj = G2.layout(incoming=True)
        G2.h = None
        # oracle weights of the transposed edges: in-row v lists the sources u of edges u -> v
        og = O.OracleGraph(V, src, dst)
        assert np.array_equal(G.orig_to_sorted(), og.fwd)
        rows = np.repeat(np.arange(V, dtype=np.int64), np.diff(iptr))
        # weight of (u -> v) from the out-CSR the oracle built: look the pair up through a dict of the out edges
        wref = og.weights(99)
        out_rows = np.repeat(np.arange(V, dtype=np.int64), np.diff(og.row_ptr))
        key_out = out_rows * V + og.adj
        order = np.argsort(key_out, kind="stable")
        key_sorted, w_sorted = key_out[order], wref[order]
        key_in = iadj.astype(np.int64) * V + rows
        pos = np.searchsorted(key_sorted, key_in)
        assert np.array_equal(key_sorted[pos], key_in)
        assert np.array_equal(w_in.to_numpy().view(np.uint32), w_sorted[pos].view(np.uint32))
        G.free()


def test_add_group_of_vertices(vgl, ctx, oracle):
    O = oracle
    V = 1 << 12
    src, dst = O.generate_edges(O.GEN_RMAT, 12, 8, 0xD8)
    G = vgl.Graph.from_edges(ctx, V, src, dst)
    ptr, _ = G.layout()
    deg = np.diff(ptr)
    F = vgl.Frontier(G)
    rng = np.random.default_rng(5)
    ids = np.unique(np.concatenate([[0, 1, 2], rng.choice(V, 500, replace=False)])).astype(np.int32)
    F.clear()
    F.add_group_of_vertices(ids[::-1].copy())  # the reference sorts the list; so does the mirror
    fi = F.info()
    assert fi.size == len(ids) and fi.neighbours == int(deg[ids].sum()) and fi.sparsity_type == 2
    assert np.array_equal(F.ids(), ids)
    td, tb = G.tiers()
    assert list(fi.tier_size) == [int((ids < tb[0]).sum()), int(((ids >= tb[0]) & (ids < tb[1])).sum()), int((ids >= tb[1]).sum())]
    bm = F.bitmap()
    expect = np.zeros(V, bool)
    expect[ids] = True
    assert np.array_equal(np.unpackbits(bm.view(np.uint8), bitorder="little")[:V].astype(bool), expect)
    assert F.reduce_sum_i32(ctx.from_numpy(np.ones(V, np.int32))) == len(ids)
    with pytest.raises(vgl.VglbError):
        F.add_group_of_vertices(ids)  # only on an empty frontier, like the reference
    F.clear()
    with pytest.raises(vgl.VglbError):
        vgl._check(vgl.lib().vglb_frontier_set_ids(ctx.h, F.h, np.array([5, 5, 9], np.int32).ctypes.data, 3, 0))  # duplicates
    F.free()
    G.free()
