"""The reference's on-disk formats (csrc/graph_io.cu): files written by libvgl_b200 must be byte-identical to the ones
the unmodified reference writes, and load into the reference; files written by the reference must load into the device
layout and give the reference's results. (.el_container: edges_container.h:58-99; .vgl: vgl_graph.hpp:109-161,
vect_csr_graph.hpp:141-180.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref_or_skip(oracle):
    if not oracle.ref_available("bfs") or not hasattr(oracle.ref_lib("bfs"), "vglref_graph_save"):
        pytest.skip("oracle/_ref with the file-format entry points was not built in this environment")


def test_el_container_round_trip(vgl, ctx, oracle, golden, tmp_path):
    name, g = golden
    src, dst = oracle.generate_edges(int(g["kind"]), int(g["scale"]), int(g["edge_factor"]), int(g["seed"]))
    V = 1 << int(g["scale"])
    ours = str(tmp_path / "ours.el_container")
    vgl.save_el_container(ours, V, src, dst)
    raw = open(ours, "rb").read()
    assert len(raw) == 16 + 8 * len(src)
    assert np.frombuffer(raw[:4], "<i4")[0] == V and np.frombuffer(raw[4:12], "<i8")[0] == len(src) and np.frombuffer(raw[12:16], "<i4")[0] == 4
    G = vgl.Graph.import_el_container(ctx, ours, vgl.GRAPH_WITH_INCOMING)
    ptr, adj = G.layout()
    assert np.array_equal(ptr, g["out_ptr"]) and np.array_equal(adj, g["out_adj"]) and np.array_equal(G.orig_to_sorted(), g["out_fwd"])
    fwd = G.orig_to_sorted()
    lv, _ = G.bfs(int(fwd[int(g["sources"][0])]), True)
    assert np.array_equal(G.to_original(lv), g["bfs_levels"][0])
    G.free()
    with pytest.raises(vgl.VglbError):
        vgl.Graph.import_el_container(ctx, str(tmp_path / "missing.el_container"))
    bad = str(tmp_path / "bad.el_container")
    open(bad, "wb").write(raw[:12] + np.int32(1).tobytes() + raw[16:])
    with pytest.raises(vgl.VglbError, match="incorrect type of graph"):
        vgl.Graph.import_el_container(ctx, bad)
    open(bad, "wb").write(raw[:-8])
    with pytest.raises(vgl.VglbError, match="truncated"):
        vgl.Graph.import_el_container(ctx, bad)


def test_el_container_matches_reference_writer_and_reader(vgl, ctx, oracle, golden, tmp_path):
    _ref_or_skip(oracle)
    name, g = golden
    src, dst = oracle.generate_edges(int(g["kind"]), int(g["scale"]), int(g["edge_factor"]), int(g["seed"]))
    V = 1 << int(g["scale"])
    ours, theirs = str(tmp_path / "ours.el_container"), str(tmp_path / "ref.el_container")
    vgl.save_el_container(ours, V, src, dst)
    oracle.ref_save_edges(theirs, V, src, dst)
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    rg = oracle.RefGraph.from_edges_file(ours)  # the reference imports OUR file
    ptr, adj, fwd, _, _ = rg.layout(0)
    assert np.array_equal(ptr, g["out_ptr"]) and np.array_equal(adj, g["out_adj"]) and np.array_equal(fwd, g["out_fwd"])
    rg.close()


def test_vgl_graph_file_byte_identical_and_loadable_both_ways(vgl, ctx, oracle, golden, tmp_path):
    _ref_or_skip(oracle)
    name, g = golden
    src, dst = oracle.generate_edges(int(g["kind"]), int(g["scale"]), int(g["edge_factor"]), int(g["seed"]))
    V = 1 << int(g["scale"])
    ours, theirs = str(tmp_path / "ours.vgl"), str(tmp_path / "ref.vgl")
    vgl.save_vgl(ctx, ours, V, src, dst)
    rg = oracle.RefGraph(V, src, dst, "bfs")
    rg.save(theirs)
    a, b = open(ours, "rb").read(), open(theirs, "rb").read()
    assert len(a) == len(b) == 16 + 2 * (16 + 8 * (V + 1) + 4 * len(src) + 8 * V + 8 * len(src))
    assert a == b, "VGL graph file differs from the reference's at byte %d" % next(i for i in range(len(a)) if a[i] != b[i])
    rg.close()
    # the reference loads OUR file and runs its own BFS on it
    s0 = int(g["sources"][0])
    rl = oracle.RefGraph.load(ours)
    assert (rl.V, rl.E) == (V, len(src))
    assert np.array_equal(rl.bfs(s0, 0)[0], g["bfs_levels"][0])
    rl.close()
    # we load THEIR file (and one written with device-resident edges) into the device layout
    also = str(tmp_path / "ours_dev.vgl")
    vgl.save_vgl(ctx, also, V, ctx.from_numpy(src), ctx.from_numpy(dst))
    assert open(also, "rb").read() == b
    G = vgl.Graph.load_vgl(ctx, theirs, vgl.GRAPH_WITH_INCOMING)
    ptr, adj = G.layout()
    assert np.array_equal(ptr, g["out_ptr"]) and np.array_equal(adj, g["out_adj"]) and np.array_equal(G.orig_to_sorted(), g["out_fwd"])
    H = vgl.Graph.from_edges(ctx, V, src, dst, vgl.GRAPH_WITH_INCOMING)
    for x, y in zip(G.layout(incoming=True), H.layout(incoming=True)):
        assert np.array_equal(x, y)  # the derived incoming direction equals the builder's
    fwd = G.orig_to_sorted()
    for dopt in (False, True):
        lv, _ = G.bfs(int(fwd[s0]), dopt)
        assert np.array_equal(G.to_original(lv), g["bfs_levels"][0])
    ranks, _ = G.pagerank(int(g["pr_iters"]))
    assert oracle.rel_l1(G.to_original(ranks), g["pr_ranks"]) <= 1e-6
    G.free(); H.free()
    with pytest.raises(vgl.VglbError):
        vgl.Graph.load_vgl(ctx, str(tmp_path / "ours.el_container"))
