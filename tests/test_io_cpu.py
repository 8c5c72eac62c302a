"""CPU side of the on-disk formats: the .el_container writer of libvgl_b200 is host-only code, so its bytes are checked here
(against the documented layout always, against the unmodified reference's writer when oracle/_ref is built)."""
import os

import numpy as np


def test_el_container_writer_bytes(vgl, oracle, tmp_path):
    src, dst = oracle.generate_edges(oracle.GEN_RMAT, 9, 6, 123)
    V = 1 << 9
    ours = str(tmp_path / "ours.el_container")
    vgl.save_el_container(ours, V, src, dst)
    raw = open(ours, "rb").read()
    # edges_container.h:58-76: int32 V, int64 E, int32 EDGES_CONTAINER (= 4, framework_types.h:49-57), src[E], dst[E]
    expect = np.int32(V).tobytes() + np.int64(len(src)).tobytes() + np.int32(4).tobytes() + src.tobytes() + dst.tobytes()
    assert raw == expect
    if oracle.ref_available("bfs") and hasattr(oracle.ref_lib("bfs"), "vglref_edges_save"):
        theirs = str(tmp_path / "ref.el_container")
        oracle.ref_save_edges(theirs, V, src, dst)
        assert open(theirs, "rb").read() == raw
        rg = oracle.RefGraph.from_edges_file(ours)  # the reference imports our file: same layout as from the arrays
        ptr, adj, fwd, _, _ = rg.layout(0)
        og = oracle.OracleGraph(V, src, dst)
        assert np.array_equal(ptr, og.row_ptr) and np.array_equal(adj, og.adj) and np.array_equal(fwd, og.fwd)
        rg.close()


def test_el_container_writer_rejects_bad_arguments(vgl, tmp_path):
    import pytest
    e = np.zeros(4, np.int32)
    with pytest.raises(vgl.VglbError):
        vgl.save_el_container(str(tmp_path / "no_such_dir" / "x.el_container"), 8, e, e)
    with pytest.raises(vgl.VglbError):
        vgl.save_el_container(str(tmp_path / "x.el_container"), 0, e, e)


def _parse_vgl(raw, V, E):
    """The VGL graph file as documented in csrc/graph_io.cu, parsed with numpy: header + outgoing + incoming container."""
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype, n, off)
        off += a.nbytes
        return a

    hdr = (int(take("<i4", 1)[0]), int(take("<i8", 1)[0]), int(take("<i4", 1)[0]))
    containers = []
    for _ in range(2):
        c = {"V": int(take("<i4", 1)[0]), "E": int(take("<i8", 1)[0]), "format": int(take("<i4", 1)[0])}
        c["ptr"], c["adj"] = take("<i8", V + 1), take("<i4", E)
        c["fwd"], c["bwd"], c["edge_order"] = take("<i4", V), take("<i4", V), take("<i8", E)
        containers.append(c)
    assert off == len(raw)
    return hdr, containers


def test_vgl_file_layout_of_the_reference(oracle, tmp_path):
    """Pins the format description on the CPU: a file written by the unmodified reference parses into the oracle's layout —
    including the quirk that the incoming container is built from the edge list in OUTGOING-CSR order, because
    VectorCSRGraph::import leaves the caller's EdgesContainer sorted (vect_csr/import.hpp:299-326)."""
    import pytest
    if not oracle.ref_available("bfs") or not hasattr(oracle.ref_lib("bfs"), "vglref_graph_save"):
        pytest.skip("oracle/_ref with the file-format entry points was not built in this environment")
    src, dst = oracle.generate_edges(oracle.GEN_KRONECKER, 9, 8, 99)
    V, E = 1 << 9, len(src)
    rg = oracle.RefGraph(V, src, dst, "bfs")
    path = str(tmp_path / "ref.vgl")
    rg.save(path)
    rg.close()
    hdr, (out, inc) = _parse_vgl(open(path, "rb").read(), V, E)
    assert hdr == (V, E, 1) and (out["V"], out["E"], out["format"]) == (V, E, 1) and (inc["V"], inc["E"], inc["format"]) == (V, E, 1)
    og = oracle.OracleGraph(V, src, dst, want_edge_order=True)
    assert np.array_equal(out["ptr"], og.row_ptr) and np.array_equal(out["adj"], og.adj)
    assert np.array_equal(out["fwd"], og.fwd) and np.array_equal(out["bwd"][out["fwd"]], np.arange(V))
    assert np.array_equal(out["edge_order"], og.edge_order)
    # the edge list as the outgoing import leaves it: outgoing-CSR order, ORIGINAL ids; the incoming container is the import of
    # its transpose
    rows = np.repeat(np.arange(V, dtype=np.int32), np.diff(og.row_ptr))
    s_sorted_order, d_sorted_order = out["bwd"][rows], out["bwd"][og.adj]
    ig = oracle.OracleGraph(V, d_sorted_order, s_sorted_order, want_edge_order=True)
    assert np.array_equal(inc["ptr"], ig.row_ptr) and np.array_equal(inc["adj"], ig.adj)
    assert np.array_equal(inc["fwd"], ig.fwd) and np.array_equal(inc["edge_order"], ig.edge_order)
    # ... and NOT the import of the transposed input in its original order (same rows, different order inside them)
    naive = oracle.OracleGraph(V, dst, src, want_edge_order=True)
    assert np.array_equal(inc["ptr"], naive.row_ptr) and not np.array_equal(inc["edge_order"], naive.edge_order)
