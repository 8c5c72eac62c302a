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
