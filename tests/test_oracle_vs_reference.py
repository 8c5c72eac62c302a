"""CPU: the C oracle against the UNMODIFIED reference (oracle/_ref/libvgl_ref_*.so, compiled from /root/reference by
oracle/Makefile) on fresh seeds — the live half of the pinning (the fixtures are the frozen half).
Skipped where oracle/_ref was never built; nothing here reads /root/reference at run time."""
import numpy as np
import pytest


def _need_ref(O, profile):
    if not O.ref_available(profile):
        pytest.skip("oracle/_ref not built in this environment")


CASES = [(0, 12, 8, 0x1234), (1, 11, 16, 0x77), (2, 11, 32, 0xBEEF), (0, 16, 16, 0xA1), (1, 17, 16, 0xA2), (2, 16, 32, 0xA3)]


@pytest.mark.parametrize("kind,scale,ef,seed", CASES)
def test_layout_bfs_cc(oracle, kind, scale, ef, seed):
    O = oracle
    _need_ref(O, "bfs")
    V = 1 << scale
    src, dst = O.generate_edges(kind, scale, ef, seed)
    rg = O.RefGraph(V, src, dst, "bfs")
    og = O.OracleGraph(V, src, dst)
    ptr, adj, fwd, ve, vc = rg.layout(0)
    assert np.array_equal(ptr, og.row_ptr) and np.array_equal(adj, og.adj) and np.array_equal(fwd, og.fwd)
    assert vc == og.thresholds(1 << 30, 64)[1]  # apps/bfs/bfs.cpp:5 VECTOR_CORE_THRESHOLD_VALUE 2*VECTOR_LENGTH
    outdeg = np.bincount(src, minlength=V)
    for s in O.pick_sources(V, outdeg, 2, seed):
        ref_lv, _ = rg.bfs(s, 0)
        assert np.array_equal(og.bfs(s)[0], ref_lv)
    rg.close()


@pytest.mark.parametrize("kind,scale,ef,seed", CASES[:2])
def test_cc(oracle, kind, scale, ef, seed):
    O = oracle
    _need_ref(O, "cc")
    V = 1 << scale
    src, dst = O.generate_edges(kind, scale, ef, seed)
    s2, d2 = O.symmetrize(src, dst)
    rg = O.RefGraph(V, s2, d2, "cc")
    ref_lab, _ = rg.cc()
    assert np.array_equal(O.OracleGraph(V, s2, d2).cc()[0], ref_lab)
    rg.close()


@pytest.mark.parametrize("kind,scale,ef,seed", CASES)
def test_sssp(oracle, kind, scale, ef, seed):
    O = oracle
    _need_ref(O, "sssp")
    V = 1 << scale
    src, dst = O.generate_edges(kind, scale, ef, seed)
    rg = O.RefGraph(V, src, dst, "sssp")
    og = O.OracleGraph(V, src, dst)
    outdeg = np.bincount(src, minlength=V)
    s = O.pick_sources(V, outdeg, 1, seed)[0]
    ref_seq, _ = rg.sssp(s, seed ^ 0x5555, 0)       # seq_dijkstra: authoritative
    ref_aa, _ = rg.sssp(s, seed ^ 0x5555, 1)        # vgl_dijkstra ALL_ACTIVE PUSH: self-healing cross-check
    mine, _ = og.sssp(s, seed ^ 0x5555)
    assert np.array_equal(ref_seq.view(np.uint32), ref_aa.view(np.uint32))
    assert np.array_equal(mine.view(np.uint32), ref_seq.view(np.uint32))
    rg.close()


@pytest.mark.parametrize("kind,scale,ef,seed", CASES[:2])
def test_pagerank(oracle, kind, scale, ef, seed):
    O = oracle
    _need_ref(O, "pr")
    V = 1 << scale
    src, dst = O.generate_edges(kind, scale, ef, seed)
    rg = O.RefGraph(V, src, dst, "pr")
    og = O.OracleGraph(V, src, dst)
    ref, _ = rg.pagerank(20)
    mine = og.pagerank_f32(20, rg.threads())
    assert O.rel_l1(mine, ref) <= 1e-7
    assert O.rel_l1(ref, og.pagerank_f64(20)) <= 5e-6   # the reference itself vs fp64 truth
    rg.close()
