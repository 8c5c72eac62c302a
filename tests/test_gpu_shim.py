"""GPU parity of the user-lambda operator API (include/vgl_b200/graph_abstractions_b200.cuh): the C++ program
tests/shim/shim_algorithms.cu implements BFS / SSSP / CC / PageRank as scatter / gather / compute / reduce /
generate_new_frontier calls with device lambdas — the way the reference's algorithms/* use VGL_GRAPH_ABSTRACTIONS —
and its results must match the oracle: bit-exact levels, distances and labels, PageRank within 1e-6 relative L1."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "shim", "shim_algorithms")


@pytest.mark.parametrize("kind,scale,ef", [(0, 10, 8), (1, 14, 16), (2, 12, 32)])
def test_lambda_api_algorithms_match_oracle(vgl, oracle, tmp_path, kind, scale, ef):
    O = oracle
    if not os.path.exists(BIN):
        from vectorgraphlibrary_b200 import build
        build.build_shim_test()
    V = 1 << scale
    seed, wseed, iters = 0xB200 + kind, 77, 20
    src, dst = O.generate_edges(kind, scale, ef, seed)
    og = O.OracleGraph(V, src, dst)
    source = O.pick_sources(V, np.bincount(src, minlength=V), 1, seed)[0]
    p = subprocess.run([BIN, str(kind), str(scale), str(ef), str(seed), str(source), str(wseed), str(iters), str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "SHIM_OK" in p.stdout and "errors_caught=2" in p.stdout, p.stdout
    deg = np.diff(og.row_ptr)
    assert f"max_degree={int(deg.max())} " in p.stdout and f"degree_sum={len(src)} " in p.stdout, p.stdout
    levels = np.fromfile(tmp_path / "bfs_levels.bin", np.int32)
    assert np.array_equal(levels, og.bfs(source)[0])
    dist = np.fromfile(tmp_path / "sssp_dist.bin", np.float32)
    assert np.array_equal(dist.view(np.uint32), og.sssp(source, wseed)[0].view(np.uint32))
    labels = np.fromfile(tmp_path / "cc_labels.bin", np.int32)
    assert np.array_equal(labels, og.cc()[0])
    ranks = np.fromfile(tmp_path / "pr_ranks.bin", np.float32)
    assert O.rel_l1(ranks, og.pagerank_f64(iters)) <= 1e-6
