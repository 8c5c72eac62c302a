"""CPU-side checks of the drop-in build (no compute call): the harness compiled against include/vgl_b200/overlay exists when the
reference tree was available at build time, loads, reports the B200 backend and exports the entry points the GPU tests use;
the overlay header itself names the reference interface it replaces."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "libvgl_dropin.so")
OVERLAY = os.path.join(ROOT, "include", "vgl_b200", "overlay", "vgl_compute_api", "gpu", "graph_abstractions_gpu.h")


def test_overlay_header_declares_the_reference_interface():
    text = open(OVERLAY).read()
    for needle in ("class GraphAbstractionsGPU : public GraphAbstractions", "GraphAbstractionsGPU(VGL_Graph &_graph, TraversalDirection",
                   "void scatter(VGL_Graph &_graph, VGL_Frontier &_frontier", "void gather(VGL_Graph &_graph, VGL_Frontier &_frontier",
                   "void compute(VGL_Graph &_graph, VGL_Frontier &_frontier", "_T reduce(VGL_Graph &_graph, VGL_Frontier &_frontier",
                   "void generate_new_frontier(VGL_Graph &_graph, VGL_Frontier &_frontier", "friend class GraphAbstractions;",
                   "void advance_worker(VectorCSRGraph &_graph, FrontierVectorCSR &_frontier", "bool _inner_mpi_processing"):
        assert needle in text, needle
    assert os.path.exists(os.path.join(os.path.dirname(OVERLAY), "vector_register", "vector_registers.h"))


@pytest.mark.skipif(not os.path.exists("/root/reference/graph_library.h") and not os.path.exists(DROPIN),
                    reason="neither the reference tree nor a prebuilt drop-in harness is present")
def test_dropin_harness_is_built_against_the_overlay():
    assert os.path.exists(DROPIN), "run `make -C oracle refgpu` (the reference's algorithms compiled against the overlay)"
    L = C.CDLL(DROPIN)
    assert L.vglgpu_is_b200_backend() == 1
    for sym in ("vglgpu_graph_create", "vglgpu_bfs", "vglgpu_pagerank", "vglgpu_sssp", "vglgpu_cc", "vglgpu_hits", "vglgpu_group_mark"):
        assert hasattr(L, sym), sym
    ref = os.path.join(ROOT, "oracle", "_ref", "libvgl_refgpu.so")
    if os.path.exists(ref):
        assert C.CDLL(ref).vglgpu_is_b200_backend() == 0
