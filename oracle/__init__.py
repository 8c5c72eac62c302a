"""oracle — TEST INFRASTRUCTURE ONLY (never imported by vectorgraphlibrary_b200).

ctypes front-ends for
  * ``liboracle.so``            — the plain-C restatement of the reference path (oracle/vgl_oracle.c), and
  * ``_ref/libvgl_ref_*.so``    — the unmodified reference (multicore/OpenMP build) wrapped by oracle/ref_harness.cpp,
                                  compiled in the build container from /root/reference (oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GEN_RMAT, GEN_KRONECKER, GEN_UNIFORM = 0, 1, 2
MASTER_SEED = 0xB200

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checkers (make -C oracle). Building the checker is not using it."""
    targets = ["oracle"] + (["ref", "refgpu"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE, "-j4"] + targets, check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.vglo_generate_edges.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.c_int, _i32p, _i32p]
        L.vglo_build_vect_csr.argtypes = [C.c_int32, C.c_int64, _i32p, _i32p, _i64p, _i32p, _i32p, _i32p, C.c_void_p]
        L.vglo_build_vect_csr.restype = C.c_int
        L.vglo_estimate_thresholds.argtypes = [C.c_int32, _i64p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.vglo_edge_weights.argtypes = [C.c_int32, _i64p, _i32p, _i32p, C.c_uint64, _f32p]
        L.vglo_bfs.argtypes = [C.c_int32, _i64p, _i32p, C.c_int32, _i32p, C.POINTER(C.c_int64)]
        L.vglo_indegree_noloops.argtypes = [C.c_int32, _i64p, _i32p, _i32p]
        L.vglo_pagerank_f32.argtypes = [C.c_int32, _i64p, _i32p, _i32p, C.c_int, C.c_int, _f32p]
        L.vglo_pagerank_f32_tree_rows.argtypes = [C.c_int32, _i64p, _i32p, _i32p, C.c_int, C.c_int, _f32p]
        L.vglo_pagerank_f64.argtypes = [C.c_int32, _i64p, _i32p, _i32p, C.c_int, _f64p]
        L.vglo_sssp.argtypes = [C.c_int32, _i64p, _i32p, _f32p, C.c_int32, _f32p, C.POINTER(C.c_int64)]
        L.vglo_sssp_frontier_bf.argtypes = [C.c_int32, _i64p, _i32p, _f32p, C.c_int32, _f32p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.vglo_cc.argtypes = [C.c_int32, _i64p, _i32p, _i32p, C.POINTER(C.c_int32)]
        L.vglo_rel_l1_f32.argtypes = [C.c_int64, _f32p, _f32p]
        L.vglo_rel_l1_f32.restype = C.c_double
        _LIB = L
    return _LIB


# ---------------------------------------------------------------------------------------------------------------------
# C oracle front-end
# ---------------------------------------------------------------------------------------------------------------------

def generate_edges(kind: int, scale: int, edge_factor: int, seed: int = MASTER_SEED, abc=(57, 19, 19)):
    E = edge_factor << scale
    src = np.empty(E, np.int32)
    dst = np.empty(E, np.int32)
    lib().vglo_generate_edges(kind, scale, E, seed, abc[0], abc[1], abc[2], src, dst)
    return src, dst


def symmetrize(src, dst):
    """CC input: append the reversed copy of every edge."""
    return np.concatenate([src, dst]), np.concatenate([dst, src])


class OracleGraph:
    """Degree-sorted CSR of one direction in VGL numbering (vect_csr/import.hpp:257-337)."""

    def __init__(self, V: int, src: np.ndarray, dst: np.ndarray, want_edge_order: bool = False):
        E = int(src.shape[0])
        self.V, self.E = V, E
        self.row_ptr = np.empty(V + 1, np.int64)
        self.adj = np.empty(E, np.int32)
        self.fwd = np.empty(V, np.int32)
        self.bwd = np.empty(V, np.int32)
        self.edge_order = np.empty(E, np.int64) if want_edge_order else None
        eo = self.edge_order.ctypes.data if want_edge_order else None
        rc = lib().vglo_build_vect_csr(V, E, np.ascontiguousarray(src), np.ascontiguousarray(dst),
                                       self.row_ptr, self.adj, self.fwd, self.bwd, eo)
        if rc != 0:
            raise MemoryError("vglo_build_vect_csr")

    @classmethod
    def from_csr(cls, row_ptr: np.ndarray, adj: np.ndarray, fwd: np.ndarray | None = None):
        """Adopt an already-built degree-sorted CSR (e.g. the one the GPU builder produced, downloaded): lets the
        O(E) oracle algorithms run at sizes where the oracle's own import would take minutes."""
        self = cls.__new__(cls)
        self.V, self.E = int(row_ptr.shape[0]) - 1, int(adj.shape[0])
        self.row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        self.adj = np.ascontiguousarray(adj, np.int32)
        self.fwd = np.arange(self.V, dtype=np.int32) if fwd is None else np.ascontiguousarray(fwd, np.int32)
        self.bwd = np.empty(self.V, np.int32)
        self.bwd[self.fwd] = np.arange(self.V, dtype=np.int32)
        self.edge_order = None
        return self

    def thresholds(self, ve_value: int, vc_value: int):
        ve, vc = C.c_int32(), C.c_int32()
        lib().vglo_estimate_thresholds(self.V, self.row_ptr, ve_value, vc_value, C.byref(ve), C.byref(vc))
        return ve.value, vc.value

    def to_original(self, arr_sorted: np.ndarray) -> np.ndarray:
        return arr_sorted[self.fwd]

    def weights(self, seed: int) -> np.ndarray:
        w = np.empty(self.E, np.float32)
        lib().vglo_edge_weights(self.V, self.row_ptr, self.adj, self.bwd, seed, w)
        return w

    def bfs(self, source_orig: int):
        levels = np.empty(self.V, np.int32)
        insp = C.c_int64()
        lib().vglo_bfs(self.V, self.row_ptr, self.adj, int(self.fwd[source_orig]), levels, C.byref(insp))
        return self.to_original(levels), insp.value

    def indegree_noloops(self) -> np.ndarray:
        d = np.empty(self.V, np.int32)
        lib().vglo_indegree_noloops(self.V, self.row_ptr, self.adj, d)
        return d

    def pagerank_f32(self, iters: int, threads: int):
        r = np.empty(self.V, np.float32)
        lib().vglo_pagerank_f32(self.V, self.row_ptr, self.adj, self.indegree_noloops(), iters, threads, r)
        return self.to_original(r)

    def pagerank_f32_tree_rows(self, iters: int, threads: int):
        """fp32 state, reference-order dangling sum over `threads` chunks, fp64-accumulated row sums (see vgl_oracle.c)."""
        r = np.empty(self.V, np.float32)
        lib().vglo_pagerank_f32_tree_rows(self.V, self.row_ptr, self.adj, self.indegree_noloops(), iters, threads, r)
        return self.to_original(r)

    def pagerank_f64(self, iters: int):
        r = np.empty(self.V, np.float64)
        lib().vglo_pagerank_f64(self.V, self.row_ptr, self.adj, self.indegree_noloops(), iters, r)
        return self.to_original(r)

    def sssp(self, source_orig: int, weight_seed: int):
        dist = np.empty(self.V, np.float32)
        rel = C.c_int64()
        lib().vglo_sssp(self.V, self.row_ptr, self.adj, self.weights(weight_seed), int(self.fwd[source_orig]), dist, C.byref(rel))
        return self.to_original(dist), rel.value

    def sssp_frontier_bf(self, source_orig: int, weight_seed: int):
        dist = np.empty(self.V, np.float32)
        rel, its = C.c_int64(), C.c_int32()
        lib().vglo_sssp_frontier_bf(self.V, self.row_ptr, self.adj, self.weights(weight_seed), int(self.fwd[source_orig]),
                                    dist, C.byref(rel), C.byref(its))
        return self.to_original(dist), rel.value, its.value

    def cc(self):
        comp = np.empty(self.V, np.int32)
        rounds = C.c_int32()
        lib().vglo_cc(self.V, self.row_ptr, self.adj, comp, C.byref(rounds))
        return self.to_original(comp), rounds.value


def rel_l1(a: np.ndarray, ref: np.ndarray) -> float:
    a64, r64 = a.astype(np.float64), ref.astype(np.float64)
    return float(np.abs(a64 - r64).sum() / np.abs(r64).sum())


def pick_sources(V: int, out_degree_orig: np.ndarray, count: int, seed: int = MASTER_SEED):
    """`count` seeded ORIGINAL ids with out-degree > 0 (vglb_source_candidate twin; apps/bfs/bfs.cpp:38)."""
    M = (1 << 64) - 1

    def mix64(z):
        z = (z + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    out, k = [], 0
    while len(out) < count:
        v = mix64(seed ^ ((0xA5A5A5A5 + k * 0x9E3779B97F4A7C15) & M)) % V
        k += 1
        if out_degree_orig[v] > 0:
            out.append(int(v))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Reference (unmodified VGL multicore build) front-end
# ---------------------------------------------------------------------------------------------------------------------

def ref_available(profile: str = "pr", timing: bool = False) -> bool:
    """oracle/_ref holds the unmodified reference compiled for `profile`; timing=True asks for the optimised build
    (oracle/Makefile TIMING_OPT) that bench.py's CPU arm times, False for the strict-IEEE parity build."""
    return os.path.exists(os.path.join(_HERE, "_ref", f"libvgl_ref_{profile}{'_timing' if timing else ''}.so"))


_REF_LIBS = {}


def ref_lib(profile: str, timing: bool = False) -> C.CDLL:
    key = profile
    if timing:
        profile = profile + "_timing"
    if profile not in _REF_LIBS:
        if int(os.environ.get("OMP_NUM_THREADS", "2")) < 2:
            raise RuntimeError("the reference segfaults with OMP_NUM_THREADS=1 (SURVEY App. A.1)")
        L = C.CDLL(os.path.join(_HERE, "_ref", f"libvgl_ref_{profile}.so"))
        L.vglref_max_threads.restype = C.c_int
        L.vglref_graph_create.argtypes = [C.c_int, C.c_longlong, _i32p, _i32p]
        L.vglref_graph_create.restype = C.c_void_p
        L.vglref_graph_destroy.argtypes = [C.c_void_p]
        L.vglref_graph_layout.argtypes = [C.c_void_p, C.c_int, _i64p, _i32p, _i32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.vglref_bfs.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_int]
        L.vglref_bfs.restype = C.c_double
        L.vglref_pagerank.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.vglref_pagerank.restype = C.c_double
        L.vglref_sssp.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, _f32p, C.c_int]
        L.vglref_sssp.restype = C.c_double
        L.vglref_cc.argtypes = [C.c_void_p, _i32p]
        L.vglref_cc.restype = C.c_double
        if hasattr(L, "vglref_graph_save"):  # file-format entry points (older prebuilt harnesses lack them)
            L.vglref_edges_save.argtypes = [C.c_char_p, C.c_int, C.c_longlong, _i32p, _i32p]
            L.vglref_graph_create_from_edges_file.argtypes = [C.c_char_p]
            L.vglref_graph_create_from_edges_file.restype = C.c_void_p
            L.vglref_graph_save.argtypes = [C.c_void_p, C.c_char_p]
            L.vglref_graph_load.argtypes = [C.c_char_p]
            L.vglref_graph_load.restype = C.c_void_p
            L.vglref_graph_vertices.argtypes = [C.c_void_p]
            L.vglref_graph_edges.argtypes = [C.c_void_p]
            L.vglref_graph_edges.restype = C.c_longlong
        _REF_LIBS[profile] = L
    return _REF_LIBS[profile]


def ref_save_edges(path: str, V: int, src: np.ndarray, dst: np.ndarray, profile: str = "bfs"):
    """EdgesContainer::save_to_binary_file of the unmodified reference."""
    rc = ref_lib(profile).vglref_edges_save(os.fsencode(path), V, len(src), np.ascontiguousarray(src, np.int32),
                                            np.ascontiguousarray(dst, np.int32))
    if rc != 0:
        raise RuntimeError("EdgesContainer::save_to_binary_file failed")


class RefGraph:
    """The reference's own VGL_Graph(VECTOR_CSR_GRAPH) built from an edge list (vgl_graph.hpp:57-68)."""

    @classmethod
    def _adopt(cls, L, handle):
        if not handle:
            raise RuntimeError("the reference could not read the file")
        self = cls.__new__(cls)
        self.L, self.h = L, handle
        self.V, self.E = L.vglref_graph_vertices(handle), L.vglref_graph_edges(handle)
        return self

    @classmethod
    def from_edges_file(cls, path: str, profile: str = "bfs"):
        """EdgesContainer::load_from_binary_file + VGL_Graph::import — the apps' `-import <file>` path."""
        L = ref_lib(profile)
        return cls._adopt(L, L.vglref_graph_create_from_edges_file(os.fsencode(path)))

    @classmethod
    def load(cls, path: str, profile: str = "bfs"):
        """VGL_Graph::load_from_binary_file."""
        L = ref_lib(profile)
        return cls._adopt(L, L.vglref_graph_load(os.fsencode(path)))

    def save(self, path: str):
        """VGL_Graph::save_to_binary_file."""
        if self.L.vglref_graph_save(self.h, os.fsencode(path)) != 0:
            raise RuntimeError("VGL_Graph::save_to_binary_file failed")

    def __init__(self, V: int, src: np.ndarray, dst: np.ndarray, profile: str = "pr", timing: bool = False):
        self.L = ref_lib(profile, timing)
        self.V, self.E = V, int(src.shape[0])
        if self.E >= 2 ** 31:
            raise ValueError("the reference cannot process E >= 2^31 (SURVEY App. A.5)")
        self.h = self.L.vglref_graph_create(V, self.E, np.ascontiguousarray(src), np.ascontiguousarray(dst))
        if not self.h:
            raise RuntimeError("vglref_graph_create failed")

    def threads(self) -> int:
        return self.L.vglref_max_threads()

    def close(self):
        if self.h:
            self.L.vglref_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def layout(self, direction: int = 0):
        ptr = np.empty(self.V + 1, np.int64)
        adj = np.empty(self.E, np.int32)
        fwd = np.empty(self.V, np.int32)
        ve, vc = C.c_int(), C.c_int()
        self.L.vglref_graph_layout(self.h, direction, ptr, adj, fwd, C.byref(ve), C.byref(vc))
        return ptr, adj, fwd, ve.value, vc.value

    def bfs(self, source_orig: int, mode: int = 0):
        lv = np.empty(self.V, np.int32)
        t = self.L.vglref_bfs(self.h, source_orig, lv, mode)
        return lv, t

    def pagerank(self, iters: int):
        r = np.empty(self.V, np.float32)
        t = self.L.vglref_pagerank(self.h, iters, r)
        return r, t

    def sssp(self, source_orig: int, weight_seed: int, mode: int = 0):
        d = np.empty(self.V, np.float32)
        t = self.L.vglref_sssp(self.h, weight_seed, source_orig, d, mode)
        return d, t

    def cc(self):
        c = np.empty(self.V, np.int32)
        t = self.L.vglref_cc(self.h, c)
        return c, t


# ---------------------------------------------------------------------------------------------------------------------
# Reference GPU build (-D __USE_GPU__) front-end: the reference's own algorithms on (a) this repo's B200 backend through the
# include-path overlay ("dropin") or (b) the reference's own CUDA backend recompiled for sm_100a ("refgpu")
# ---------------------------------------------------------------------------------------------------------------------

def gpu_ref_available(which: str = "dropin") -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libvgl_{which}.so"))


_GPU_LIBS = {}


def gpu_ref_lib(which: str) -> C.CDLL:
    if which not in _GPU_LIBS:
        L = C.CDLL(os.path.join(_HERE, "_ref", f"libvgl_{which}.so"))
        L.vglgpu_graph_create.argtypes = [C.c_int, C.c_longlong, _i32p, _i32p]
        L.vglgpu_graph_create.restype = C.c_void_p
        L.vglgpu_graph_destroy.argtypes = [C.c_void_p]
        L.vglgpu_bfs.argtypes = [C.c_void_p, C.c_int, _i32p]
        L.vglgpu_bfs.restype = C.c_double
        L.vglgpu_pagerank.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p]
        L.vglgpu_pagerank.restype = C.c_double
        L.vglgpu_sssp.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, _f32p, C.c_int]
        L.vglgpu_sssp.restype = C.c_double
        L.vglgpu_cc.argtypes = [C.c_void_p, _i32p]
        L.vglgpu_cc.restype = C.c_double
        L.vglgpu_hits.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, C.c_void_p, C.c_void_p]
        L.vglgpu_hits.restype = C.c_double
        L.vglgpu_group_mark.argtypes = [C.c_void_p, _i32p, C.c_int, _i32p]
        L.vglgpu_group_mark.restype = C.c_longlong
        _GPU_LIBS[which] = L
    return _GPU_LIBS[which]


class GpuRefGraph:
    """VGL_Graph of the reference's GPU build; `which` = "dropin" (B200 backend through the overlay) or "refgpu"."""

    def __init__(self, V: int, src: np.ndarray, dst: np.ndarray, which: str = "dropin"):
        self.L = gpu_ref_lib(which)
        self.which, self.V, self.E = which, V, int(src.shape[0])
        assert self.L.vglgpu_is_b200_backend() == (1 if which == "dropin" else 0)
        self.h = self.L.vglgpu_graph_create(V, self.E, np.ascontiguousarray(src, np.int32), np.ascontiguousarray(dst, np.int32))
        if not self.h:
            raise RuntimeError("vglgpu_graph_create failed")

    def close(self):
        if self.h:
            self.L.vglgpu_graph_destroy(self.h)
            self.h = None

    def _checked(self, t):
        if t < 0:
            raise RuntimeError(f"the reference ({self.which}) threw; see stderr")
        return t

    def bfs(self, source_orig: int):
        lv = np.empty(self.V, np.int32)
        return lv, self._checked(self.L.vglgpu_bfs(self.h, source_orig, lv))

    def pagerank(self, iters: int, push: bool = False):
        r = np.empty(self.V, np.float32)
        return r, self._checked(self.L.vglgpu_pagerank(self.h, iters, 1 if push else 0, r))

    def sssp(self, source_orig: int, weight_seed: int, mode: int = 2):
        d = np.empty(self.V, np.float32)
        return d, self._checked(self.L.vglgpu_sssp(self.h, weight_seed, source_orig, d, mode))

    def cc(self):
        c = np.empty(self.V, np.int32)
        return c, self._checked(self.L.vglgpu_cc(self.h, c))

    def hits(self, steps: int, with_seq: bool = True):
        a, hb = np.empty(self.V, np.float32), np.empty(self.V, np.float32)
        sa, sh = np.empty(self.V, np.float32), np.empty(self.V, np.float32)
        t = self._checked(self.L.vglgpu_hits(self.h, steps, a, hb, sa.ctypes.data if with_seq else None, sh.ctypes.data if with_seq else None))
        return a, hb, (sa if with_seq else None), (sh if with_seq else None), t

    def group_mark(self, ids_orig):
        ids = np.ascontiguousarray(ids_orig, np.int32)
        marks = np.empty(self.V, np.int32)
        total = self.L.vglgpu_group_mark(self.h, ids, len(ids), marks)
        if total < 0:
            raise RuntimeError(f"the reference ({self.which}) threw; see stderr")
        return marks, int(total)
