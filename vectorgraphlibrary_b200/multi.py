"""1D-partitioned drivers (one process per GPU) for bench.py and the multi-rank tests, plus the host-side mirror of
the partition arithmetic of csrc/partition.cu.

Partition (see include/vgl_b200.h, "multi-GPU"): the reference's degree-sorted ids s = 0..V-1 are dealt round-robin,
owner(s) = s mod P, local row = s div P; replicated vertex arrays are indexed by column id
column(s) = (s mod P) * vp + s div P with vp = rows per rank rounded up to a multiple of 32.
Reference: per-rank vertex ranges of the MPI build, vect_csr/mpi_api.hpp:6-26, get_api.hpp:66-94.
"""
from __future__ import annotations

import time

import numpy as np

from .dist import _stats, pick_sources


def rows_per_rank(V: int, P: int) -> int:
    """Slice stride vp: ceil(V / P) rounded up to whole 32-bit bitmap words."""
    return ((-(-V // P) + 31) // 32) * 32


def local_rows(V: int, P: int, rank: int) -> int:
    """Number of sorted ids rank, rank + P, ... below V."""
    return (V - rank + P - 1) // P if V > rank else 0


def column_of_sorted(s, P: int, vp: int):
    s = np.asarray(s)
    return (s % P) * vp + s // P


def sorted_of_column(c, P: int, vp: int):
    c = np.asarray(c)
    return (c % vp) * P + c // vp


def owner_of_column(c, vp: int):
    return np.asarray(c) // vp


class PartitionedRunner:
    """N > 1: rank r owns the rows {s : s mod N == r} of a graph N times the single-GPU workload (weak scaling: scale +
    log2 N, same edge factor), built in place from the counter-based generator (no edge list ever leaves a GPU)."""

    ncu_traffic = None
    bfs_direction_optimising = True

    def __init__(self, vgl, ctx, comm, workload, kind, scale, ef, pr_iters, weak=True, vcomm=None):
        self.vgl, self.ctx, self.workload, self.pr_iters = vgl, ctx, workload, pr_iters
        self.tcomm = comm  # torch.distributed plumbing (unique id exchange, host barriers)
        world, rank = comm.world, comm.rank
        extra = int(np.log2(world))
        if (1 << extra) != world:
            raise SystemExit("bench.py: --gpus must be a power of two")
        # weak scaling: the graph grows with the GPU count (scale + log2 N, per-GPU work fixed);
        # strong scaling: the SAME graph at every N (BASELINE config 4 "at 1/2/4/8", the reference's strong_scalability.sh)
        self.weak = weak
        self.scale = scale + extra if weak else scale
        self._own_comm = vcomm is None
        self.comm = vcomm or vgl.Comm(ctx, rank, world, exchange=lambda b: comm.broadcast_bytes(b, vgl.UNIQUE_ID_BYTES, 0))
        flags = vgl.GRAPH_WITH_INCOMING if workload == "bfs" else 0
        self.g = vgl.Graph.from_generator_partitioned(ctx, self.comm, kind, self.scale, ef, flags, symmetrize=(workload == "cc"))
        g = self.g
        self.exchange = "n/a"
        if workload == "pr":
            self.exchange = self._choose_exchange(g)
        self.V_total, self.E_total = g.V_global, g.E_global
        self.adj_bytes_per_gpu = 4 * g.E
        self.partition = (f"1D vertex partition over {world} GPUs: sorted ids dealt round-robin (owner = id mod {world}), "
                          f"{'weak' if weak else 'strong'} scaling (scale {self.scale}), rank 0 holds {g.V} rows / {g.E} edges"
                          + (f"; PageRank exchange: {self.exchange}" if workload == "pr" else ""))
        self.dtype = "f32" if workload in ("pr", "sssp") else "int32"
        self.iters_per_step = pr_iters if workload == "pr" else 1
        self.edges_per_step = g.E_global * self.iters_per_step
        self.out = ctx.empty(max(1, g.V), np.float32 if self.dtype == "f32" else np.int32)
        self.dominant_kernel = {"pr": "pr_sweep_kernel", "bfs": "bfs_td_kernel + bfs_bu_kernel (whole run)",
                                "sssp": "sssp_relax_flat_kernel + sssp_select_kernel (whole run)", "cc": "cc_hook_kernel + cc_jump_kernel (whole run)"}[workload]
        self.weights = g.synthetic_weights(vgl.MASTER_SEED ^ 0x5555) if workload == "sssp" else None
        self.sources = None
        if workload in ("bfs", "sssp"):
            self.sources = self._pick_sources(16)
        self._host = None

    def _pick_sources(self, count):
        """`count` seeded ORIGINAL ids with out-degree > 0, as column ids; identical on every rank (vglb_source_candidate,
        include/vglb_synth.h). The out-degree of a candidate is looked up in the allgathered row lengths on the device —
        a few 4-byte reads instead of a V-sized host array."""
        from .dist import mix64, _M64
        vgl, g, ctx = self.vgl, self.g, self.ctx
        ptr, _unused = g._d2h(g.info.d_out_ptr, g.V + 1, np.int64), None
        host = np.zeros(g.cols, np.int32)
        host[g.col0:g.col0 + g.V] = np.diff(ptr)
        full = ctx.from_numpy(host)
        del host, ptr
        vgl._check(vgl.lib().vglb_comm_allgather(self.comm.h, full.ptr, g.vp * 4))
        ctx.synchronize()
        out, k, seed = [], 0, vgl.MASTER_SEED
        while len(out) < count:
            v = mix64(seed ^ ((0xA5A5A5A5 + k * 0x9E3779B97F4A7C15) & _M64)) % g.V_global
            k += 1
            col = int(g._d2h(g.info.d_orig_to_sorted + 4 * v, 1, np.int32)[0])
            if int(g._d2h(full.ptr + 4 * col, 1, np.int32)[0]) > 0:
                out.append(col)
            if k > 64 * count + 1024:
                raise RuntimeError("no vertex with out-degree > 0")
        full.free()
        return out

    def _choose_exchange(self, g):
        """Peer stores fused into the sweep (CUDA IPC over NVLink) unless VGLB_PR_EXCHANGE=nccl; every rank must end up
        in the same mode, so a rank whose IPC mapping fails takes everybody back to ncclAllGather (and says so)."""
        import os
        import sys
        vgl = self.vgl
        if os.environ.get("VGLB_PR_EXCHANGE", "p2p") == "nccl":
            return "ncclAllGather per sweep"
        ok = 1
        try:
            g.set_exchange(vgl.EXCHANGE_P2P)
        except vgl.VglbError as ex:
            ok = 0
            sys.stderr.write(f"[rank {self.tcomm.rank}] peer-store exchange unavailable: {ex}\n")
        if self.tcomm.sum_int(ok) == self.tcomm.world:
            return "peer stores from the sweep epilogue (CUDA IPC over NVLink) + allreduce of the dangling mass"
        g.set_exchange(vgl.EXCHANGE_NCCL)
        return "ncclAllGather per sweep (peer-store mapping failed)"

    def step(self, i):
        w = self.workload
        if w == "pr":
            _, st = self.g.pagerank(self.pr_iters, 0.85, self.out)
            return _stats(st, self.pr_iters, st.algorithmic_bytes)
        if w == "bfs":
            _, st = self.g.bfs(self.sources[i % len(self.sources)], self.bfs_direction_optimising, self.out)
        elif w == "sssp":
            _, st = self.g.sssp(self.weights, self.sources[i % len(self.sources)], self.out)
        else:
            _, st = self.g.cc(self.out)
        return _stats(st, 1, st.algorithmic_bytes)

    # ---- end to end through the C ABI with HOST buffers: every rank uploads its own part ----
    def _host_arrays(self):
        if self._host is None:
            vgl, g = self.vgl, self.g
            ptr, adj = g.layout()
            H = {"ptr": vgl.pinned_array(g.V + 1, np.int64), "adj": vgl.pinned_array(g.E, np.int32),
                 "fwd": vgl.pinned_array(g.V_global, np.int32), "in_ptr": None, "in_adj": None, "w": None,
                 "out": vgl.pinned_array(max(1, g.V), np.float32 if self.dtype == "f32" else np.int32)}
            H["ptr"][:], H["adj"][:], H["fwd"][:] = ptr, adj, g.orig_to_sorted()
            del ptr, adj
            if self.workload == "bfs":
                iptr, iadj = g.layout(incoming=True)
                H["in_ptr"], H["in_adj"] = vgl.pinned_array(g.V + 1, np.int64), vgl.pinned_array(len(iadj), np.int32)
                H["in_ptr"][:], H["in_adj"][:] = iptr, iadj
            if self.workload == "sssp":
                H["w"] = vgl.pinned_array(g.E, np.float32)
                H["w"][:] = self.weights.to_numpy()
            self._host = H
        return self._host

    def e2e(self, steps):
        vgl, ctx = self.vgl, self.ctx
        H = self._host_arrays()
        L = vgl.lib()
        h2d = sum(H[k].nbytes for k in ("ptr", "adj", "fwd", "in_ptr", "in_adj", "w") if H[k] is not None)
        d2h = H["out"].nbytes
        times = []
        for i in range(steps + 1):
            self.comm.barrier()
            t0 = time.perf_counter()
            g = vgl.Graph.from_csr_partitioned(ctx, self.comm, self.g.V_global, H["ptr"], H["adj"], H["fwd"], H["in_ptr"], H["in_adj"])
            out = ctx.empty(max(1, g.V), H["out"].dtype)
            if self.workload == "pr":
                g.pagerank(self.pr_iters, 0.85, out)
            elif self.workload == "bfs":
                g.bfs(self.sources[i % len(self.sources)], self.bfs_direction_optimising, out)
            elif self.workload == "sssp":
                w = ctx.empty(g.E, np.float32)
                vgl._check(L.vglb_memcpy_h2d(ctx.h, w.ptr, H["w"].ctypes.data, H["w"].nbytes))
                g.sssp(w, self.sources[i % len(self.sources)], out)
                w.free()
            else:
                g.cc(out)
            vgl._check(L.vglb_memcpy_d2h(ctx.h, H["out"].ctypes.data, out.ptr, d2h))
            ctx.synchronize()
            self.comm.barrier()
            dt = time.perf_counter() - t0
            out.free()
            g.free()
            if i > 0:
                times.append(dt)
        # bytes of the whole job: every rank moves its own part
        return {"seconds": float(np.mean(times)), "h2d": int(self.tcomm.sum_int(h2d)), "d2h": int(self.tcomm.sum_int(d2h))}

    def extras(self):
        return {"rows_rank0": self.g.V, "edges_rank0": self.g.E, "columns": self.g.cols}

    def close(self):
        if self.weights is not None:
            self.weights.free()
        self.out.free()
        self.g.free()  # a collective: mappings closed on every rank before anybody frees what it exported
        if self._host:
            for a in self._host.values():
                self.vgl.pinned_free(a)
        self._host = None
        if self._own_comm:
            self.comm.close()
