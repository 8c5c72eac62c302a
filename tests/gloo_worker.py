"""One rank of the CPU (gloo) multi-process test of the host-side multi-GPU logic: launched by torchrun from
tests/test_multirank_cpu.py with WORLD_SIZE=2. No compute call is made (libvgl_b200 has no CPU path)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from vectorgraphlibrary_b200 import dist as vdist, multi
    c = vdist.Communicator("gloo")
    rank, world = c.rank, c.world
    assert world == int(os.environ["WORLD_SIZE"]) and rank == int(os.environ["RANK"])
    # the plumbing bench.py / PartitionedRunner rely on
    uid = bytes(range(128)) if rank == 0 else None
    assert c.broadcast_bytes(uid, 128, 0) == bytes(range(128))          # NCCL unique id hand-out
    assert c.sum_int(rank + 1) == world * (world + 1) // 2               # agreement on the exchange mode, byte totals
    assert c.max_float(10.0 + rank) == 10.0 + world - 1                  # max-over-ranks timing
    got = c.allgather_bytes(bytes([rank]) * 4)
    assert got == [bytes([r]) * 4 for r in range(world)]
    c.barrier()
    # the round-robin deal: the ranks' row sets partition the sorted ids, columns are a bijection onto non-padding slots
    for V in (1, 31, 32, 1000, 4096 + 5):
        vp = multi.rows_per_rank(V, world)
        mine = np.arange(rank, V, world)
        assert len(mine) == multi.local_rows(V, world, rank)
        cols = multi.column_of_sorted(mine, world, vp)
        assert np.array_equal(cols, rank * vp + np.arange(len(mine)))
        assert np.all(multi.owner_of_column(cols, vp) == rank)
        assert np.array_equal(multi.sorted_of_column(cols, world, vp), mine)
        total = c.sum_int(len(mine))
        assert total == V
    # seeded sources are identical on every rank
    deg = (np.arange(1000) % 3).astype(np.int64)
    s = vdist.pick_sources(1000, deg, 8, 0xB200)
    blob = np.asarray(s, np.int64).tobytes()
    assert all(b == blob for b in c.allgather_bytes(blob))
    c.close()
    print(f"GLOO_WORKER_OK rank {rank}/{world}", flush=True)


if __name__ == "__main__":
    main()
