"""dev (torchrun): one traced partitioned SSSP run; rank 0's per-round lines go to stderr."""
import os, sys, numpy as np
sys.path.insert(0, ".")
import torch
import vectorgraphlibrary_b200 as vgl
from vectorgraphlibrary_b200 import dist as vdist, multi
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
tcomm = vdist.Communicator.from_env(local)
ctx = vgl.Context(local)
r = multi.PartitionedRunner(vgl, ctx, tcomm, "bfs", 1, 26, 16, 20)
for i in range(3):
    st = r.step(i)
if rank == 0:
    os.environ["VGLB_BFS_TRACE"] = "1"
st = r.step(0)
if rank == 0:
    print("ms", st["seconds"] * 1e3, "rounds", st["iterations"], "launches", st["kernel_launches"], flush=True)
r.close(); tcomm.close()
