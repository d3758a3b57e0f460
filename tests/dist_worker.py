"""Worker for the world_size-2 gloo test of the multi-GPU host logic (no CUDA needed)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import dist as MD  # noqa: E402
from mcedm_b200.runner import Trainer  # noqa: E402

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)

# ---- sampling: contiguous row blocks per rank, one gather at the end
rows = torch.arange(10 * 3, dtype=torch.float64).reshape(10, 3)
lo, hi = MD.shard_rows(10, rank, world)
assert (lo, hi) == ((0, 5) if rank == 0 else (5, 10))
mine = rows[lo:hi] * 2.0
full = MD.gather_rows(mine, 10)
assert torch.equal(full, rows * 2.0)
lo, hi = MD.shard_rows(7, rank, world)          # ragged
full = MD.gather_rows(rows[:7][lo:hi] + 1, 7)
assert torch.equal(full, rows[:7] + 1)
assert MD.rank_seed(1, rank) == 1 + rank

# ---- training: one flat all-reduce, mean over ranks
tr = Trainer(max_epochs=1)
ps = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(5))]
for p in ps:
    p.grad = torch.full_like(p, float(rank + 1))
tr._allreduce_grads(ps)
for p in ps:
    assert torch.allclose(p.grad, torch.full_like(p, 1.5))
dist.barrier()
if rank == 0:
    print("OK")
dist.destroy_process_group()
