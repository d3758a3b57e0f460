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

# ---- sample_edm_sharded: each rank samples its contiguous row block (micro-batched), one gather returns all rows in order
class FakeModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def sample_edm(self, hu, cond, hu_mask, sparams, return_last=True):
        self.calls.append(hu.shape[0])
        return (cond * 3.0 + hu_mask).double().unsqueeze(1)


fm = FakeModule()
cond = torch.arange(11 * 4, dtype=torch.float32).reshape(11, 4)
msk = torch.ones(11, 4)
full = MD.sample_edm_sharded(fm, torch.zeros(11, 4), cond, msk, None, chunk=4)
assert torch.equal(full, (cond * 3.0 + 1.0).double().unsqueeze(1))
assert fm.calls == ([4, 2] if rank == 0 else [4, 1]), fm.calls      # 6 / 5 rows per rank in micro-batches of 4

# ---- training: one flat all-reduce, mean over ranks
tr = Trainer(max_epochs=1)
ps = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(5))]
for p in ps:
    p.grad = torch.full_like(p, float(rank + 1))
tr._allreduce_grads(ps)
for p in ps:
    assert torch.allclose(p.grad, torch.full_like(p, 1.5))

# ---- FusedAdam.flat_grads() is idempotent within a step: the tensor that was all-reduced is the one step() consumes
# (the eager path used to gather the local .grad tensors a second time over the reduced buffer).  The Adam
# kernel itself is CUDA-only, so the flat parameter buffer is stubbed here and only the gradient contract is exercised.
from mcedm_b200.optim import FusedAdam  # noqa: E402

ps = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(5))]
opt = FusedAdam(ps, lr=1e-3)
opt._flat_p = torch.zeros(17)
for p in ps:
    p.grad = torch.full_like(p, float(rank + 1))
g = opt.flat_grads()
dist.all_reduce(g)
g2 = opt.flat_grads()
assert g2 is g and torch.equal(g2, torch.full((17,), 3.0)), "second flat_grads() re-gathered the local gradients"
opt.zero_grad()
assert opt._pending_g is None
for p in ps:
    p.grad = torch.full_like(p, 7.0)
assert torch.equal(opt.flat_grads(), torch.full((17,), 7.0))      # a new step fetches fresh gradients
dist.barrier()
if rank == 0:
    print("OK")
dist.destroy_process_group()
