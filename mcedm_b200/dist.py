"""Multi-GPU plumbing for the two places the hot path shards.

Sampling: every row of the (n_samples * b) batch is an independent trajectory (GroupNorm and attention
are per sample), so rows are split into contiguous blocks, one block per rank, with NO communication
during the 50-step loop and ONE all-gather of the final fp64 fields at the end (the reference stacks
all rows on one GPU, models/mcedm.py:352-376).  Training: identical replicas, ONE all-reduce over a
flat gradient buffer per step (runner.Trainer._allreduce_grads).  One process per GPU, NCCL backend
on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the process group if needed."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) block of rank `rank`; the first (n_rows % world) ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank generator seed for scaling runs (parity with the reference is defined at 1 GPU)."""
    return seed + rank


def gather_rows(local: torch.Tensor, n_rows: int) -> torch.Tensor:
    """All ranks receive the [n_rows, ...] tensor assembled from the per-rank row blocks (one all-gather)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [shard_rows(n_rows, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in counts)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    if all(hi - lo == max_rows for lo, hi in counts):
        return out
    return torch.cat([out[r * max_rows: r * max_rows + (hi - lo)] for r, (lo, hi) in enumerate(counts)], dim=0)


def sample_rows(module, hu, cond, hu_mask, sparams, return_last=True, chunk=None):
    """`module.sample_edm` over the given rows in micro-batches of `chunk` rows (None: one call).  Rows are independent
    trajectories (bit-identical alone or inside any batch: tests/test_gpu_parity.py::test_full_size_micro_batch_properties),
    so the micro-batching only bounds the activation workspace; cond / mask may live in pinned host memory."""
    n = hu.shape[0]
    if chunk is None or chunk >= n:
        dev = next(module.parameters()).device
        return module.sample_edm(hu, cond.to(dev, non_blocking=True), hu_mask.to(dev, non_blocking=True), sparams,
                                 return_last=return_last)
    dev = next(module.parameters()).device
    outs = []
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        outs.append(module.sample_edm(hu[lo:hi], cond[lo:hi].to(dev, non_blocking=True),
                                      hu_mask[lo:hi].to(dev, non_blocking=True), sparams, return_last=return_last))
    return torch.cat(outs, dim=0)


def sample_edm_sharded(module, hu, cond, hu_mask, sparams, return_last=True, chunk=None):
    """`module.sample_edm` on this rank's row block, then one gather: returns xs for ALL rows on every rank.
    hu / cond / hu_mask hold ALL rows (the reference stacks n_samples * b rows on one GPU, models/mcedm.py:352-376);
    each rank touches only its [lo, hi) block."""
    n = hu.shape[0]
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sample_rows(module, hu, cond, hu_mask, sparams, return_last, chunk)
    lo, hi = shard_rows(n, dist.get_rank(), dist.get_world_size())
    xs = sample_rows(module, hu[lo:hi], cond[lo:hi], hu_mask[lo:hi], sparams, return_last, chunk)
    return gather_rows(xs.contiguous(), n)
