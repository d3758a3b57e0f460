"""Host-side pieces of the reference module surface that carry no heavy arithmetic:

    Normalizer            models/normalizer.py:5-29
    EmaModel              models/ddim_blocks.py:38-59
    NoiseEstimationLoss   models/losses.py:39-59   (forward value; the fused fwd+grad kernel is K6)
    MaskedLoss            models/losses.py:62-78   (evaluation metric on the sampled fields)
    CorrelationLoss       models/losses.py:96-128  (evaluation metric of the single-task modules)

They keep the reference's class names, constructor arguments, buffer names and return values so
checkpoints and callers are interchangeable.  Everything here is a handful of elementwise torch ops on
tiny tensors (2 channels); the hot path never goes through them except MaskedLoss at the very end of
test_step.
"""
from __future__ import annotations

import copy

import torch
from torch import nn


class Normalizer(nn.Module):
    """(x - subtract) / divide with the statistics held as buffers (`subtract`, `divide`)."""

    def __init__(self, subtract=None, divide=None, stats_shape=()):
        super().__init__()
        subtract = torch.zeros(stats_shape) if subtract is None else torch.as_tensor(subtract)
        divide = torch.ones(stats_shape) if divide is None else torch.as_tensor(divide)
        self.register_buffer("subtract", subtract)
        self.register_buffer("divide", divide)

    def set_stats(self, subtract, divide):
        self.subtract = torch.as_tensor(subtract).to(self.subtract.device)
        self.divide = torch.as_tensor(divide).to(self.divide.device)

    def forward(self, x, inverse=False):
        if inverse:
            return x * self.divide + self.subtract
        return (x - self.subtract) / self.divide


class EmaModel(nn.Module):
    """Exponential moving average of a model's trainable parameters; `ma_model` is a deep copy."""

    def __init__(self, model, beta):
        super().__init__()
        self.beta = beta
        self.ma_model = copy.deepcopy(model)

    @torch.no_grad()
    def update(self, current_model):
        if isinstance(current_model, nn.parallel.DistributedDataParallel):
            current_model = current_model.module
        cur = [p for p in current_model.parameters()]
        ma = [p for p in self.ma_model.parameters()]
        sel = [(c, m) for c, m in zip(cur, ma) if c.requires_grad]
        if not sel:
            return
        if len(sel) == len(cur) and cur[0].is_cuda:
            # one kernel over flat buffers (train_small.cu ema_update); no torch arithmetic on the GPU path
            from .optim import ema_update_

            if not hasattr(self, "_flat_state"):
                object.__setattr__(self, "_flat_state", {})
            ema_update_(ma, cur, self.beta, self._flat_state)
            return
        # ema = ema * beta + (1 - beta) * p   (ddim_blocks.py:53-56), one fused multi-tensor pass
        mas = [m for _, m in sel]
        torch._foreach_mul_(mas, self.beta)
        torch._foreach_add_(mas, [c.detach() for c, _ in sel], alpha=1 - self.beta)

    def update_average(self, old, new):
        if old is None:
            return new
        return old * self.beta + (1 - self.beta) * new

    def forward(self, *args, **kwargs):
        return self.ma_model(*args, **kwargs)


class NoiseEstimationLoss(nn.Module):
    """sum over (c,h,w) of weight * (pred - target)^2, then mean / sum / none over the batch."""

    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, pred, target, weight=1.0):
        per_sample = torch.sum(weight * (pred - target) ** 2, dim=(1, 2, 3))
        if self.reduction == "mean":
            return torch.mean(per_sample)
        if self.reduction == "sum":
            return torch.sum(per_sample)
        return per_sample


class MaskedLoss(nn.Module):
    """Mean absolute (or squared) error over the entries where mask == 1."""

    def __init__(self, loss="l1"):
        super().__init__()
        self.l1 = loss == "l1"

    def forward(self, pred, target, mask, loss_dim=None):
        pred = pred * mask
        target = target * mask
        if loss_dim is not None:
            pred, target, mask = pred[..., loss_dim], target[..., loss_dim], mask[..., loss_dim]
        diff = pred - target
        total = diff.abs().sum() if self.l1 else (diff * diff).sum()
        return total / torch.sum(mask)


class CorrelationLoss(nn.Module):
    """Pearson correlation per channel between prediction and target, averaged over the batch."""

    def __init__(self, reduction="none"):
        super().__init__()
        self.reduction = reduction

    @staticmethod
    def calculate_correlation(x, y):
        x_bar = x - torch.mean(x, dim=1, keepdim=True)
        y_bar = y - torch.mean(y, dim=1, keepdim=True)
        cov = torch.sum(y_bar * x_bar, dim=1)
        denominator = torch.sqrt(torch.sum(x_bar * x_bar, dim=1) * torch.sum(y_bar * y_bar, dim=1))
        denominator = torch.where(denominator == 0, denominator + 1e-7, denominator)
        return torch.mean(cov / denominator, dim=0)

    def forward(self, pred, target):
        pred = pred.reshape(pred.shape[0], -1, pred.shape[-1])
        target = target.reshape(target.shape[0], -1, target.shape[-1])
        corr = self.calculate_correlation(pred, target)
        if self.reduction == "mean":
            return torch.mean(corr)
        if self.reduction == "sum":
            return torch.sum(corr)
        return corr
