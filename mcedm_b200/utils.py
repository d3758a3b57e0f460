"""Small host-side helpers shared by tests, bench and the runner."""
from __future__ import annotations

import hashlib
from typing import Dict

import torch

# the reference zero-initialises these tensors (models/adm_blocks.py:145,157,222,317): a freshly built
# network outputs exactly 0, which makes parity checks and benchmarks on it vacuous
_ZERO_INIT_SUFFIXES = ("conv1.weight", "proj.weight", "out_conv.weight")


def randomize_zero_init(module_or_sd, seed: int = 2, gain: float = 1.0) -> None:
    """Overwrites every zero-initialised weight with N(0, gain^2 / fan_in), deterministically.

    Works on any module or state_dict that uses the reference's parameter names (the reference model,
    the oracle's state_dict or `mcedm_b200.DhariwalUNet`), so all three can be given identical
    "stress" weights.  Draws come from a private CPU generator in key order.
    """
    sd: Dict[str, torch.Tensor] = module_or_sd if isinstance(module_or_sd, dict) else dict(module_or_sd.state_dict())
    gen = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name in sd:
            if name.endswith(_ZERO_INIT_SUFFIXES):
                t = sd[name]
                fan_in = t[0].numel()
                new = torch.randn(t.shape, generator=gen, dtype=torch.float32) * (gain / fan_in ** 0.5)
                t.copy_(new.to(device=t.device, dtype=t.dtype))


def state_hash(sd: Dict[str, torch.Tensor], prefix: str = "") -> str:
    """sha256 over the fp32 bytes of all floating-point entries under `prefix`, in key order."""
    h = hashlib.sha256()
    for k, v in sd.items():
        if not k.startswith(prefix) or not torch.is_floating_point(v):
            continue
        h.update(k[len(prefix):].encode())
        h.update(v.detach().to("cpu", torch.float32).contiguous().numpy().tobytes())
    return h.hexdigest()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b|| / ||b|| in fp64."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
