"""Times training steps of PlMcedm on one GPU (synthetic SWE batch): forward+loss / backward / optimizer+EMA.
usage: python scripts/train_bench.py [B] [steps]"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L  # noqa: E402
from mcedm_b200 import data as D  # noqa: E402
from mcedm_b200.config import compose  # noqa: E402
from mcedm_b200.mcedm import PlMcedm  # noqa: E402
from mcedm_b200.utils import randomize_zero_init  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
cfg = compose("config_adm_edm_mcedm_res32", ["system=swe_per"])
torch.manual_seed(1)
pl = PlMcedm(copy.deepcopy(cfg.model.hparams))
randomize_zero_init(pl.model, 2)
pl = pl.to(dev).train()
opt = pl.configure_optimizers()["optimizer"]
opt.max_grad_norm = 1.0
h, tg, xg, u, mask = D.make_batch("swe_per", B, "train", seed=0)
batch = tuple(t.to(dev) for t in (h, tg, xg, u, mask))
st = D.field_stats("swe_per", 16)
pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))


def step(ev=None):
    opt.zero_grad(set_to_none=True)
    if ev: ev[0].record()
    loss = pl.training_step(batch, 0)
    if ev: ev[1].record()
    loss.backward()
    if ev: ev[2].record()
    pl.optimizer_step(0, 0, opt)
    if ev: ev[3].record()
    return loss


for _ in range(3):
    loss = step()
torch.cuda.synchronize()
L.check_watchdog()
n0 = L.LAUNCHES[0]
tot = [0.0, 0.0, 0.0]
import time
t0 = time.perf_counter()
for _ in range(steps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    loss = step(ev)
    torch.cuda.synchronize()
    for i in range(3):
        tot[i] += ev[i].elapsed_time(ev[i + 1])
wall = (time.perf_counter() - t0) / steps * 1e3
ms = [t / steps for t in tot]
print(f"B={B} loss={float(loss):.4f} fwd+loss {ms[0]:.2f} ms, bwd {ms[1]:.2f} ms, opt+ema {ms[2]:.2f} ms, "
      f"total {sum(ms):.2f} ms (wall {wall:.2f}) -> {B / sum(ms) * 1e3:.1f} samples/s; "
      f"{(L.LAUNCHES[0] - n0) // steps} launches/step; {56.305e9 * B / sum(ms) / 1e9:.1f} TFLOP/s")
