"""Per-kernel time of one training step's forward / backward (CUDA events around every C-ABI call, eager, after warm-up).
usage: python scripts/train_breakdown.py [B] [plan]     plan = fused16 | fp32"""
import collections
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200.adm_blocks import DhariwalUNet  # noqa: E402
from mcedm_b200.config import compose  # noqa: E402
from mcedm_b200.utils import randomize_zero_init  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
plan = sys.argv[2] if len(sys.argv) > 2 else "fused16"
dev = torch.device("cuda:0")
cfg = compose("config_adm_edm_mcedm_res32")
torch.manual_seed(1)
net = DhariwalUNet(copy.deepcopy(cfg.model.hparams))
randomize_zero_init(net, 2)
net = net.to(dev).train()
eng = net.engine()
eng.train_plan = plan
x = torch.randn(B, 2, 128, 128, device=dev)
c = torch.randn(B, 2, 128, 128, device=dev)
nl = torch.randn(B, device=dev) * 0.5
dF = torch.randn(B, 2, 128, 128, device=dev) / B

for _ in range(3):
    eng.forward_train(x, nl, c)
    eng.backward(dF)
torch.cuda.synchronize()
real = eng.lib
recs = []


def level_of(name, a):
    """resolution tag of a launch from its integer arguments (first H-like value)"""
    vals = [v if isinstance(v, int) else getattr(v, "value", None) for v in a]
    ints = [v for v in vals if isinstance(v, int)]
    for cand in (128, 64, 32):
        if cand in ints:
            return cand
    return 0


class Proxy:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if not name.startswith("mcedm_") or name in ("mcedm_flat_geometry", "mcedm_last_error", "mcedm_gn_bwd_ctas_per_img",
                                                     "mcedm_wgrad_ctas"):
            return fn

        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            recs.append((phase[0], name[6:], level_of(name, a), e0, e1))
            return rc
        return call


phase = ["fwd"]
REP = 5
tot_ev = []
eng.lib = Proxy()
try:
    for _ in range(REP):
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        phase[0] = "fwd"
        t0.record()
        eng.forward_train(x, nl, c)
        t1.record()
        phase[0] = "bwd"
        eng.backward(dF)
        t2.record()
        tot_ev.append((t0, t1, t2))
finally:
    eng.lib = real
torch.cuda.synchronize()
groups = collections.OrderedDict()
for ph, name, lvl, e0, e1 in recs:
    g = groups.setdefault((ph, name, lvl), [0, 0.0])
    g[0] += 1
    g[1] += e0.elapsed_time(e1)
fw = sum(a.elapsed_time(b) for a, b, _ in tot_ev) / REP
bw = sum(b.elapsed_time(c_) for _, b, c_ in tot_ev) / REP
print(f"plan={plan} B={B}: eager fwd {fw:.2f} ms, bwd {bw:.2f} ms (host-bound when eager; the per-launch sums below are device times)")
for ph in ("fwd", "bwd"):
    tot = sum(v[1] for k, v in groups.items() if k[0] == ph) / REP
    print(f"--- {ph}: {sum(v[0] for k, v in groups.items() if k[0] == ph) // REP} calls, {tot:.3f} ms")
    for (p_, name, lvl), (n, ms) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        if p_ != ph:
            continue
        print(f"  {name:24s} @{lvl:3d} x{n // REP:3d} {ms / REP * 1e3:8.1f} us {100 * ms / REP / tot:5.1f} %  {ms / n * 1e3:7.1f} us each")
