"""Per-kernel time of one fused evaluation (CUDA events per launch, eager): python scripts/eval_breakdown.py [B]"""
import os, sys, copy, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200.adm_blocks import DhariwalUNet
from mcedm_b200.config import compose
from mcedm_b200.utils import randomize_zero_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
cfg = compose("config_adm_edm_mcedm_res32")
torch.manual_seed(1)
net = DhariwalUNet(copy.deepcopy(cfg.model.hparams))
randomize_zero_init(net, 2)
net = net.to(dev).eval()
x = torch.randn(B, 2, 128, 128, device=dev)
c = torch.randn(B, 2, 128, 128, device=dev)
nl = torch.tensor([0.3], device=dev)
prof = net.engine().profile_kernels(x, nl, c, repeats=5)
groups = collections.OrderedDict()
for p in prof:
    key = p["name"]
    if p["flops"]:
        key += f" [{p['flops'] / B / 1e9:.2f} GF/sample]"
    g = groups.setdefault(key, [0, 0.0, 0.0, 0.0])
    g[0] += 1; g[1] += p["ms"]; g[2] += p["flops"]; g[3] += p["bytes"]
tot = sum(p["ms"] for p in prof)
print(f"B={B}: {len(prof)} launches, {tot:.3f} ms (sum of per-launch CUDA-event times)")
for k, (n, ms, fl, by) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:55s} x{n:2d} {ms * 1e3:8.1f} us {100 * ms / tot:5.1f} %  {ms * 1e3 / n:7.1f} us each  "
          f"{fl / ms / 1e9 if fl else 0:7.0f} TFLOP/s {by / ms / 1e6:7.0f} GB/s")
