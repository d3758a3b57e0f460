"""One U-Net evaluation at a given batch (for ncu launch lists): python scripts/one_eval.py [B] [n_evals]"""
import os, sys, copy
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200.adm_blocks import DhariwalUNet
from mcedm_b200.config import compose
from mcedm_b200.utils import randomize_zero_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
cfg = compose("config_adm_edm_mcedm_res32")
torch.manual_seed(1)
net = DhariwalUNet(copy.deepcopy(cfg.model.hparams))
randomize_zero_init(net, 2)
net = net.to(dev).eval()
x = torch.randn(B, 2, 128, 128, device=dev)
c = torch.randn(B, 2, 128, 128, device=dev)
nl = torch.tensor([0.3], device=dev)
with torch.no_grad():
    for _ in range(n):
        y = net(x, nl, c)
    torch.cuda.synchronize()
    if len(sys.argv) > 3:      # live timing through the CUDA-graph replay path (not under a profiler)
        out = torch.empty_like(y)
        eng = net.engine()
        eng.forward_static(x, nl, c, out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.forward_static(x, nl, c, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"B={B}: {ms:.3f} ms per evaluation (graph replay) = {18.797e9 * B / ms / 1e9:.1f} TFLOP/s")
print("ok", float(y.abs().mean()))
