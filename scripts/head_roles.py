"""Per-role barrier-wait breakdown of conv_head_fused (conv_rows<16, fused>): MCEDM_DBG=32 python scripts/head_roles.py [B]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcedm_b200 import _lib as L
from mcedm_b200.engine import pack_conv3x3
lib = L.lib(); dev = torch.device("cuda:0"); dt = torch.float16
B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 128
x = torch.randn(B, H, 128, 64, device=dev).to(dt)
coef = torch.cat([torch.rand(B, 64, device=dev) + 0.5, torch.randn(B, 64, device=dev) * 0.3], 1).contiguous()
w16 = torch.zeros(16, 64, 3, 3, device=dev); w16[:2] = torch.randn(2, 64, 3, 3, device=dev) / 24
w = pack_conv3x3(w16, dtype=dt); bias = torch.zeros(16, device=dev)
out = torch.empty(B, 2, H, 128, device=dev)
run = lambda: L.check(lib.mcedm_conv_head_fused(L.ptr(x), L.ptr(coef), L.ptr(w), L.ptr(bias), B, H, 2, L.ptr(out), 1, L.stream_ptr()))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
print(f"DBG={os.environ.get('MCEDM_DBG', '0')}: {e0.elapsed_time(e1) * 100:.1f} us per launch")
buf = np.zeros((160, 8), dtype=np.int64)
L.check(lib.mcedm_debug_rows(buf.ctypes.data_as(C.c_void_p)))
rows = B * H / 148
m = buf[:148].mean(0) / rows
print(f"per-row cycles: producer total {m[7]:.0f} (wait h_empty {m[0]:.0f}) | MMA total {m[6]:.0f} (wait acc_empty {m[1]:.0f}, h_ready {m[2]:.0f}) | "
      f"epilogue total {m[5]:.0f} (wait acc_full {m[3]:.0f}) | transform wait h_full {m[4]:.0f}")
