"""Per-role barrier-wait breakdown of conv_rows_fused (bring-up): MCEDM_DBG=32 python scripts/conv_rows_roles.py [B]
(MCEDM_DBG=64 instead times the MMA-issue and commit sections of the issuing lane: slots [1], [2])."""
import ctypes as C, os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcedm_b200 import _lib as L
from mcedm_b200.engine import pack_conv3x3
lib = L.lib(); dev = torch.device("cuda:0"); dt = torch.float16
B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 128
x = torch.randn(B, H, 128, 64, device=dev).to(dt)
coef = torch.cat([torch.rand(B, 64, device=dev) + 0.5, torch.randn(B, 64, device=dev) * 0.3], 1).contiguous()
w = pack_conv3x3(torch.randn(64, 64, 3, 3, device=dev) / 24, dtype=dt); bias = torch.randn(64, device=dev)
out = torch.empty(B, H, 128, 64, device=dev, dtype=dt); st = torch.empty(B * H, 4, 16, 2, device=dev)
cptr = (C.c_void_p * 1)(coef.data_ptr())
for _ in range(3):
    L.check(lib.mcedm_conv_rows_fused(L.ptr_array([x]), cptr, 1, None, 0, L.ptr(w), L.ptr(bias), B, H, 64, 0, 64, L.ptr(out), 1, None, 0, 0, 0, L.ptr(st), 1, L.stream_ptr()))
buf = np.zeros((160, 8), dtype=np.int64)
L.check(lib.mcedm_debug_rows(buf.ctypes.data_as(C.c_void_p)))
rows = B * H / 148
m = buf[:148].mean(0) / rows
print(f"DBG={os.environ.get('MCEDM_DBG')} per-row cycles: producer total {m[7]:.0f} (wait h_empty {m[0]:.0f}) | MMA total {m[6]:.0f} (wait acc_empty {m[1]:.0f}, h_ready {m[2]:.0f}) | "
      f"epilogue total {m[5]:.0f} (wait acc_full {m[3]:.0f}) | transform wait h_full {m[4]:.0f}")
