"""Launches conv_rows<64> (64->64 3x3 at 128x128, B=64, fp32 residual) a few times; ncu target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L
from mcedm_b200.engine import pack_conv3x3
dev = torch.device("cuda:0"); lib = L.lib()
B, H, N = 64, 128, 64
a = torch.randn(B, H, 128, 64, device=dev).to(torch.bfloat16)
w = pack_conv3x3(torch.randn(N, 64, 3, 3, device=dev) / 24)
bias = torch.randn(N, device=dev); res = torch.randn(B, H, 128, N, device=dev)
out = torch.empty(B, H, 128, N, device=dev); st = torch.empty(B * H, 4, 16, 2, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    L.check(lib.mcedm_conv_rows(L.ptr_array([a]), 1, None, 0, L.ptr(w), L.ptr(bias), B, H, N, L.ptr(out), 0, L.ptr(res), 1, L.ptr(st), 0, L.stream_ptr()))
torch.cuda.synchronize(); print("ok")
