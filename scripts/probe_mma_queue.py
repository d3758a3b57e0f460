"""How deep does tcgen05.mma queue?  cycles per iteration of (n_mma stacked MMAs + commit + `idle` cycles of nothing)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda:0")
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
iters = 400
for n_mma in (12, 24):
    for idle in (0, 100, 200, 400, 800, 1200, 1600):
        out = torch.zeros(n_sm, 3, dtype=torch.int64, device=dev)
        for _ in range(2):
            L.check(L.check_lib().mcedm_probe_mma_queue(iters, n_mma, idle, L.ptr(out), L.stream_ptr()))
        torch.cuda.synchronize()
        L.check_watchdog()
        m = out.double().mean(0) / iters
        print(f"n_mma={n_mma} (tensor {96 * n_mma} cyc) idle={idle}: total {m[0]:.0f}  issue {m[1]:.0f}  commit {m[2]:.0f} cycles/iteration")
