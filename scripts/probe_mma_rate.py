"""Tensor-pipe + shared-memory operand-fetch ceiling of the conv kernels' MMA pattern (no TMA, no epilogue)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda:0")
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
for N in (64, 128):
    for tiles in (110, 1100):
        cyc = torch.zeros(n_sm, dtype=torch.int64, device=dev)
        L.check(L.check_lib().mcedm_probe_mma_rate(N, tiles, L.ptr(cyc), L.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.check_lib().mcedm_probe_mma_rate(N, tiles, L.ptr(cyc), L.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        L.check_watchdog()
        ms = e0.elapsed_time(e1)
        c = cyc.float()
        fl = 2.0 * 128 * N * 576 * tiles * n_sm
        print(f"N={N} tiles={tiles}: {float(c.mean()) / tiles:.0f} cycles/tile (ideal {36 * N // 2}), kernel {ms * 1e3:.1f} us, "
              f"{fl / ms / 1e9:.0f} TFLOP/s")
