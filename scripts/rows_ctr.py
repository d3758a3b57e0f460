"""conv_rows_fused conv1 with the 1x1 skip projection of two raw sources (dec.128x128_block*): time per centre-ring depth.
python scripts/rows_ctr.py [B]"""
import ctypes as C, os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 2:
    import torch
    from mcedm_b200 import _lib as L
    from mcedm_b200.engine import pack_conv3x3
    lib = L.lib(); dev = torch.device("cuda:0"); dt = torch.float16
    B, H = int(sys.argv[1]), 128
    x = torch.randn(B, H, 128, 64, device=dev).to(dt)
    ctr = [torch.randn(B, H, 128, 64, device=dev).to(dt) for _ in range(2)]
    coef = torch.cat([torch.rand(B, 64, device=dev) + 0.5, torch.randn(B, 64, device=dev) * 0.3], 1).contiguous()
    w = torch.cat([pack_conv3x3(torch.randn(64, 64, 3, 3, device=dev) / 24, dtype=dt),
                   pack_conv3x3(torch.randn(64, 128, 1, 1, device=dev) / 11, dtype=dt)], 0).contiguous()
    bias = torch.randn(64, device=dev)
    out = torch.empty(B, H, 128, 64, device=dev, dtype=dt); st = torch.empty(B * H, 4, 16, 2, device=dev)
    cptr = (C.c_void_p * 1)(coef.data_ptr())
    def run():
        L.check(lib.mcedm_conv_rows_fused(L.ptr_array([x]), cptr, 1, L.ptr_array(ctr), 2, L.ptr(w), L.ptr(bias), B, H, 64, 0, 64,
                                          L.ptr(out), 1, None, 0, 0, 0, L.ptr(st), 1, L.stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    L.check_watchdog()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"CSLOTS={os.environ.get('MCEDM_CSLOTS', 'auto'):>4}: {us:7.1f} us  {2.0 * B * H * 128 * 64 * (576 + 128) / us / 1e6:6.0f} TFLOP/s "
          f"{B * H * 128 * 64 * 2 * 4 / us / 1e3:6.0f} GB/s")
else:
    B = sys.argv[1] if len(sys.argv) > 1 else "256"
    for cs in ("2", "3", "4", "auto"):
        env = dict(os.environ)
        if cs != "auto":
            env["MCEDM_CSLOTS"] = cs
        subprocess.run([sys.executable, __file__, B, "x"], env=env)
