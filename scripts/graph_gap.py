"""Per-node cost of a dependent kernel chain replayed from a CUDA graph: python scripts/graph_gap.py
(what one of the ~100 launches of an evaluation costs beyond its own work; the lever PDL / fewer launches acts on)."""
import torch

dev = torch.device("cuda:0")
x = torch.zeros(1024, device=dev)
for n in (100, 400):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            x.add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(n):
                x.add_(1.0)
    torch.cuda.synchronize()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{n} dependent tiny kernels per graph: {e0.elapsed_time(e1) / 10 / n * 1e3:.2f} us per node")
