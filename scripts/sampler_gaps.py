"""Where does a sampler step spend its time?  CUDA events around every network evaluation inside sample_edm.
python scripts/sampler_gaps.py [B] [steps]"""
import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcedm_b200.config import compose
from mcedm_b200.mcedm import PlMcedm
from mcedm_b200.utils import randomize_zero_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
cfg = compose("config_adm_edm_mcedm_res32")
torch.manual_seed(1)
pl = PlMcedm(copy.deepcopy(cfg.model.hparams))
randomize_zero_init(pl.model, 2)
pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
pl = pl.to(dev).eval()
sp = copy.deepcopy(cfg.diff_sampler)
sp.timesteps = steps
cond = torch.randn(B, 2, 128, 128, device=dev)
mask = torch.zeros(B, 2, 128, 128, device=dev)
mask[:, 1] = 1.0
hu = torch.zeros(B, 2, 128, 128, device=dev)
for _ in range(2):
    pl.sample_edm(hu, cond, mask, sp)
eng = pl.ema_model.ma_model.engine()
orig = eng.forward_static
evs = []

def timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(*a, **k)
    e1.record()
    evs.append((e0, e1))
    return r

eng.forward_static = timed
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
import time
h0 = time.perf_counter()
t0.record()
pl.sample_edm(hu, cond, mask, sp)
t1.record()
h1 = time.perf_counter()
torch.cuda.synchronize()
tot = t0.elapsed_time(t1)
inside = [a.elapsed_time(b) for a, b in evs]
gaps = [evs[i][1].elapsed_time(evs[i + 1][0]) for i in range(len(evs) - 1)]
print(f"B={B}: trajectory {tot:.2f} ms (host issue time {1e3 * (h1 - h0):.2f} ms); {len(evs)} evaluations: sum {sum(inside):.2f} ms, "
      f"mean {sum(inside) / len(inside):.3f}, min {min(inside):.3f}, max {max(inside):.3f}; gaps between evaluations: sum {sum(gaps):.2f} ms, "
      f"mean {sum(gaps) / len(gaps):.3f} ms; head {t0.elapsed_time(evs[0][0]):.3f} ms, tail {evs[-1][1].elapsed_time(t1):.3f} ms")
print("first evaluations (ms):", " ".join(f"{v:.3f}" for v in inside[:8]))
print("first gaps (ms):", " ".join(f"{v:.3f}" for v in gaps[:8]))
