"""Benchmark of the m-cedm hot path on B200: EDM-sampled fields / second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): `mcedm edm_sampler Heun sampling, n_samples=1024 of 128x128 SWE
fields` — stochastic Heun, 50 steps (99 U-Net evaluations per field), observed-state mask blending,
synthetic SWE-periodic-shaped (h,u) fields, random-init ("stress") weights of the ADM U-Net.  One STEP =
one pass of the hot path over one batch: every rank samples `--fields` rows (default 1024; weak scaling:
the rows per GPU are fixed as N grows) in micro-batches of `--chunk` rows.  `value` = rows sampled by
all ranks / max-over-ranks device time, with inputs resident in HBM.  `e2e` = the same through the
public API (`PlMcedm.sample_edm` via mcedm_b200.dist) from pinned HOST buffers, with the H2D copies of
(cond, mask) and the D2H copy of the fp64 result inside the timed region.

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
Inputs (several GB of 16-bit activations per 256-row chunk) exceed the 126 MB L2 many times over, so no
explicit L2 flush is needed between iterations (stated in config.l2).
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_EVAL = 18.797e9          # per sample per U-Net forward
EVALS_PER_FIELD = 99


_STDOUT = sys.stdout


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fields", type=int, default=1024, help="rows sampled per GPU per step")
    ap.add_argument("--chunk", type=int, default=256, help="micro-batch of rows per sampler launch sequence")
    ap.add_argument("--timesteps", type=int, default=50)
    ap.add_argument("--ref-fields", type=int, default=1, help="fields per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-fields", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--train-batch", type=int, default=32, help="training samples per GPU per step (config 3: 256 / 8)")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput measurement")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tensor=float(d["bf16_tflops"]), tensor_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(tensor=1590.0, tensor_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample_fields(n_fields, timesteps, seed=0):
    """Times `n_fields` full trajectories of the oracle's sample_edm on all host cores. Returns (seconds, threads)."""
    import torch

    from mcedm_b200 import data as D
    from mcedm_b200.adm_blocks import DhariwalUNet
    from mcedm_b200.config import compose
    from mcedm_b200.utils import randomize_zero_init
    from oracle import edm_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = compose("config_adm_edm_mcedm_res32")
    torch.manual_seed(1)
    net = DhariwalUNet(copy.deepcopy(cfg.model.hparams))
    randomize_zero_init(net, 2)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    h, tg, xg, u, masks = D.make_batch("swe_per", n_fields, "eval", seed=seed)
    st = D.field_stats("swe_per", 16)
    state = torch.cat([(h - st["input_mean"]) / st["input_std"], (u - st["target_mean"]) / st["target_std"]], -1)
    mask = masks["u"]
    g = torch.Generator().manual_seed(seed)
    cond = O.get_cond_in(state, mask, torch.randn(state.shape, generator=g)).permute(0, 3, 1, 2).contiguous()
    mask_c = mask.permute(0, 3, 1, 2).contiguous()
    sp = dict(cfg.diff_sampler)
    sp["timesteps"] = timesteps
    noise = torch.randn(mask_c.shape, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.sample_edm(sd, dict(cfg.model.hparams.model), noise, cond, mask_c, sp,
                     lambda i, x: torch.randn(x.shape, dtype=x.dtype, generator=g))
    return time.perf_counter() - t0, threads


def cpu_train_step(batch, seed=0):
    """Times one reference-algorithm training step (loss + backward through the oracle, fp32 autograd) on all host
    cores. Returns (seconds, threads)."""
    import torch

    from mcedm_b200 import data as D
    from mcedm_b200.adm_blocks import DhariwalUNet
    from mcedm_b200.config import compose
    from mcedm_b200.utils import randomize_zero_init
    from oracle import edm_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = compose("config_adm_edm_mcedm_res32")
    torch.manual_seed(1)
    net = DhariwalUNet(copy.deepcopy(cfg.model.hparams))
    randomize_zero_init(net, 2)
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "resample" not in k)
          for k, v in net.state_dict().items()}
    h, tg, xg, u, mask = D.make_batch("swe_per", batch, "train", seed=seed)
    st = D.field_stats("swe_per", 16)
    state = torch.cat([(h - st["input_mean"]) / st["input_std"], (u - st["target_mean"]) / st["target_std"]], -1)
    g = torch.Generator().manual_seed(seed)
    x = state.permute(0, 3, 1, 2).contiguous()
    mask_c = mask.permute(0, 3, 1, 2).contiguous()
    cond = O.get_cond_in(state, mask, torch.randn(state.shape, generator=g)).permute(0, 3, 1, 2).contiguous()
    noise = torch.randn(x.shape, generator=g)
    sigma = (torch.randn([batch, 1, 1, 1], generator=g) * 1.2 - 1.2).exp()
    t0 = time.perf_counter()
    loss, _ = O.training_loss(sd, dict(cfg.model.hparams.model), x, sigma, noise, cond, mask_c)
    loss.backward()
    return time.perf_counter() - t0, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(min(args.warmup, 1)):
        cpu_sample_fields(1, max(2, args.timesteps // 10))
    times = []
    for _ in range(args.steps):
        dt, threads = cpu_sample_fields(args.ref_fields, args.timesteps)
        times.append(dt)
    total = sum(times)
    value = args.ref_fields * args.steps / total
    sample = f"{args.ref_fields} field(s) x {args.timesteps} Heun steps ({2 * args.timesteps - 1} net evals) per step, fp32, torch CPU"
    line = dict(impl="reference", metric="edm_sampled_fields_per_sec", value=value, unit="fields/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * total / args.steps, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="mcedm edm_sampler Heun sampling of 128x128 SWE fields (BASELINE configs[1])",
                            timesteps=args.timesteps, fields_per_step=args.ref_fields),
                cpu_baseline=dict(value=value, unit="fields/s", cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit="fields/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), file=_STDOUT, flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from mcedm_b200 import _lib as L
    from mcedm_b200 import data as D
    from mcedm_b200 import dist as MD
    from mcedm_b200.config import compose
    from mcedm_b200.mcedm import PlMcedm
    from mcedm_b200.utils import randomize_zero_init

    rank, world, local = MD.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = compose("config_adm_edm_mcedm_res32", ["system=swe_per", f"diff_sampler.timesteps={args.timesteps}"])
    torch.manual_seed(1)
    pl = PlMcedm(copy.deepcopy(cfg.model.hparams))
    randomize_zero_init(pl.model, 2)
    pl.ema_model.ma_model.load_state_dict(pl.model.state_dict())
    pl = pl.to(dev).eval()
    pl.use_cuda_graph = not args.no_graph
    sp = cfg.diff_sampler
    st = D.field_stats("swe_per", 16)
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))

    # synthetic conditioning: 16 distinct fields, each repeated fields/16 times (test_step's repeat(n_samples))
    rows, chunk = args.fields, min(args.chunk, args.fields)
    b = 16 if rows % 16 == 0 else 1
    h, tg, xg, u, masks = D.make_batch("swe_per", b, "eval", seed=1000 * rank)
    state = pl.data_transform(h.to(dev), u.to(dev))
    mask = masks["u"].to(dev)
    torch.manual_seed(MD.rank_seed(cfg.seed, rank))
    cond1 = pl.get_cond_in(state, mask, None, None).permute(0, 3, 1, 2).contiguous()
    cond_all = cond1.repeat(rows // b, 1, 1, 1).contiguous()
    mask_all = mask.permute(0, 3, 1, 2).repeat(rows // b, 1, 1, 1).contiguous()
    hu_shape = torch.empty(chunk, 2, 128, 128, device=dev)
    host_cond = cond_all.cpu().pin_memory()
    host_mask = mask_all.cpu().pin_memory()
    host_out = torch.empty(rows, 1, 128, 128, 2, dtype=torch.float64).pin_memory()

    def one_step(e2e: bool):
        outs = []
        for lo in range(0, rows, chunk):
            hi = min(rows, lo + chunk)
            if e2e:
                c = host_cond[lo:hi].to(dev, non_blocking=True)
                m = host_mask[lo:hi].to(dev, non_blocking=True)
            else:
                c, m = cond_all[lo:hi], mask_all[lo:hi]
            xs = pl.sample_edm(hu_shape[: hi - lo], c, m, sp, return_last=True)
            if e2e:
                host_out[lo:hi].copy_(xs, non_blocking=True)
            else:
                outs.append(xs[0, 0, 0, 0, 0])
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            one_step(e2e)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        one_step(False)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = L.LAUNCHES[0]
    ms = timed(False, args.steps)
    launches = L.LAUNCHES[0] - launches0
    clk = clocks.stop() if rank == 0 else None
    one_step(True)
    ms_e2e = timed(True, args.steps)
    L.check_watchdog()

    total_rows = rows * world
    value = total_rows * args.steps / (ms / 1e3)
    e2e_value = total_rows * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: conv_igemm (3x3 implicit GEMM, N = 64), CUDA events per launch
    roof = None
    if rank == 0:
        pk = peaks()
        unet = pl.ema_model.ma_model
        x = torch.randn(chunk, 2, 128, 128, device=dev)
        prof = unet.engine().profile_kernels(x, torch.tensor([0.1], device=dev), cond_all[:chunk])
        fused = unet.engine().fused and unet.engine().infer_fmt == 1

        def grp(names, pred=lambda p: True):
            sel = [p for p in prof if p["name"] in names and pred(p)]
            return sel, sum(p["flops"] for p in sel), sum(p["bytes"] for p in sel), sum(p["ms"] for p in sel)

        # dominant kernel: the row-resident 3x3 implicit GEMM of the 128x128 level (68 % of the network's FLOPs)
        dom_name = "conv_rows_fused" if fused else "conv_rows"
        dom, fl, by, t = grp((dom_name,), lambda p: p["flops"] > 1e9 * chunk)      # N = 64 launches (not the 16-wide head)
        flat, fl_f, by_f, t_f = grp(("conv_flat_fused", "conv_flat"))
        att, fl_a, by_a, t_a = grp(("attention",))
        gn, fl_g, by_g, t_g = grp(("gn_apply16", "gn_apply"))
        cin, fl_i, by_i, t_i = grp(("conv_in_tc16", "conv_in16", "conv_in"))
        total_ms = sum(p["ms"] for p in prof)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")    # dram bytes per launch from an `ncu --set full` capture
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("conv_rows_kernel<64,fused>" if fused else "conv_rows_kernel<64>")
        achieved = fl / (t * 1e-3) / 1e12
        roof = dict(bound="tensor", achieved=achieved, peak=pk["tensor_sustained"], unit="TFLOP/s",
                    frac=achieved / pk["tensor_sustained"], traffic=traffic,
                    kernel=("conv_rows_kernel<64, fused> (GroupNorm+SiLU-fused 3x3 implicit GEMM at 128x128, 64 input "
                            "channels -> 64 per launch, 16-bit activations)" if fused else
                            "conv_rows_kernel<64> (3x3 implicit GEMM at 128x128, 64->64 channels per launch)"),
                    launches_per_eval=len(dom), flops_per_eval=fl, bytes_per_eval=by, ms_per_eval=t,
                    achieved_gbs=by / max(1e-9, t) / 1e6, share_of_eval_ms=t / total_ms, eval_ms_sum_of_launches=total_ms,
                    launches_in_eval=len(prof),
                    peak_source=pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                    other_kernels=[
                        dict(kernel="conv_flat_kernel<64> (3x3 at 64x64 / 32x32, padded-flat)", bound="tensor", unit="TFLOP/s",
                             achieved=fl_f / max(1e-9, t_f) / 1e9, peak=pk["tensor_sustained"], ms_per_eval=t_f,
                             launches_per_eval=len(flat)),
                        dict(kernel="attn_kernel (softmax(QK^T/8)V at 32x32, L=1024)", bound="tensor", unit="TFLOP/s",
                             achieved=fl_a / max(1e-9, t_a) / 1e9, peak=pk["tensor_sustained"], ms_per_eval=t_a,
                             launches_per_eval=len(att)),
                        dict(kernel="gn_apply (stand-alone GroupNorm+SiLU+resample passes that remain)", bound="hbm",
                             unit="GB/s", achieved=by_g / max(1e-9, t_g) / 1e6, peak=pk["hbm"], ms_per_eval=t_g,
                             launches_per_eval=len(gn)),
                        dict(kernel="conv_in_tc (first 3x3 conv on cat([cond, x]), tensor cores, 2 MMAs per input row)", bound="hbm", unit="GB/s",
                             achieved=by_i / max(1e-9, t_i) / 1e6, peak=pk["hbm"], ms_per_eval=t_i,
                             launches_per_eval=len(cin))],
                    end_to_end_tflops=value * EVALS_PER_FIELD * FLOPS_PER_EVAL * (args.timesteps * 2 - 1) / 99 / 1e12 / world)
        roof["other_kernels"] += pde_kernel_rooflines(pl, dev, pk)
    train = None
    if not args.no_train:
        train = measure_training(args, cfg, dev, rank, world, barrier)
    cpu = None
    # the CPU baseline is a rank-0, N = 1 measurement: under torchrun the other ranks spin in the NCCL barrier on the
    # same host cores (and OMP_NUM_THREADS is forced to 1), which turned a 3 s sample into a 280 s one at N = 8
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if train is not None:
            dt, threads = cpu_train_step(2)
            train["cpu_baseline"] = dict(value=2 / dt, unit="samples/s", cores=threads, kind="port",
                                         sample="one training step (loss + backward) of 2 samples through the oracle "
                                                f"port of the reference algorithm, torch CPU fp32 autograd, {threads} threads")
        dt, threads = cpu_sample_fields(args.cpu_baseline_fields, args.timesteps)
        cpu = dict(value=args.cpu_baseline_fields / dt, unit="fields/s", cores=threads, kind="port",
                   sample=f"{args.cpu_baseline_fields} field x {args.timesteps} Heun steps of the same workload "
                          f"(oracle port of the reference algorithm, torch CPU fp32, {threads} threads)")
    if rank == 0:
        line = dict(metric="edm_sampled_fields_per_sec", value=value, unit="fields/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="fp16" if pl.ema_model.ma_model.engine().infer_fmt else "bf16", data="synthetic",
                    config=dict(operands="fp16 tensor-core operands AND fp16 activation storage in HBM (saturating), GroupNorm+SiLU "
                                         "applied inside the convs; fp32 accumulation / statistics, fp64 sampler state; per-step "
                                         "denoiser error 1.4e-3 vs the reference (bar 1e-2)",
                                workload="mcedm edm_sampler Heun sampling, n_samples=1024 of 128x128 SWE fields per GPU "
                                         "(BASELINE configs[1])", fields_per_gpu=rows, micro_batch=chunk,
                                timesteps=args.timesteps, net_evals_per_field=2 * args.timesteps - 1,
                                weights="random init (seed 1) + randomised zero-init tensors (seed 2)",
                                parallelism=f"independent rows sharded over {world} GPU(s), no data-path collective",
                                cuda_graph=not args.no_graph,
                                l2="inputs per iteration (GBs of activations) exceed the 126 MB L2; no explicit flush"),
                    clocks=clk,
                    e2e=dict(value=e2e_value, unit="fields/s", h2d_bytes_per_step=int(2 * rows * 2 * 128 * 128 * 4),
                             d2h_bytes_per_step=int(rows * 2 * 128 * 128 * 8), ms_per_step=ms_e2e / args.steps),
                    gpu_launches=launches, roofline=roof, cpu_baseline=cpu, train=train)
        print(json.dumps(line), file=_STDOUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def pde_kernel_rooflines(pl, dev, pk, B=256, n=20):
    """Achieved HBM GB/s of the PDE-residual kernels (K6) on a sampled batch in the sampler's own layout (float64 NCHW
    state): algorithmic bytes = 2 planes x 8 B read per cell (+ 8 B written per cell for the gradient), CUDA events
    around `n` launches on the launching stream; the 268 MB batch exceeds nothing but is streamed once per launch."""
    import torch

    x = torch.randn(B, 2, 128, 128, device=dev, dtype=torch.float64)
    f = pl.pde_loss
    nh, nu = pl.normalizer_input, pl.normalizer_target
    out = []
    for name, fn, byts in (
            ("swe_fv_loss_kernel (FORCE finite-volume residual + sum, fp64 NCHW state in)",
             lambda: f.residual(x[:, 0], x[:, 1], nh, nu), B * 128 * 128 * 16.0),
            ("swe_fv_grad_kernel (analytic residual gradient, one image row per CTA)",
             lambda: f.gradient(x[:, 0], x[:, 1], nh, nu, mode=0), B * 128 * 128 * 24.0)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out.append(dict(kernel=name, bound="hbm", unit="GB/s", achieved=byts / ms / 1e6, peak=pk["hbm"], ms_per_launch=ms,
                        batch=B))
    return out


def measure_training(args, cfg, dev, rank, world, barrier):
    """U-Net train samples/s (BASELINE metric, second half; configs[2]: masked mixed-conditioning EDM training, bf16
    tensor-core operands, data-parallel): every rank runs `--train-batch` samples per step through
    PlMcedm.training_step -> loss.backward() -> [NCCL all-reduce of the flat gradient] -> fused clip+Adam -> EMA."""
    import torch
    import torch.distributed as dist

    from mcedm_b200 import _lib as L
    from mcedm_b200 import data as D
    from mcedm_b200.mcedm import PlMcedm
    from mcedm_b200.utils import randomize_zero_init

    B = args.train_batch
    torch.manual_seed(1)
    pl = PlMcedm(copy.deepcopy(cfg.model.hparams))
    randomize_zero_init(pl.model, 2)
    pl = pl.to(dev).train()
    opt = pl.configure_optimizers()["optimizer"]
    opt.max_grad_norm = 1.0
    opt.grad_scale = 1.0 / world
    st = D.field_stats("swe_per", 16)
    pl.normalizer_input.set_stats(st["input_mean"].to(dev), st["input_std"].to(dev))
    pl.normalizer_target.set_stats(st["target_mean"].to(dev), st["target_std"].to(dev))
    host = [t.pin_memory() for t in D.make_batch("swe_per", B, "train", seed=17 + rank)]
    resident = tuple(t.to(dev) for t in host)

    def step(e2e):
        batch = tuple(t.to(dev, non_blocking=True) for t in host) if e2e else resident
        opt.zero_grad(set_to_none=True)
        loss = pl.training_step(batch, 0)
        loss.backward()
        if world > 1:
            dist.all_reduce(opt.flat_grads())
        pl.optimizer_step(0, 0, opt)
        return loss

    def timed(e2e, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = step(e2e)
            if e2e:
                loss.item()                                  # device -> host read of the step's loss
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(5):
        step(False)
    n0 = L.LAUNCHES[0]
    ms = timed(False, args.train_steps)
    launches = L.LAUNCHES[0] - n0
    ms_e2e = timed(True, args.train_steps)
    L.check_watchdog()
    per = ms / args.train_steps
    bytes_in = sum(t.numel() * t.element_size() for t in host)
    return dict(metric="unet_train_samples_per_sec", value=B * world * args.train_steps / (ms / 1e3), unit="samples/s",
                ms_per_step=per, batch_per_gpu=B, global_batch=B * world, steps=args.train_steps, dtype="bf16",
                tflops_per_gpu=56.305e9 * B / per / 1e9,
                e2e=dict(value=B * world * args.train_steps / (ms_e2e / 1e3), unit="samples/s",
                         h2d_bytes_per_step=bytes_in, d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.train_steps),
                gpu_launches=launches,
                config=dict(workload="masked mixed-conditioning EDM training step (BASELINE configs[2]): forward + loss + "
                                     "backward + grad all-reduce + clip + Adam + EMA", optimizer="Adam lr 2e-4, clip 1.0, EMA 0.999",
                            parallelism=f"data-parallel over {world} GPU(s), one NCCL all-reduce of the flat fp32 gradient",
                            cuda_graph=True))


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: anything a library prints on the way goes to stderr
    _STDOUT = sys.stdout
    sys.stdout = sys.stderr
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
