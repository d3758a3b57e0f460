"""Turns the outputs of scripts/capture_profiles.sh (gpurun_out/r2f_*) into the tracked summaries under profiles/:
    python scripts/make_profiles.py        (needs ncu on PATH to read the .ncu-rep files; no GPU)"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
O = "gpurun_out"


def summ(rep, picks):
    return subprocess.run([sys.executable, "scripts/ncu_summary.py", rep] + [str(p) for p in picks], capture_output=True, text=True).stdout


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(r)))
    return rows[0], rows[1], rows[2:]


def val(h, r, k):
    return float(r[h.index(k)].replace(",", ""))


def dram_bytes(h, u, r):
    tot = 0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(k)
        tot += float(r[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[i]]
    return int(tot)


def lines(path, pred):
    return "".join(l + "\n" for l in open(path, errors="ignore").read().splitlines() if pred(l))


TEN = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
DUR = "gpu__time_duration.sum"

# ------------------------------------------------------------------ conv_rows
h, u, b = raw(f"{O}/r2f_rows.ncu-rep")
t = lambda i: val(h, b[i], TEN)
d = lambda i: val(h, b[i], DUR)
md = f"""# ncu --set full: conv_rows_kernel<64, fused, RM> (round 2, FINAL kernels, B = 256 = the bench's micro-batch)

Command (one warm-up + one measured launch per configuration, `scripts/prof_fused.py`; the plain run exited 0 first;
`scripts/capture_profiles.sh 1`, summarised by `scripts/make_profiles.py`):
`ncu --set full --clock-control none --import-source on -k regex:conv_rows -c 12 -o gpurun_out/r2f_rows python scripts/prof_fused.py 256 1 rows`

Launches below: 1 = 64->64 3x3 at 128x128, operands already normalised (no transform; the data-gradient launches of the
training backward); 3 = GroupNorm+SiLU transform on, no residual (the common conv0); 5 = transform + 16-bit residual (the
common conv1); 7 = one N = 32 pass of a 128-channel conv (two halo sources); 11 = transform + 1x1 skip projection of two
raw sources (conv1 of the 128-channel decoder blocks).  The same program timed with CUDA events outside the profiler
(`gpurun_out/r2f_rows_plain.log`, single launches; 10-launch averages of launches 3 / 5 in `scripts/rows_ablate.py`:
301 / 363 us = 1026 / 853 TFLOP/s):

```
{lines(f'{O}/r2f_rows_plain.log', lambda l: l.startswith('rows'))}```

Headline: `sm__pipe_tensor_cycles_active` **{t(3):.1f} %** with the transform, no residual ({d(3):.0f} us under the profiler; round 1:
40.0 %, first half of round 2 at B = 128: 52.1 %), **{t(5):.1f} %** with the residual (33.9 % / 44.5 %), {t(1):.1f} % without the transform,
{t(11):.1f} % with the skip projection.  DRAM traffic per launch equals the algorithmic bytes (launch 3: {dram_bytes(h, u, b[3]) / 1e9:.2f} GB for
0.537 + 0.537 algorithmic - the written part trails because some of the output is still in L2 when the kernel ends;
launch 5: {dram_bytes(h, u, b[5]) / 1e9:.2f} GB for 1.61; launch 11: {dram_bytes(h, u, b[11]) / 1e9:.2f} GB for 2.15): every element is fetched once.

"""
md += summ(f"{O}/r2f_rows.ncu-rep", [1, 3, 5, 7, 11])
open("profiles/r2_ncu_conv_rows.md", "w").write(md)
traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the FINAL round-2 `ncu --set full` captures "
                       "(profiles/r2_ncu_conv_rows.md, r2_ncu_conv_flat.md), B = 256",
           "conv_rows_kernel<64,fused>": dram_bytes(h, u, b[5]),
           "conv_rows_kernel<64,fused> (no residual)": dram_bytes(h, u, b[3]),
           "conv_rows_kernel<64,fused> (1x1 skip of two raw sources)": dram_bytes(h, u, b[11]), "batch": 256}

# ------------------------------------------------------------------ conv_flat
h, u, b = raw(f"{O}/r2f_flat.ncu-rep")
t = lambda i: val(h, b[i], TEN)
md = f"""# ncu --set full: conv_flat_kernel<64, fused, FM> (round 2, FINAL kernels: warpgroup roles + setmaxnreg, 8 transform warps loading from global memory; B = 256)

Command (`scripts/capture_profiles.sh 1`; the plain run exited 0 first):
`ncu --set full --clock-control none --import-source on -k regex:conv_flat -c 10 -o gpurun_out/r2f_flat python scripts/prof_fused.py 256 1 flat`

Launches: 1 = 64x64, no transform; 3 = 64x64 transform, no residual; 5 = 64x64 transform + residual; 7 = 32x32 no
transform; 9 = 32x32 transform + residual.  CUDA-event timing of the same program outside the profiler (single launches,
`gpurun_out/r2f_flat_plain.log`; 10-launch averages are ~15 % lower: 101 / 116 / 36.4 us for launches 3 / 5 / 9):

```
{lines(f'{O}/r2f_flat_plain.log', lambda l: l.startswith('flat'))}```

`sm__pipe_tensor_cycles_active`: {t(1):.1f} % without the transform, {t(3):.1f} % with it, {t(5):.1f} % with the residual at 64x64; {t(7):.1f} % /
{t(9):.1f} % at 32x32 (8 tiles per CTA: prologue and drain are a third of the launch).  DRAM bytes = algorithmic bytes
(launch 3: {dram_bytes(h, u, b[3]) / 1e6:.0f} MB for 2 x 143 MB of padded-flat tensor incl. padding positions; the 72 KB of weights per CTA come from L2).

"""
md += summ(f"{O}/r2f_flat.ncu-rep", [1, 3, 5, 7, 9])
open("profiles/r2_ncu_conv_flat.md", "w").write(md)
traffic["conv_flat_kernel<64,fused> 64x64 residual"] = dram_bytes(h, u, b[5])
traffic["conv_flat_kernel<64,fused> 64x64"] = dram_bytes(h, u, b[3])
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)

# ------------------------------------------------------------------ attention
h, u, b = raw(f"{O}/r2f_attn.ncu-rep")
md = f"""# ncu --set full: attn_kernel (round 2, FINAL; B = 256, L = 1024, d = 64, fp16)

Command (`scripts/capture_profiles.sh 2`; the plain run exited 0 first):
`ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 4 -o gpurun_out/r2f_attn python scripts/attn_bench.py 256 x`

Launch 2 = `attn_kernel<single pass, fp16-pair exponentials, 4 softmax warps per lane quarter>` (the inference kernel:
2048 one-tile CTAs of 576 threads), launch 3 = the flagged-tile two-pass fallback (a 148-CTA grid whose CTAs read all their
flags in one round trip and leave when none is raised: {val(h, b[3], DUR):.1f} us).  CUDA events outside the profiler, single pass +
fallback per call (`gpurun_out/r2f_attn_plain.log`):

```
{lines(f'{O}/r2f_attn_plain.log', lambda l: l.startswith('MCEDM'))}```

The tensor pipe is active {val(h, b[2], TEN):.1f} % of the kernel: QK^T and PV are 0.27 GFLOP per sample, the time goes into one `MUFU.EX2` per
score (16 per clock per SM, measured on the box for the f32 and the f16x2 form alike: 268 M exponentials per call = 67 us at
1.7 GHz on 148 SMs) and into the per-tile prologue / epilogue of 13.8 waves of one-tile CTAs.

"""
md += summ(f"{O}/r2f_attn.ncu-rep", [2, 3])
open("profiles/r2_ncu_attn.md", "w").write(md)

# ------------------------------------------------------------------ training backward kernels
h, u, b = raw(f"{O}/r2f_gnbwd.ncu-rep")
g_read = val(h, b[1], "dram__bytes_read.sum")
md = f"""# ncu --set full: gn_bwd16_fused_kernel and conv_wgrad_kernel (training backward, round 2 FINAL; 32 samples per GPU)

Commands (`scripts/capture_profiles.sh 2`; plain runs exited 0 first):
`ncu --set full --clock-control none --import-source on -k regex:gn_bwd16 -c 4 -o gpurun_out/r2f_gnbwd python scripts/gn_bwd_bench.py 32`
`ncu --set full --clock-control none --import-source on -k regex:conv_wgrad --launch-skip 198 -c 8 -o gpurun_out/r2f_wgrad python scripts/train_breakdown.py 32 fused16`

## gn_bwd16 (GroupNorm + SiLU backward on raw fp16 activations: both passes in one kernel, cp.async-staged loads), 128x128, 64 channels

CUDA events outside the profiler (`gpurun_out/r2f_gnbwd_plain.log`; the two-kernel form with register-resident batches
measured 105.5 / 113.7 us at 128x128 on the same inputs):

```
{lines(f'{O}/r2f_gnbwd_plain.log', lambda l: 'us per call' in l)}```

DRAM: {g_read:.0f} MB read per launch = x and dy (67 MB each) TWICE: with all 32 samples in flight the pass-1 stream (134 MB)
exceeds what one L2 partition keeps, so pass 2 misses although it re-reads the same slice ~40 us later; what the one-kernel
form buys is the launch and, with the per-thread cp.async ring (three 64-byte batches per thread in flight, loads no longer
alternate with ~400 instructions of arithmetic per thread), 14-16 % of the time.  Achieved occupancy stays at 21 % (256 CTAs
of 256 threads, 2 resident per SM).  Waves of 8 samples (4 MB each) so that pass 2 hits L2 were tried and measured SLOWER (108.5 vs 88.1 us: four rounds of
per-sample hand-over and pipeline ramp per CTA cost more than the L2 hits return).

"""
md += summ(f"{O}/r2f_gnbwd.ncu-rep", [1])
h, u, b = raw(f"{O}/r2f_wgrad.ncu-rep")
md += f"""
## conv_wgrad_kernel (weight gradient with the GroupNorm+SiLU operand formed in shared memory)

Final form: straight-line MMA issue (one election per row, descriptors by 32-bit adds), kx-stacked N = 192 B operand
(three MN blocks ONE PIXEL = 128 B apart: 2 MMAs per K block instead of 6), two MMA-issuing warps on alternate rows, two
transform / epilogue warp sets.  Launches 0, 1 = 3x3 at 128x128 (conv1 / conv0 of a 128x128 block, 38.7 GFLOP each):
{val(h, b[0], DUR):.0f} / {val(h, b[1], DUR):.0f} us under the profiler, tensor pipe {val(h, b[0], TEN):.1f} % active (the first version of this round: 83-87 us, 28.8 % -
its single issuing thread was the pacer: ~3800 cycles of elections, 64-bit descriptor builds and reconvergence barriers per
row for 1536 cycles of MMAs, and an N = 64 MMA fetched 48 shared-memory wavefronts for 32 cycles of math); launches 2, 3 =
the 1x1 skip projection's two sources (HBM-bound: 139 MB in {val(h, b[2], DUR):.0f} us).  CUDA events outside the profiler
(`scripts/wgrad_bench.py`, `gpurun_out/r2f_wgrad_plain.log`; the first version: 82 / 267 / 54 / 34 / 29 us for the first
five lines below):

```
{lines(f'{O}/r2f_wgrad_plain.log', lambda l: l.startswith('B=') and 'xf=1' in l)}```

"""
md += summ(f"{O}/r2f_wgrad.ncu-rep", [0, 2])
open("profiles/r2_ncu_train_bwd.md", "w").write(md)

# ------------------------------------------------------------------ launch lists
shutil.copy(f"{O}/r2f_launches_eval_B256.csv", "profiles/r2_launches_fused_B256.csv")
shutil.copy(f"{O}/r2f_launches_train_B32.csv", "profiles/r2_launches_train_B32.csv")
ls = lambda *a: subprocess.run([sys.executable, "scripts/launch_summary.py", *a], capture_output=True, text=True).stdout
ev = ls("profiles/r2_launches_fused_B256.csv", "emb_mlp", "2")
tr = ls("profiles/r2_launches_train_B32.csv", "pack_gather", "2")
eval_live = lines(f"{O}/r2f_one_eval_plain.log", lambda l: "per evaluation" in l).strip()
train_live = lines(f"{O}/r2f_train_plain.log", lambda l: "samples/s" in l and l.startswith("B=")).strip()
open("profiles/r2_launches_fused_B256.md", "w").write(f"""# ncu launch list (gpu__time_duration.sum, --clock-control none): one U-Net evaluation, B = 256, fused 16-bit inference plan (round 2, FINAL tree)

Command (`scripts/capture_profiles.sh 1`): `python scripts/one_eval.py 256 4 t > plain.log && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches_eval_B256.csv python scripts/one_eval.py 256 4`
(`profiles/r2_launches_fused_B256.csv`; cold-cache, serialised launches: compare SHARES, not absolutes; window = the third
evaluation, from one `emb_mlp_kernel` to the next: `python scripts/launch_summary.py <csv> emb_mlp 2`.  The same evaluation
replayed from its CUDA graph and timed with CUDA events outside the profiler in the plain run of the same command on the
same box: `{eval_live}`; other boxes of the pool measured 7.97-8.17 ms for the same tree, 8.6-9.0 ms inside the 40 s bench
loop under `sw_power_cap`; round 1: 8.66 / 9.28 ms.)

{ev}
Template arguments: `conv_rows_kernel<N, fused, residual mode, epilogue>` (residual mode 0 none - includes the two launches
with the 1x1 skip projection of two raw sources -, 1 same-resolution 16-bit residual, 2 nearest-x2 residual; epilogue 0 =
16-bit staged, the shipped one); `conv_flat_kernel<64, fused, epilogue variant>`; `attn_kernel<single pass, fp16-pair
exponentials, softmax warps per lane quarter, row sums by MMA>`: `<1,1,4,0>` is the inference kernel, `<0,0,2,0>` the
flagged-tile two-pass fallback.  `gn_finalize_kernel` = `mcedm_gn_coef`, one per GroupNorm; `gn_apply_kernel<1>` = the 8
stand-alone GroupNorm passes that remain (4 resampling conv0 inputs, 4 qkv inputs; per shape in
`scripts/gn_apply_bench.py`: 197 / 143 / 65 / 46 / 4 x 37 us with L2 flushed = 4.1 / 4.7 / 3.1 / 3.6 / 1.8 TB/s);
`conv_igemm_kernel<64>` = 4 attention output projections (shifts instead of 64-bit divisions in the epilogue: 60 -> 41-47 us
each) + 4 skip projections at 64x64 / 32x32; `conv_in_tc_kernel<4>` with two builder warp sets (242 -> 168 us live),
`conv_rows_kernel<16, ...>` = the output head with two transform warp sets (212 -> 160 us live).

Shares against the FLOPs: the 128x128 convs (`conv_rows<64>`) take ~45 % of the time for 68 % of the FLOPs, the padded-flat
convs ~24 % for 25 %, everything else ~31 % for 7 %.
""")
open("profiles/r2_launches_train_B32.md", "w").write(f"""# ncu launch list (gpu__time_duration.sum, --clock-control none): one training step, 32 samples per GPU, "fused16" plan (round 2, FINAL tree)

Command (`scripts/capture_profiles.sh 1`): `python scripts/train_bench.py 32 5 > plain.log && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches_train_B32.csv python scripts/train_bench.py 32 2`
(`profiles/r2_launches_train_B32.csv`; window = one graph-replayed step, from one `pack_gather_kernel` to the next:
`python scripts/launch_summary.py <csv> pack_gather 2`).  Live, outside the profiler (plain run of the same program):
`{train_live}` - SHORTER than the serialised launches below add up to, because the weight-gradient launches run on a parallel
branch of the graph and fill the gaps and tails of the data-gradient chain.  Round 1: 9.6 ms per step.

{tr}
`conv_wgrad_kernel` ~22 % (66 launches: 16 at 128x128 of ~50 us, 50 at 64x64 / 32x32 - fixed-cost bound at
32 samples; 30 % / ~75 us before the issue-loop rewrite), `gn_bwd16` ~24 % (both passes in one kernel; `<0,0>` = the four resampling blocks), the data-gradient convs
(`conv_rows<64,1,0,0>` x11, `conv_flat<64,1,0>` x26: the forward kernels on 16-bit gradients) ~11 %, the forward's fused
convs ~12 %.  Per-call CUDA-event times of the same step run eagerly (`scripts/train_breakdown.py 32 fused16`):

```
{open(f'{O}/r2f_tb_fused16.txt').read()}```
""")
print("profiles/ regenerated")
