// K8 — fused validation / test reductions.
//
// Replaces, in PlMcedm.test_step / validation_step (models/mcedm.py:283-441) and the single-task / joint-state modules'
// test steps (models/ddim.py:372-533, :1700-1770):
//
//   xs_mean = mean over n_samples of xs                                   (mcedm.py:385-386)
//   MaskedLoss('l1')(xs_mean, state_gt, mask, loss_dim)                   (losses.py:62-78, mcedm.py:398)
//   MaskedLoss('l1')(inverse_data_transform(xs_mean), gt_unnorm, mask, loss_dim)  (mcedm.py:400-408, normalizer.py:28-29)
//   CorrelationLoss()(pred, target)                                       (losses.py:96-128)
//   scale_each_min_max(state)'s per-(sample, channel) min / max           (ddim.py:689-698)
//
// which are ~25 torch launches over the fp64 fields (mean, two multiplies by the mask per loss, abs-diff, three
// reductions, the inverse normalisation, ...) per mask name.  Here: ONE streaming pass over xs for both masked errors and
// the sample mean (HBM-bound: 8 n + 16 B read, 8 B written per field element), plus one pass for the correlation /
// range statistics.  All arithmetic in fp64 in the reference's expression order per element; partial sums are folded
// in a fixed order, so results are deterministic (and equal to the torch expressions to ~1e-15 relative).
#include "ptx.cuh"
#include "runtime.cuh"
#include "../../include/mcedm_b200.h"

namespace mcedm {

constexpr int kMaeThreads = 256;

// xs fp64 [n][b][P][C] (channel-last fields, n-major as `rearrange('(n b) ...')`), gt fp32 [b][P][C] (normalised),
// un-normalised ground truth as the two tensors the datamodule delivers: gt_un_a fp32 [b][P][Ca] (h) and gt_un_b fp32
// [b][P][C-Ca] (u) — no torch.cat —, mask fp32 [b][P][C]; loss channels [c0, c1).
// partial fp64 [gridDim.x][3] = (sum |m*mask - gt*mask|, sum |(m*div+sub)*mask - gt_un*mask|, sum mask) over loss channels.
__global__ void __launch_bounds__(kMaeThreads)
masked_mae_mean_kernel(const double* __restrict__ xs, int n, long long bpc, int C, const float* __restrict__ gt,
                       const float* __restrict__ gt_un_a, const float* __restrict__ gt_un_b, int Ca,
                       const float* __restrict__ mask, int c0, int c1,
                       const double* __restrict__ sub, const double* __restrict__ div, int clamp01,
                       double* __restrict__ mean_out, double* __restrict__ partial) {
  __shared__ double sh[kMaeThreads / 32][3];
  double a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const double inv_n = 1.0 / (double)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < bpc; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s = __dadd_rn(s, xs[(long long)k * bpc + i]);
    const double m = n == 1 ? s : __dmul_rn(s, inv_n) ;
    if (mean_out) mean_out[i] = m;
    const int c = (int)(i % C);
    if (c >= c0 && c < c1) {
      const float mk = mask[i];
      // pred * mask is fp64 * fp32 -> fp64; target * mask is fp32 * fp32 -> fp32 (losses.py:69-70)
      const double p1 = __dmul_rn(m, (double)mk);
      const double t1 = (double)__fmul_rn(gt[i], mk);
      a1 += fabs(__dsub_rn(p1, t1));
      if (gt_un_a) {
        const long long pix = i / C;
        const float tgt = c < Ca ? gt_un_a[pix * Ca + c] : gt_un_b[pix * (C - Ca) + (c - Ca)];
        double un = m;
        if (clamp01) un = fmin(fmax(un, 0.0), 1.0);                  // min_max normalisation clamps first (mcedm.py:192-194)
        un = __dadd_rn(__dmul_rn(un, div[c]), sub[c]);               // Normalizer inverse: x * divide + subtract
        const double p2 = __dmul_rn(un, (double)mk);
        const double t2 = (double)__fmul_rn(tgt, mk);
        a2 += fabs(__dsub_rn(p2, t2));
      }
      a3 += (double)mk;
    }
  }
  for (int off = 16; off; off >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    a3 += __shfl_xor_sync(0xffffffffu, a3, off);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sh[w][0] = a1;
    sh[w][1] = a2;
    sh[w][2] = a3;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int k = 0; k < kMaeThreads / 32; ++k) t += sh[k][threadIdx.x];
    partial[(long long)blockIdx.x * 3 + threadIdx.x] = t;
  }
}

// out[0] = sum1 / n_el, out[1] = sum2 / n_el, out[2] = n_el   (fixed-order fold of the per-CTA partials)
__global__ void masked_mae_fold_kernel(const double* __restrict__ partial, int n_cta, double* __restrict__ out) {
  if (threadIdx.x == 0) {
    double s[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < n_cta; ++i)
      for (int k = 0; k < 3; ++k) s[k] += partial[(long long)i * 3 + k];
    out[0] = s[0] / s[2];
    out[1] = s[1] / s[2];
    out[2] = s[2];
  }
}

// per (sample b, channel c) of pred fp64 [b][P][C] against target fp32 [b][P][C]: the Pearson correlation exactly as
// CorrelationLoss.calculate_correlation forms it (centred sums, zero denominators get 1e-7), and min / max of pred.
// grid = b * C CTAs, two passes over the channel's P values (the second one hits L2).
__global__ void __launch_bounds__(256)
corr_minmax_kernel(const double* __restrict__ pred, const float* __restrict__ target, long long P, int C,
                   double* __restrict__ corr, double* __restrict__ pmin, double* __restrict__ pmax) {
  __shared__ double sh[8][5];
  __shared__ double mean_x, mean_y;
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  const double* x = pred + (long long)b * P * C + c;
  const float* y = target + (long long)b * P * C + c;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double sx = 0.0, sy = 0.0, mn = INFINITY, mx = -INFINITY;
  for (long long i = threadIdx.x; i < P; i += blockDim.x) {
    const double v = x[i * C];
    sx += v;
    if (corr) sy += (double)y[i * C];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  for (int off = 16; off; off >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, off);
    sy += __shfl_xor_sync(0xffffffffu, sy, off);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if (lane == 0) {
    sh[w][0] = sx;
    sh[w][1] = sy;
    sh[w][2] = mn;
    sh[w][3] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, bb = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int k = 0; k < 8; ++k) {
      a += sh[k][0];
      bb += sh[k][1];
      lo = fmin(lo, sh[k][2]);
      hi = fmax(hi, sh[k][3]);
    }
    mean_x = a / (double)P;
    mean_y = bb / (double)P;
    if (pmin) pmin[blockIdx.x] = lo;
    if (pmax) pmax[blockIdx.x] = hi;
  }
  __syncthreads();
  if (!corr) return;
  double cxy = 0.0, cxx = 0.0, cyy = 0.0;
  const double mxv = mean_x;
  const float myf = (float)mean_y;          // torch.mean of the fp32 target is fp32; y_bar = y - mean stays fp32
  for (long long i = threadIdx.x; i < P; i += blockDim.x) {
    const double xb = x[i * C] - mxv;
    const double yb = (double)__fsub_rn(y[i * C], myf);
    cxy += yb * xb;
    cxx += xb * xb;
    cyy += yb * yb;
  }
  for (int off = 16; off; off >>= 1) {
    cxy += __shfl_xor_sync(0xffffffffu, cxy, off);
    cxx += __shfl_xor_sync(0xffffffffu, cxx, off);
    cyy += __shfl_xor_sync(0xffffffffu, cyy, off);
  }
  __syncthreads();
  if (lane == 0) {
    sh[w][0] = cxy;
    sh[w][1] = cxx;
    sh[w][2] = cyy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, bx = 0.0, by = 0.0;
    for (int k = 0; k < 8; ++k) {
      a += sh[k][0];
      bx += sh[k][1];
      by += sh[k][2];
    }
    double den = sqrt(bx * by);
    if (den == 0.0) den += 1e-7;
    corr[blockIdx.x] = a / den;
  }
}

}  // namespace mcedm

extern "C" int mcedm_masked_mae_mean(const double* xs, int n_samples, int b, long long pixels, int C, const float* gt,
                                     const float* gt_unnorm_a, const float* gt_unnorm_b, int Ca, const float* mask, int c0,
                                     int c1, const double* sub, const double* div, int clamp01, double* mean_out,
                                     double* partial_scratch, int n_cta, double* out3, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(n_samples >= 1 && b >= 1 && pixels >= 1 && C >= 1 && c0 >= 0 && c1 <= C && c0 < c1 && n_cta >= 1,
                "masked_mae_mean: bad arguments");
  MCEDM_REQUIRE(gt_unnorm_a == nullptr || (sub != nullptr && div != nullptr && Ca >= 1 && Ca <= C &&
                                           (Ca == C || gt_unnorm_b != nullptr)),
                "masked_mae_mean: un-normalised targets need statistics and both channel blocks");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  masked_mae_mean_kernel<<<n_cta, kMaeThreads, 0, st>>>(xs, n_samples, (long long)b * pixels * C, C, gt, gt_unnorm_a, gt_unnorm_b, Ca,
                                                       mask, c0, c1, sub, div, clamp01, mean_out, partial_scratch);
  MCEDM_CUDA(cudaGetLastError());
  masked_mae_fold_kernel<<<1, 32, 0, st>>>(partial_scratch, n_cta, out3);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_corr_minmax(const double* pred, const float* target, int b, long long pixels, int C, double* corr_bc,
                                 double* min_bc, double* max_bc, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(b >= 1 && pixels >= 1 && C >= 1, "corr_minmax: bad arguments");
  corr_minmax_kernel<<<b * C, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, target, pixels, C, corr_bc, min_bc,
                                                                                 max_bc);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}
