// K7 — the pieces the DDPM U-Net (`ddim_blocks.Model`, models/ddim_blocks.py:222-470) needs on top
// of the ADM kernels.  Its 3x3 / 1x1 convolutions, residuals, decoder concat and attention run on the fused 16-bit
// kernels as they are (conv_rows_fused / conv_flat_fused take per-(sample, channel) coefficients y = silu(a*x + b) and
// do not care how the groups were formed); what differs from the ADM network is
//
//   * Normalize = GroupNorm(32 groups, eps 1e-6) (:60-61): 2 channels per group at 64 channels, 4 at 128 — the conv
//     epilogues' partial sums are per 4-channel group, so the statistics come from a pass over the stored 16-bit tensor
//     that keeps them PER CHANNEL (gn_stats16), and gn_coef_groups forms any group size from them;
//   * ResnetBlock adds temb_proj(swish(temb)) per (sample, channel) between conv1 and norm2 (:140-146).  That sum is
//     never materialised: with per-channel sums the statistics of h + t follow analytically
//     (S' = S + n t, Q' = Q + 2 t S + n t^2) and y = a (h + t) + beta - mean' a = a h + (a t + beta - mean' a) is again
//     an affine map of the STORED h, i.e. one more term in the coefficient b;
//   * Downsample = pad (0,1,0,1) + stride-2 3x3 conv (:97-101) = the stride-1 "same" conv sampled at the odd positions
//     (2i+1, 2j+1): run the existing conv on the raw tensor and keep a quarter of it (decimate16);
//   * the timestep embedding: sinusoidal (sin | cos, :12-30) -> dense 64->256 -> swish -> dense 256->256, and per
//     ResnetBlock temb_proj(swish(temb)) 256 -> 64 (ddpm_temb).
#include "ptx.cuh"
#include "runtime.cuh"
#include "../../include/mcedm_b200.h"

#include <cuda_fp16.h>

namespace mcedm {

// ---------------------------------------------------------------------------------------------------------------------
// per-(sample, channel) sum and sum of squares of a 16-bit [B][npos][64] tensor (dense NHWC: npos = H*W; padded-flat:
// npos = block positions, whose padding is stored zeros and adds nothing).  grid = (nsplit, B), 256 threads =
// 8 channel chunks (16 B) x 32 position lanes; fp32 per thread over <= a few hundred positions, fixed-order fold.
__global__ void __launch_bounds__(256) gn_stats16_kernel(const uint4* __restrict__ x, long long npos, int fmt,
                                                         float* __restrict__ partial) {
  __shared__ float sh[32][8][16];
  const int b = blockIdx.y, nsplit = gridDim.x;
  const int c8 = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const long long p0 = npos * blockIdx.x / nsplit, p1 = npos * (blockIdx.x + 1) / nsplit;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  const uint4* base = x + (long long)b * npos * 8;
  for (long long pos = p0 + pl; pos < p1; pos += 32) {
    const uint4 v = __ldg(base + pos * 8 + c8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float lo, hi;
      if (fmt) {
        const float2 f = unpack_f16x2(w[e]);
        lo = f.x;
        hi = f.y;
      } else {
        lo = bf16_lo(w[e]);
        hi = bf16_hi(w[e]);
      }
      s[2 * e] += lo;
      q[2 * e] = fmaf(lo, lo, q[2 * e]);
      s[2 * e + 1] += hi;
      q[2 * e + 1] = fmaf(hi, hi, q[2 * e + 1]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sh[pl][c8][2 * e] = s[e];
    sh[pl][c8][2 * e + 1] = q[e];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = threadIdx.x >> 1, which = threadIdx.x & 1;     // channel, (sum | sum of squares)
    float acc = 0.f;
    for (int l = 0; l < 32; ++l) acc += sh[l][c >> 3][2 * (c & 7) + which];
    partial[(((long long)b * nsplit + blockIdx.x) * 64 + c) * 2 + which] = acc;
  }
}

// coefficients of y = act(a*x + b) for GroupNorm over groups of `cpg` consecutive channels of a 64-channel tensor
// (a slice of a wider normalisation: gamma / beta point at the slice), with an optional per-(sample, channel) shift t
// added to x BEFORE the normalisation (statistics and b corrected analytically, see the header).  grid = B, 64 threads.
__global__ void __launch_bounds__(64) gn_coef_groups_kernel(const float* __restrict__ partial, int nsplit, double count,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int cpg, float eps, const float* __restrict__ shift,
                                                            int shift_stride, float* __restrict__ coef) {
  __shared__ double sS[64], sQ[64];
  const int b = blockIdx.x, c = threadIdx.x;
  double S = 0.0, Q = 0.0;
  for (int i = 0; i < nsplit; ++i) {
    const float2 v = *reinterpret_cast<const float2*>(partial + (((long long)b * nsplit + i) * 64 + c) * 2);
    S += (double)v.x;
    Q += (double)v.y;
  }
  const double t = shift ? (double)shift[(long long)b * shift_stride + c] : 0.0;
  sS[c] = S + count * t;
  sQ[c] = Q + 2.0 * t * S + count * t * t;
  __syncthreads();
  const int g0 = c / cpg * cpg;
  double gs = 0.0, gq = 0.0;
  for (int i = 0; i < cpg; ++i) {
    gs += sS[g0 + i];
    gq += sQ[g0 + i];
  }
  const double n = count * cpg;
  const double mean = gs / n;
  double var = gq / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float a = rstd * gamma[c];
  const float bb = beta[c] - (float)mean * a;
  coef[(long long)b * 128 + c] = a;
  coef[(long long)b * 128 + 64 + c] = fmaf(a, (float)t, bb);
}

// out[b, i, j, :] = src[b, 2i+1, 2j+1, :] (16-bit, 64 channels); src / out dense NHWC (pitch 0) or padded-flat.
__global__ void __launch_bounds__(256) decimate16_kernel(const uint4* __restrict__ src, int in_pitch, int in_blk, int H,
                                                         int W, uint4* __restrict__ out, int out_pitch, int out_blk,
                                                         long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (b, y, x, chunk) of the OUTPUT
  if (i >= total) return;
  const int Ho = H >> 1, Wo = W >> 1;
  const int c8 = (int)(i & 7);
  long long r = i >> 3;
  const int xo = (int)(r % Wo);
  r /= Wo;
  const int yo = (int)(r % Ho);
  const int b = (int)(r / Ho);
  const int ys = 2 * yo + 1, xs = 2 * xo + 1;
  const long long sp = in_pitch > 0 ? (long long)b * in_blk + (long long)(ys + 1) * in_pitch + xs
                                    : ((long long)b * H + ys) * W + xs;
  const long long dp = out_pitch > 0 ? (long long)b * out_blk + (long long)(yo + 1) * out_pitch + xo
                                     : ((long long)b * Ho + yo) * Wo + xo;
  out[dp * 8 + c8] = __ldg(src + sp * 8 + c8);
}

__device__ __forceinline__ float swish_f(float x) { return x / (1.0f + expf(-x)); }

// temb = dense1(swish(dense0(sincos(t))))  (ddim_blocks.py:12-30, :422-425), then for every ResnetBlock
// out[blk][b][0:64] = temb_proj[blk](swish(temb))  (:140).  grid = Bt, 256 threads; fp32 throughout.
__global__ void __launch_bounds__(256) ddpm_temb_kernel(const float* __restrict__ t, const float* __restrict__ w0,
                                                        const float* __restrict__ b0, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ wp,
                                                        const float* __restrict__ bp, int nblk, int Bt,
                                                        float* __restrict__ out) {
  __shared__ float e[64], h0[256], h1[256];
  const int b = blockIdx.x, j = threadIdx.x;
  if (j < 32) {
    // emb = exp(arange(32) * -(log(10000) / 31)); [sin(t*emb) | cos(t*emb)]
    const float f = expf((float)j * -(logf(10000.0f) / 31.0f));
    const float a = t[b] * f;
    e[j] = sinf(a);
    e[32 + j] = cosf(a);
  }
  __syncthreads();
  {
    float acc = b0[j];
    for (int k = 0; k < 64; ++k) acc = fmaf(w0[j * 64 + k], e[k], acc);
    h0[j] = swish_f(acc);
  }
  __syncthreads();
  {
    float acc = b1[j];
    for (int k = 0; k < 256; ++k) acc = fmaf(w1[j * 256 + k], h0[k], acc);
    h1[j] = swish_f(acc);                    // every consumer applies swish(temb) first
  }
  __syncthreads();
  for (int o = j; o < nblk * 64; o += 256) {
    const int blk = o >> 6, c = o & 63;
    const float* w = wp + ((long long)blk * 64 + c) * 256;
    float acc = bp[blk * 64 + c];
    for (int k = 0; k < 256; ++k) acc = fmaf(w[k], h1[k], acc);
    out[((long long)blk * Bt + b) * 64 + c] = acc;
  }
}

}  // namespace mcedm

extern "C" int mcedm_gn_stats16(const void* x16, long long positions_per_img, int B, int op_fmt, int n_split,
                                float* partial, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(B >= 1 && positions_per_img >= 1 && n_split >= 1 && n_split <= 1024, "gn_stats16: bad arguments");
  dim3 grid(n_split, B);
  gn_stats16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(x16),
                                                                             positions_per_img, op_fmt ? 1 : 0, partial);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_gn_coef_groups(const float* partial, int n_split, long long pixels_per_img, const float* gamma,
                                    const float* beta, int channels_per_group, float eps, const float* shift,
                                    int shift_batch_stride, int B, float* coef_out, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(channels_per_group >= 1 && 64 % channels_per_group == 0, "gn_coef_groups: channels_per_group=%d",
                channels_per_group);
  MCEDM_REQUIRE(B >= 1 && n_split >= 1 && pixels_per_img >= 1, "gn_coef_groups: bad arguments");
  gn_coef_groups_kernel<<<B, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, n_split, (double)pixels_per_img, gamma, beta, channels_per_group, eps, shift, shift_batch_stride, coef_out);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_decimate16(const void* src16, int in_pitch, int in_blk, int B, int H, int W, void* out16,
                                int out_pitch, int out_blk, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(B >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "decimate16: even H, W");
  const long long total = (long long)B * (H / 2) * (W / 2) * 8;
  const int threads = 256;
  decimate16_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(src16), in_pitch, in_blk, H, W, reinterpret_cast<uint4*>(out16), out_pitch, out_blk,
      total);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_ddpm_temb(const float* t, int Bt, const float* w0, const float* b0, const float* w1, const float* b1,
                               const float* w_proj, const float* b_proj, int n_blocks, float* out, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(Bt >= 1 && n_blocks >= 1, "ddpm_temb: bad arguments");
  ddpm_temb_kernel<<<Bt, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, w0, b0, w1, b1, w_proj, b_proj, n_blocks, Bt,
                                                                          out);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}
