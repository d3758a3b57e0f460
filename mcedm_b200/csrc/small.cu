// Small kernels around the U-Net trunk:
//   * emb_mlp      — noise embedding + all per-block `affine` projections in one launch
//                    (models/adm_blocks.py:185-199 PositionalEmbedding, :367-379 map_layer0/1,
//                     :163 affine(emb) of every UNetBlock).
//   * conv_in      — first 3x3 conv on cat([cond, x]) (models/adm_blocks.py:319-340, :384-385):
//                    NCHW fp32 in (<= 8 channels) -> NHWC fp32 64-channel out + GroupNorm partials.
//                    fp32 CUDA-core math: K = 9*Cin <= 72 is far too small for the tensor cores and
//                    the layer is bound by its 256 B/pixel store.
//   * head_to_nchw — picks the first Cout channels of the padded NHWC out_conv result -> NCHW F_x.
#include "ptx.cuh"
#include "runtime.cuh"
#include "../../include/mcedm_b200.h"

namespace mcedm {

// ------------------------------------------- emb MLP -------------------------------------------
// grid = (Bemb, n_aff), block = 128 (4 warps).  Every CTA recomputes the 64-wide embedding of its sample (two 64x64
// mat-vecs) and then applies ONE block's affine (128x64); mat-vec rows are spread over the warps with coalesced
// weight reads and a shuffle reduction.  (One CTA per sample doing all 16 affines serially with one strided weight
// row per thread took 80 us per evaluation in the sampling loop, where Bemb = 1.)
__device__ __forceinline__ float warp_dot64(const float* __restrict__ w_row, const float* x, int lane) {
  float acc = fmaf(w_row[lane], x[lane], w_row[lane + 32] * x[lane + 32]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  return acc;
}

__global__ void __launch_bounds__(128)
emb_mlp_kernel(const float* __restrict__ c_noise, const float* __restrict__ freqs, const float* __restrict__ w0,
               const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ b1,
               const float* __restrict__ aff_w, const float* __restrict__ aff_b, int n_aff, int Bemb,
               float* __restrict__ emb_out, float* __restrict__ out) {
  __shared__ float e[64], h0[64], h1[64];
  const int b = blockIdx.x, a = blockIdx.y, t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  if (t < 32) {
    const float ang = c_noise[b] * freqs[t];
    e[t] = cosf(ang);
    e[t + 32] = sinf(ang);
  }
  __syncthreads();
  for (int r = warp; r < 64; r += 4) {
    const float acc = warp_dot64(w0 + r * 64, e, lane);
    if (lane == 0) h0[r] = silu_f(acc + b0[r]);
  }
  __syncthreads();
  for (int r = warp; r < 64; r += 4) {
    const float acc = warp_dot64(w1 + r * 64, h0, lane);
    if (lane == 0) {
      const float v = silu_f(acc + b1[r]);
      h1[r] = v;
      if (emb_out && a == 0) emb_out[b * 64 + r] = v;
    }
  }
  __syncthreads();
  if (a < n_aff) {
    const float* w = aff_w + (long long)a * 128 * 64;
    for (int r = warp; r < 128; r += 4) {
      const float acc = warp_dot64(w + r * 64, h1, lane);
      if (lane == 0) out[((long long)a * Bemb + b) * 128 + r] = acc + aff_b[a * 128 + r];
    }
  }
}

// ------------------------------------------- conv_in -------------------------------------------
// Persistent CTAs loop over 128-pixel tiles (128/W image rows).  warp = four horizontally adjacent pixels at a
// time, lane = 2 output channels whose 2*Cin*9 weights stay in registers for the CTA's lifetime.  The input patch
// is staged in shared memory with a 16-byte-aligned row pitch so one LDS.128 + two LDS.32 feed the 3 taps of
// 4 pixels: 1 shared load per 8 FMAs (the previous one-pixel-per-warp version issued 1 per 2).
constexpr int kMaxCin = 8;

template <int CIN>
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ x, int Cx, const float* __restrict__ cond, int Cc,
               const float* __restrict__ w, const float* __restrict__ bias, int H, int W, int n_tiles,
               float* __restrict__ out, float* __restrict__ stats, int out16, int fmt) {
  extern __shared__ float patch[];  // [Cin][rows+2][W+8]; image column xx lives at index xx + 4
  __shared__ float sm[8][16][2];
  const int rows = 128 / W;
  const int tiles_per_img = H * W / 128;
  const int PW = W + 8, PH = rows + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // weights of this lane's 2 output channels: w[co][ci][ky][kx]
  float w0[CIN * 9], w1[CIN * 9];
#pragma unroll
  for (int i = 0; i < CIN * 9; ++i) {
    w0[i] = w[(2 * lane) * CIN * 9 + i];
    w1[i] = w[(2 * lane + 1) * CIN * 9 + i];
  }
  const float bz0 = bias ? bias[2 * lane] : 0.f, bz1 = bias ? bias[2 * lane + 1] : 0.f;
  const int quads_per_row = W >> 2;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int y0 = (tile - b * tiles_per_img) * rows;
    __syncthreads();                       // previous tile's patch / statistics are no longer read
    for (int rr = warp; rr < CIN * PH; rr += 8) {
      const int c = rr / PH, r = rr - c * PH;
      const int gy = y0 + r - 1;
      const bool row_ok = gy >= 0 && gy < H;
      const float* src = (c < Cc) ? cond + (((long long)b * Cc + c) * H + gy) * W
                                  : x + (((long long)b * Cx + (c - Cc)) * H + gy) * W;
      float* dst = patch + rr * PW + 3;    // image column -1
      for (int xx = lane; xx < W + 2; xx += 32) {
        const int gx = xx - 1;
        dst[xx] = (row_ok && gx >= 0 && gx < W) ? src[gx] : 0.f;
      }
    }
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int qd = warp; qd < 32; qd += 8) {
      const int r = qd / quads_per_row, x0 = (qd - r * quads_per_row) * 4;
      float a0[4] = {bz0, bz0, bz0, bz0}, a1[4] = {bz1, bz1, bz1, bz1};
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float* pr = patch + (c * PH + r + ky) * PW + x0 + 3;
          const float4 mid = *reinterpret_cast<const float4*>(pr + 1);
          const float v[6] = {pr[0], mid.x, mid.y, mid.z, mid.w, pr[5]};
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float u0 = w0[c * 9 + ky * 3 + kx], u1 = w1[c * 9 + ky * 3 + kx];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              a0[j] = fmaf(v[j + kx], u0, a0[j]);
              a1[j] = fmaf(v[j + kx], u1, a1[j]);
            }
          }
        }
      }
      const long long pix = (long long)tile * 128 + r * W + x0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (out16) reinterpret_cast<uint32_t*>(out)[(pix + j) * 32 + lane] = pack_op2(a0[j], a1[j], fmt);
        else *reinterpret_cast<float2*>(out + (pix + j) * 64 + 2 * lane) = make_float2(a0[j], a1[j]);
        s1 += a0[j] + a1[j];
        s2 += a0[j] * a0[j] + a1[j] * a1[j];
      }
    }
    if (stats) {
      // group of 4 channels = lanes (2g, 2g+1)
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
      if ((lane & 1) == 0) {
        sm[warp][lane >> 1][0] = s1;
        sm[warp][lane >> 1][1] = s2;
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
        float t = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) t += sm[ww][g][k];
        stats[(long long)tile * 32 + threadIdx.x] = t;
      }
    }
  }
}

// ----------------------------------------- head_to_nchw ----------------------------------------
__global__ void head_to_nchw_kernel(const float* __restrict__ src, int Cs, int Cout, long long HW, long long total,
                                    float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over B*Cout*HW (NCHW order)
  if (i >= total) return;
  const long long pix = i % HW;
  const int c = (int)((i / HW) % Cout);
  const long long b = i / (HW * Cout);
  dst[i] = src[(b * HW + pix) * Cs + c];
}

}  // namespace mcedm

extern "C" int mcedm_emb_mlp(const float* c_noise, const float* freqs, const float* w0, const float* b0,
                             const float* w1, const float* b1, const float* aff_w, const float* aff_b, int n_aff,
                             int Bemb, float* emb_out, float* out, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(Bemb >= 1 && n_aff >= 0, "emb_mlp: bad sizes");
  emb_mlp_kernel<<<dim3(Bemb, n_aff > 0 ? n_aff : 1), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(c_noise, freqs, w0, b0, w1, b1, aff_w,
                                                                            aff_b, n_aff, Bemb, emb_out, out);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

static int conv_in_impl(const float* x, int Cx, const float* cond, int Cc, const float* w, const float* bias,
                        int B, int H, int W, float* out, float* stats_partial, int out16, int fmt, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(Cx >= 1 && Cc >= 0 && Cx + Cc <= kMaxCin, "conv_in: %d+%d input channels exceed %d", Cx, Cc, kMaxCin);
  MCEDM_REQUIRE(Cc == 0 || cond != nullptr, "conv_in: cond channels without a cond tensor");
  MCEDM_REQUIRE(W >= 8 && W <= 128 && 128 % W == 0 && (H * W) % 128 == 0, "conv_in: W=%d H=%d unsupported", W, H);
  const int rows = 128 / W;
  const int smem = (Cx + Cc) * (rows + 2) * (W + 8) * (int)sizeof(float);
  const int n_tiles = B * (H * W / 128);
  const unsigned grid = (unsigned)(n_tiles < 4 * num_sms() ? n_tiles : 4 * num_sms());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define MCEDM_CONV_IN_CASE(C) \
  case C: conv_in_kernel<C><<<grid, 256, smem, st>>>(x, Cx, cond, Cc, w, bias, H, W, n_tiles, out, stats_partial, out16, fmt); break;
  switch (Cx + Cc) {
    MCEDM_CONV_IN_CASE(1) MCEDM_CONV_IN_CASE(2) MCEDM_CONV_IN_CASE(3) MCEDM_CONV_IN_CASE(4)
    MCEDM_CONV_IN_CASE(5) MCEDM_CONV_IN_CASE(6) MCEDM_CONV_IN_CASE(7) MCEDM_CONV_IN_CASE(8)
  }
#undef MCEDM_CONV_IN_CASE
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_conv_in(const float* x, int Cx, const float* cond, int Cc, const float* w, const float* bias,
                             int B, int H, int W, float* out, float* stats_partial, void* stream) {
  return conv_in_impl(x, Cx, cond, Cc, w, bias, B, H, W, out, stats_partial, 0, 0, stream);
}

extern "C" int mcedm_conv_in16(const float* x, int Cx, const float* cond, int Cc, const float* w, const float* bias,
                               int B, int H, int W, void* out16, float* stats_partial, int op_fmt, void* stream) {
  return conv_in_impl(x, Cx, cond, Cc, w, bias, B, H, W, reinterpret_cast<float*>(out16), stats_partial, 1,
                      op_fmt ? 1 : 0, stream);
}

extern "C" int mcedm_head_to_nchw(const float* src, int Cs, int Cout, int B, int H, int W, float* dst, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(Cout >= 1 && Cout <= Cs, "head_to_nchw: Cout=%d Cs=%d", Cout, Cs);
  const long long HW = (long long)H * W, total = HW * Cout * B;
  head_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, Cs, Cout, HW, total, dst);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}
