// fp32-accuracy inference mode ("1e-4 in fp32", BASELINE north_star): helpers.
//
// The mode keeps the tensor cores: every 16-bit operand is split into two fp16 terms (x = hi + lo, 22 significand
// bits) and every GEMM runs as hi*W_hi + lo*W_hi + hi*W_lo with fp32 accumulation (the dropped lo*W_lo term is 2^-22
// relative), on the unfused fp32-stream plan (hard part 3: plain TF32 misses 1e-4, a 3-term split
// passes).  This file holds what that plan needs beyond the regular kernels:
//   * split16       — fp32 tensor -> (hi, lo) fp16 pair (operands that no GroupNorm pass produces: attention output);
//   * attention_f32 — softmax(q k^T / 8) v in fp32 on the CUDA cores for fp32 q|k|v (AttentionOp computes its softmax
//                     in fp32, models/adm_blocks.py:103-109; the tcgen05 attention rounds P to 16 bits).
// Speed is secondary here: this is the validation-grade mode, not the throughput path.
#include "ptx.cuh"
#include "runtime.cuh"
#include "../../include/mcedm_b200.h"

namespace mcedm {

__global__ void __launch_bounds__(256) split16_kernel(const float4* __restrict__ x, long long n4, uint2* __restrict__ hi,
                                                      uint2* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 v = x[i];
  uint2 h;
  h.x = pack_f16x2(v.x, v.y);
  h.y = pack_f16x2(v.z, v.w);
  const float2 a = unpack_f16x2(h.x), b = unpack_f16x2(h.y);
  uint2 l;
  l.x = pack_f16x2(v.x - a.x, v.y - a.y);
  l.y = pack_f16x2(v.z - b.x, v.w - b.y);
  hi[i] = h;
  lo[i] = l;
}

// One warp per query; K and V rows are read through L2 (8 queries per CTA share them in L1).
__global__ void __launch_bounds__(256) attention_f32_kernel(const float* __restrict__ qkv, int L, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (qi >= L) return;
  const float* base = qkv + (long long)b * L * 192;
  // lane holds channels (2 lane, 2 lane + 1) of the query; a key's score is a warp reduction
  const float2 q2 = *reinterpret_cast<const float2*>(base + (long long)qi * 192 + 2 * lane);
  float m = -INFINITY;
  for (int j = 0; j < L; ++j) {
    const float2 k2 = *reinterpret_cast<const float2*>(base + (long long)j * 192 + 64 + 2 * lane);
    float s = fmaf(q2.x, k2.x, q2.y * k2.y);
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    m = fmaxf(m, s * 0.125f);
  }
  float l = 0.f;
  float2 acc = make_float2(0.f, 0.f);
  for (int j = 0; j < L; ++j) {
    const float2 k2 = *reinterpret_cast<const float2*>(base + (long long)j * 192 + 64 + 2 * lane);
    float s = fmaf(q2.x, k2.x, q2.y * k2.y);
#pragma unroll
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float pj = expf(s * 0.125f - m);
    l += pj;
    const float2 v2 = *reinterpret_cast<const float2*>(base + (long long)j * 192 + 128 + 2 * lane);
    acc.x = fmaf(pj, v2.x, acc.x);
    acc.y = fmaf(pj, v2.y, acc.y);
  }
  const float inv = 1.0f / l;
  *reinterpret_cast<float2*>(out + ((long long)b * L + qi) * 64 + 2 * lane) = make_float2(acc.x * inv, acc.y * inv);
}

}  // namespace mcedm

extern "C" int mcedm_split16(const float* x, long long n, void* hi16, void* lo16, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(n > 0 && n % 4 == 0, "split16: element count %lld must be a positive multiple of 4", n);
  const long long n4 = n / 4;
  split16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), n4, reinterpret_cast<uint2*>(hi16), reinterpret_cast<uint2*>(lo16));
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_attention_f32(const float* qkv_f32, int B, int L, float* out_f32, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(B >= 1 && L >= 1, "attention_f32: bad sizes");
  dim3 grid((L + 7) / 8, B);
  attention_f32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(qkv_f32, L, out_f32);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}
