// Bring-up / checker kernels. NOT on the product path:
//   * mcedm_probe_umma      — one-CTA tcgen05 experiment used by tests/ to pin down UMMA descriptor
//                             behaviour on real sm_100a silicon (row-shifted K-major A tiles with the
//                             descriptor base_offset field; MN-major B tiles), which decides what the
//                             fast kernels may rely on.
//   * mcedm_conv_direct_ref — plain CUDA-core direct convolution with the same segment semantics as
//                             mcedm_conv_igemm, used by the GPU tests as an on-device cross-check.
#include "ptx.cuh"
#include "runtime.cuh"
#include "../../include/mcedm_b200_check.h"

#include <cuda_bf16.h>

namespace mcedm {

// A: [a_rows x 64] bf16 row-major (a_rows <= 256), loaded by one TMA box into a 1024-aligned SW128 tile.
// Bm: [64 x 64] bf16 row-major. b_mn_major == 0: Bm is [N][K] (K-major);  == 1: Bm is [K][N] (MN-major).
// D[128 x 64] = A[row_shift : row_shift+128, :] * B^T, using an A descriptor whose start address is
// advanced by row_shift*128 bytes and whose base_offset field is `base_offset`.
__global__ void __launch_bounds__(128, 1)
probe_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int a_rows,
                  int row_shift, int base_offset, int b_mn_major, float* out, unsigned int* err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;                 // up to 256 rows * 128 B = 32 KB
  uint8_t* b_smem = smem + 32768;         // 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], (uint32_t)(a_rows * 128 + 8192));
    tma_load_2d(a_smem, &tm_a, &bars[0], 0, 0);
    tma_load_2d(b_smem, &tm_b, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0, err, 0x900);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, b_mn_major);
    const uint32_t a_base = smem_u32(a_smem) + row_shift * 128;
    const uint32_t b_base = smem_u32(b_smem);
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = umma_desc_k_sw128(a_base + k * 32, (uint32_t)base_offset);
      const uint64_t bd = b_mn_major ? umma_desc_mn_sw128(b_base + k * 2048, 8192) : umma_desc_k_sw128(b_base + k * 32);
      umma_f16(tmem_base, ad, bd, idesc, (uint32_t)(k != 0));
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0, err, 0x901);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

__global__ void conv_direct_ref_kernel(const __nv_bfloat16* s0, const __nv_bfloat16* s1, const __nv_bfloat16* s2,
                                       const __nv_bfloat16* s3, const int* seg, int n_seg,
                                       const __nv_bfloat16* w, const float* bias, int B, int H, int W, int N,
                                       float* out, const float* res, int res_mode) {
  const long long total = (long long)B * H * W * N;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx % N);
  const long long pix = idx / N;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const int b = (int)(pix / ((long long)W * H));
  const __nv_bfloat16* srcs[4] = {s0, s1, s2, s3};
  float acc = bias ? bias[n] : 0.f;
  for (int s = 0; s < n_seg; ++s) {
    const int yy = y + seg[3 * s + 1], xx = x + seg[3 * s + 2];
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const __nv_bfloat16* a = srcs[seg[3 * s]] + (((long long)b * H + yy) * W + xx) * 64;
    const __nv_bfloat16* wr = w + ((long long)s * N + n) * 64;
    float part = 0.f;
    for (int c = 0; c < 64; ++c) part += __bfloat162float(a[c]) * __bfloat162float(wr[c]);
    acc += part;
  }
  if (res_mode == 1) {
    acc += res[pix * N + n];
  } else if (res_mode == 2) {
    acc += res[(((long long)b * (H / 2) + y / 2) * (W / 2) + x / 2) * N + n];
  } else if (res_mode == 3) {
    const long long Ws = 2 * W;
    const float* r0 = res + (((long long)b * 2 * H + 2 * y) * Ws + 2 * x) * N + n;
    acc += 0.25f * ((r0[0] + r0[N]) + (r0[Ws * N] + r0[Ws * N + N]));
  }
  out[idx] = acc;
}

}  // namespace mcedm

namespace mcedm {
// Pure issue-rate probe: one CTA per SM issues `n_tiles` x 36 tcgen05.mma (M = 128, N, K = 16, both operands in
// shared memory with the conv kernels' descriptor pattern: 9 row-shifted A views x 4 K steps, 9 weight segments) and
// nothing else - no TMA, no epilogue, uninitialised operands.  cycles[cta] = clock64 ticks for the whole loop.  This
// is the tensor-pipe + shared-memory operand-fetch ceiling the conv kernels are measured against.
template <int N>
__global__ void __launch_bounds__(64, 1) probe_mma_rate_kernel(int n_tiles, long long* cycles, unsigned int* err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                          // 9 x N x 128 B
  uint8_t* a_smem = smem + 9 * N * 128;            // 3 input rows of 17 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + 3 * 17408);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_16(128, N, 0, 0, 1);
    const uint32_t w_lo = (smem_u32(w_smem) >> 4) | (1u << 16);
    const uint32_t a_lo = (smem_u32(a_smem) >> 4) | (1u << 16);
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(kHi) << 32) | lo; };
    const long long t0 = clock64();
    for (int t = 0; t < n_tiles; ++t) {
      const uint32_t buf = (uint32_t)t & 3u;
      if (t >= 4) mbar_wait(&bars[buf], (((uint32_t)t >> 2) - 1u) & 1u, err, 0x910 + buf);   // accumulator's previous use done
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint64_t ad = desc(a_lo + ky * (17408 >> 4) + kx * 8);
            const uint64_t bd = desc(w_lo + (ky * 3 + kx) * (N * 128 >> 4));
            umma_f16(tmem_base + buf * N, ad, bd, idesc, (ky | kx) != 0 ? 1u : 0u);
            umma_f16(tmem_base + buf * N, ad + 2, bd + 2, idesc, 1u);
            umma_f16(tmem_base + buf * N, ad + 4, bd + 4, idesc, 1u);
            umma_f16(tmem_base + buf * N, ad + 6, bd + 6, idesc, 1u);
          }
        }
        umma_commit(&bars[buf]);
      }
      __syncwarp();
    }
    // drain: wait for the last use of every accumulator buffer
    for (int t = (n_tiles > 4 ? n_tiles - 4 : 0); t < n_tiles; ++t)
      mbar_wait(&bars[t & 3], ((uint32_t)t >> 2) & 1u, err, 0x920 + (t & 3));
    const long long t1 = clock64();
    if (elect_one()) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}
// CUDA-core fp32 checker (tests only): one warp per query.
__global__ void attn_ref_kernel(const __nv_bfloat16* __restrict__ qkv, int L, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (qi >= L) return;
  const __nv_bfloat16* base = qkv + (long long)b * L * 192;
  float qv[64];
  for (int c = 0; c < 64; ++c) qv[c] = __bfloat162float(base[(long long)qi * 192 + c]);
  float m = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    float s = 0.f;
    for (int c = 0; c < 64; ++c) s += qv[c] * __bfloat162float(base[(long long)j * 192 + 64 + c]);
    m = fmaxf(m, s * 0.125f);
  }
  for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  float l = 0.f, acc[64];
  for (int c = 0; c < 64; ++c) acc[c] = 0.f;
  for (int j = lane; j < L; j += 32) {
    float s = 0.f;
    for (int c = 0; c < 64; ++c) s += qv[c] * __bfloat162float(base[(long long)j * 192 + 64 + c]);
    const float p = expf(s * 0.125f - m);
    l += p;
    for (int c = 0; c < 64; ++c) acc[c] += p * __bfloat162float(base[(long long)j * 192 + 128 + c]);
  }
  for (int off = 16; off; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
  for (int c = 0; c < 64; ++c) {
    float a = acc[c];
    for (int off = 16; off; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) out[((long long)b * L + qi) * 64 + c] = a / l;
  }
}

}  // namespace mcedm

extern "C" int mcedm_probe_mma_rate(int N, int n_tiles, long long* cycles_per_cta, void* stream) {
  using namespace mcedm;
  unsigned int* err = watchdog_ptr();
  MCEDM_REQUIRE(err != nullptr, "probe_mma_rate: no watchdog word");
  MCEDM_REQUIRE(n_tiles >= 1 && (N == 64 || N == 128 || N == 256), "probe_mma_rate: N in {64,128,256}");
  const int smem = 1024 + 9 * N * 128 + 3 * 17408 + 256;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 64) {
    MCEDM_CUDA(cudaFuncSetAttribute(probe_mma_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_mma_rate_kernel<64><<<num_sms(), 64, smem, st>>>(n_tiles, cycles_per_cta, err);
  } else if (N == 128) {
    MCEDM_CUDA(cudaFuncSetAttribute(probe_mma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_mma_rate_kernel<128><<<num_sms(), 64, smem, st>>>(n_tiles, cycles_per_cta, err);
  } else {
    MCEDM_REQUIRE(smem <= 232448, "probe_mma_rate: N=256 weights do not fit");
    MCEDM_CUDA(cudaFuncSetAttribute(probe_mma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_mma_rate_kernel<256><<<num_sms(), 64, smem, st>>>(n_tiles, cycles_per_cta, err);
  }
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

namespace mcedm {
// Issue-queue probe: per iteration one elected lane issues `n_mma` tcgen05.mma (M = 128, N = 192, K = 16, the stacked
// shape of conv_rows_fused) + one commit, then the warp idles `idle` cycles (what the issuing warp's waits / address
// arithmetic / loop control look like to the tensor pipe).  If MMAs queue deeply, cycles/iteration = max(tensor time,
// idle + issue); if issue is (nearly) synchronous with execution, it is their SUM.  out[cta] = {total, issue, commit}.
__global__ void __launch_bounds__(64, 1) probe_mma_queue_kernel(int iters, int n_mma, int idle, long long* out,
                                                                unsigned int* err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_smem = smem;                          // 192 x 128 B x 4 K-steps worth (re-read)
  uint8_t* a_smem = smem + 72 * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + 3 * 17408);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_16(128, 192, 0, 0, 1);
    const uint32_t w_lo = (smem_u32(w_smem) >> 4) | (1u << 16);
    const uint32_t a_lo = (smem_u32(a_smem) >> 4) | (1u << 16);
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(kHi) << 32) | lo; };
    long long t_issue = 0, t_commit = 0;
    const long long t0 = clock64();
    for (int t = 0; t < iters; ++t) {
      const long long c0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < n_mma; ++i) {
          const int kx = (i >> 2) % 3, ks = i & 3;
          umma_f16(tmem_base + (uint32_t)((t & 1) * 192), desc(a_lo + kx * 8 + ks * 2),
                   desc(w_lo + kx * 3 * (64 * 128 >> 4) + ks * 2), idesc, i != 0 ? 1u : 0u);
        }
      }
      __syncwarp();
      const long long c1 = clock64();
      if (elect_one()) umma_commit(&bars[t & 3]);
      __syncwarp();
      const long long c2 = clock64();
      t_issue += c1 - c0;
      t_commit += c2 - c1;
      while (clock64() - c2 < idle) {}
    }
    for (int t = (iters > 4 ? iters - 4 : 0); t < iters; ++t) {
      const int uses_before = t / 4;      // completions of bars[t & 3] before iteration t
      mbar_wait(&bars[t & 3], (uint32_t)uses_before & 1u, err, 0x930 + (t & 3));
    }
    const long long t1 = clock64();
    if (elect_one()) {
      out[blockIdx.x * 3 + 0] = t1 - t0;
      out[blockIdx.x * 3 + 1] = t_issue;
      out[blockIdx.x * 3 + 2] = t_commit;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
}  // namespace mcedm

extern "C" int mcedm_probe_mma_queue(int iters, int n_mma, int idle, long long* out3_per_cta, void* stream) {
  using namespace mcedm;
  unsigned int* err = watchdog_ptr();
  MCEDM_REQUIRE(err != nullptr, "probe_mma_queue: no watchdog word");
  MCEDM_REQUIRE(iters >= 1 && n_mma >= 1 && idle >= 0, "probe_mma_queue: bad arguments");
  const int smem = 1024 + 72 * 1024 + 3 * 17408 + 256;
  MCEDM_CUDA(cudaFuncSetAttribute(probe_mma_queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_mma_queue_kernel<<<num_sms(), 64, smem, reinterpret_cast<cudaStream_t>(stream)>>>(iters, n_mma, idle,
                                                                                        out3_per_cta, err);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_probe_umma(const void* a, int a_rows, const void* bm, int row_shift, int base_offset,
                                int b_mn_major, float* out, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(a_rows >= 128 && a_rows <= 256 && row_shift >= 0 && row_shift + 128 <= a_rows,
                "probe_umma: bad a_rows/row_shift");
  CUtensorMap tm_a, tm_b;
  int rc = make_tmap_rows64_bf16(&tm_a, a, a_rows, a_rows);
  if (rc) return rc;
  rc = make_tmap_rows64_bf16(&tm_b, bm, 64, 64);
  if (rc) return rc;
  unsigned int* err = watchdog_ptr();
  MCEDM_REQUIRE(err != nullptr, "probe_umma: no watchdog word");
  const int smem = 1024 + 32768 + 8192 + 64;
  MCEDM_CUDA(cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_umma_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tm_a, tm_b, a_rows, row_shift,
                                                                               base_offset, b_mn_major, out, err);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_conv_direct_ref(const void* const* src, int n_src, const int* seg_dev, int n_seg,
                                     const void* w_packed, const float* bias, int B, int H, int W, int N, float* out,
                                     const float* res, int res_mode, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(n_src >= 1 && n_src <= 4, "conv_direct_ref: n_src");
  const __nv_bfloat16* s[4];
  for (int i = 0; i < 4; ++i) s[i] = reinterpret_cast<const __nv_bfloat16*>(src[i < n_src ? i : 0]);
  const long long total = (long long)B * H * W * N;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  conv_direct_ref_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      s[0], s[1], s[2], s[3], seg_dev, n_seg, reinterpret_cast<const __nv_bfloat16*>(w_packed), bias, B, H, W, N, out,
      res, res_mode);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_attention_ref(const void* qkv_bf16, int B, int L, float* out_f32, void* stream) {
  using namespace mcedm;
  dim3 grid((L + 3) / 4, B);
  attn_ref_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), L, out_f32);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}
