// K1c — 3x3 convolution for the narrow levels (W = 64, 32, 16): implicit GEMM on tcgen05 over a
// ZERO-PADDED FLAT pixel sequence, every input element fetched once.
//
// Same math as conv_igemm.cu / conv_rows.cu (models/adm_blocks.py:65-81 Conv2d.forward + fused residual
// :171).  A 128-row UMMA tile must be 128 consecutive shared-memory rows, which for W < 128 spans
// several image rows; with a dense layout a +-1 pixel shift would leak across row ends.  The bf16
// operand is therefore stored by the producing GroupNorm pass (gn.cu) in a padded layout:
//
//     image block (blk positions, a multiple of 128):  [P zeros][row 0: W px | 8 zeros][row 1 ...] ... [zeros]
//     P = W + 8 = row pitch;  position(b, y, x) = b*blk + (y+1)*P + x
//
// so the whole tensor is ONE flat sequence of 128-byte pixels in which the filter tap (dy, dx) is the
// constant offset dy*P + dx and every out-of-image neighbour is a stored zero.  A tile = 128
// consecutive positions; its 9 A operands are row-shifted UMMA descriptors into a ring of 128-position
// chunks (chunk t-1, t, t+1 are adjacent in the ring; two mirror slots keep them adjacent across the
// wrap).  Each chunk is one 16 KB TMA load and serves 3 tiles x 9 taps.  Rows that fall on padding
// (11 % at 64x64, 27 % at 32x32) are computed and dropped by the epilogue.
//
// Warp roles / epilogue exactly as conv_rows.cu (8 epilogue warps, 4-deep TMEM accumulator ring,
// per-(tile, lane quarter) GroupNorm partial sums).
#include "ptx.cuh"
#include "runtime.cuh"
#include "../../include/mcedm_b200.h"

#include <cuda_bf16.h>

namespace mcedm {

constexpr int kChunkBytes = 16 * 1024;

struct FlatParams {
  int n_slots;             // ring depth S (physical slots = S + 2 mirrors)
  int H, W, P;             // image size, row pitch of the padded operand
  int tiles_per_img;       // blk / 128
  long long total_tiles;   // B * tiles_per_img
  const float* bias;
  float* out;              // fp32 NHWC [B,H,W,N] (unpadded)
  const float* res;
  int res_mode;            // 0 none, 1 same res, 2 nearest-x2 of [B,H/2,W/2,N], 3 2x2 mean of [B,2H,2W,N]
  float* stats;            // [total_tiles][4][N/4][2]
  int fmt;                 // 16-bit operand format: 0 bf16, 1 fp16
  unsigned int* err;
};

template <int N>
struct FlatCfg {
  static constexpr int CH = 32;
  static constexpr int NCH = N / CH;
  static constexpr int U = 8;
  static constexpr int W_SEG_BYTES = N * 128;
  static constexpr int EPI_WARPS = 4 * NCH;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int STAGE_BYTES = EPI_WARPS * 32 * CH * 4;
  static constexpr int ACC_BUFS = 4;
  static constexpr int TMEM_COLS = (ACC_BUFS * N <= 256) ? 256 : 512;
};

__device__ __forceinline__ uint64_t flat_desc(uint32_t addr) {
  constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t lo = ((addr & 0x3FFFFu) >> 4) | (1u << 16);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ float4 flat_residual(const FlatParams& p, int N, int b, int y, int x, int c0) {
  if (p.res_mode == 1) {
    return *reinterpret_cast<const float4*>(p.res + (((long long)b * p.H + y) * p.W + x) * N + c0);
  } else if (p.res_mode == 2) {
    return *reinterpret_cast<const float4*>(
        p.res + (((long long)b * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) * N + c0);
  } else {
    const int Ws = p.W << 1;
    const float* r0 = p.res + (((long long)b * (p.H << 1) + 2 * y) * Ws + 2 * x) * N + c0;
    const float4 r00 = *reinterpret_cast<const float4*>(r0);
    const float4 r01 = *reinterpret_cast<const float4*>(r0 + N);
    const float4 r10 = *reinterpret_cast<const float4*>(r0 + (long long)Ws * N);
    const float4 r11 = *reinterpret_cast<const float4*>(r0 + (long long)Ws * N + N);
    return make_float4(0.25f * ((r00.x + r01.x) + (r10.x + r11.x)), 0.25f * ((r00.y + r01.y) + (r10.y + r11.y)),
                       0.25f * ((r00.z + r01.z) + (r10.z + r11.z)), 0.25f * ((r00.w + r01.w) + (r10.w + r11.w)));
  }
}

template <int N>
__global__ void __launch_bounds__(FlatCfg<N>::THREADS, 1)
conv_flat_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_a,
                 const FlatParams p) {
  using Cfg = FlatCfg<N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.n_slots;
  uint8_t* w_smem = smem;
  uint8_t* ring = w_smem + 9 * Cfg::W_SEG_BYTES;               // (S + 2) chunk slots, contiguous
  uint8_t* stage_smem = ring + (S + 2) * kChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_smem + Cfg::STAGE_BYTES);
  uint64_t* w_full = bars;
  uint64_t* acc_full = bars + 1;
  uint64_t* acc_empty = acc_full + Cfg::ACC_BUFS;
  uint64_t* c_full = acc_empty + Cfg::ACC_BUFS;                 // S
  uint64_t* c_empty = c_full + S;                               // S
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_empty + S);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long t_begin = p.total_tiles * blockIdx.x / gridDim.x;
  const long long t_end = p.total_tiles * (blockIdx.x + 1) / gridDim.x;
  const int n_tiles = (int)(t_end - t_begin);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_a);
    mbar_init(w_full, 1);
    for (int i = 0; i < Cfg::ACC_BUFS; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 32 * Cfg::EPI_WARPS);
    }
    for (int i = 0; i < S; ++i) {
      mbar_init(&c_full[i], 1);
      mbar_init(&c_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // local chunk k (0 .. n_tiles+1) = global chunk t_begin - 1 + k; tile j uses chunks j, j+1, j+2
    if (lane == 0) {
      mbar_expect_tx(w_full, (uint32_t)(9 * Cfg::W_SEG_BYTES));
      for (int s = 0; s < 9; ++s) tma_load_2d(w_smem + s * Cfg::W_SEG_BYTES, &tm_w, w_full, 0, s * N);
      for (int k = 0; k < n_tiles + 2; ++k) {
        const uint32_t slot = (uint32_t)k % (uint32_t)S, ph = ((uint32_t)k / (uint32_t)S) & 1u;
        mbar_wait(&c_empty[slot], ph ^ 1u, p.err, 0x3100 + slot);
        const bool mirror = slot < 2 && k >= S;
        mbar_expect_tx(&c_full[slot], mirror ? 2 * kChunkBytes : kChunkBytes);
        const long long row0 = (t_begin - 1 + k) * 128;       // may be -128 or past the end: zero-filled
        tma_load_2d(ring + slot * kChunkBytes, &tm_a, &c_full[slot], 0, (int)row0);
        if (mirror) tma_load_2d(ring + (S + slot) * kChunkBytes, &tm_a, &c_full[slot], 0, (int)row0);
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    // warp-uniform control flow, one elected lane issues (see conv_rows.cu)
    {
      const uint32_t idesc = umma_idesc_16(128, N, 0, 0, p.fmt);
      mbar_wait(w_full, 0, p.err, 0x3200);
      tc_fence_after();
      const uint32_t w_base = smem_u32(w_smem);
      const uint32_t ring_base = smem_u32(ring);
      int waited = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t buf = (uint32_t)j % Cfg::ACC_BUFS, aph = ((uint32_t)j / Cfg::ACC_BUFS) & 1u;
        mbar_wait(&acc_empty[buf], aph ^ 1u, p.err, 0x3300 + buf);
        while (waited < j + 3) {
          const uint32_t slot = (uint32_t)waited % (uint32_t)S, ph = ((uint32_t)waited / (uint32_t)S) & 1u;
          mbar_wait(&c_full[slot], ph, p.err, 0x3400 + slot);
          ++waited;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * N;
        // window = slots (j % S), +1, +2 (mirrors keep them contiguous); the tile itself is the middle chunk
        const uint32_t centre = ring_base + ((uint32_t)j % (uint32_t)S) * kChunkBytes + 128 * 128;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint64_t ad = flat_desc(centre + (uint32_t)(((ky - 1) * p.P + (kx - 1)) * 128));
            const uint64_t bd = flat_desc(w_base + (ky * 3 + kx) * Cfg::W_SEG_BYTES);
            if (elect_one()) {
              umma_f16(d_tmem, ad, bd, idesc, (ky | kx) != 0 ? 1u : 0u);
              umma_f16(d_tmem, ad + 2, bd + 2, idesc, 1u);
              umma_f16(d_tmem, ad + 4, bd + 4, idesc, 1u);
              umma_f16(d_tmem, ad + 6, bd + 6, idesc, 1u);
            }
          }
        }
        if (elect_one()) {
          umma_commit(&c_empty[(uint32_t)j % (uint32_t)S]);        // chunk j has no further user
          if (j == n_tiles - 1) {
            umma_commit(&c_empty[(uint32_t)(j + 1) % (uint32_t)S]);
            umma_commit(&c_empty[(uint32_t)(j + 2) % (uint32_t)S]);
          }
          umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================================= epilogue =======================================
    const int q = warp & 3;
    const int ew = warp - 2;
    const int ch = ew >> 2;
    uint8_t* my_stage = stage_smem + ew * (32 * Cfg::CH * 4);
    const int unit = lane & 7;
    const int row_in_it = lane >> 3;
    const int c0 = ch * Cfg::CH + unit * 4;
    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bz = *reinterpret_cast<const float4*>(p.bias + c0);
    for (int j = 0; j < n_tiles; ++j) {
      const long long tile = t_begin + j;
      const uint32_t buf = (uint32_t)j % Cfg::ACC_BUFS, aph = ((uint32_t)j / Cfg::ACC_BUFS) & 1u;
      const int b = (int)(tile / p.tiles_per_img);
      const int pos0 = (int)(tile - (long long)b * p.tiles_per_img) * 128 + q * 32;
      // decode this lane's 8 rows (position -> image row / column; padding rows are dropped)
      long long opix[8];
      float4 rr[8];
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        const int pos = pos0 + itr * 4 + row_in_it;
        const int row = pos / p.P;
        const int x = pos - row * p.P;
        const bool valid = (row >= 1) && (row <= p.H) && (x < p.W);
        opix[itr] = valid ? (((long long)b * p.H + (row - 1)) * p.W + x) : -1;
        rr[itr] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && p.res_mode != 0) rr[itr] = flat_residual(p, N, b, row - 1, x, c0);
      }
      mbar_wait(&acc_full[buf], aph, p.err, 0x3500 + buf);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * N + ch * Cfg::CH, v);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int pj = jj ^ (lane & 7);
        *reinterpret_cast<uint4*>(my_stage + lane * 128 + pj * 16) =
            make_uint4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
      }
      __syncwarp();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        const int row = itr * 4 + row_in_it;
        const int pu = unit ^ (row & 7);
        float4 a = *reinterpret_cast<const float4*>(my_stage + row * 128 + pu * 16);
        if (opix[itr] >= 0) {
          a.x += bz.x; a.y += bz.y; a.z += bz.z; a.w += bz.w;
          a.x += rr[itr].x; a.y += rr[itr].y; a.z += rr[itr].z; a.w += rr[itr].w;
          s1 += (a.x + a.y) + (a.z + a.w);
          s2 += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
          *reinterpret_cast<float4*>(p.out + opix[itr] * N + c0) = a;
        }
      }
      if (p.stats) {
#pragma unroll
        for (int off = 8; off < 32; off <<= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, off);
          s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        if (lane < 8)
          *reinterpret_cast<float2*>(p.stats + ((tile * 4 + q) * (N / 4) + ch * 8 + lane) * 2) = make_float2(s1, s2);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace mcedm

extern "C" int mcedm_flat_geometry(int H, int W, int* pitch, int* block_positions) {
  using namespace mcedm;
  MCEDM_REQUIRE(W >= 8 && W <= 64 && H >= 1, "flat layout supports 8 <= W <= 64 (W=%d)", W);
  const int P = W + 8;
  *pitch = P;
  *block_positions = ((H + 2) * P + 127) / 128 * 128;
  return 0;
}

extern "C" int mcedm_conv_flat(const void* src_flat, const void* w_packed, const float* bias, int B, int H, int W, int N,
                               float* out, const float* res, int res_mode, float* stats_partial, int op_fmt,
                               void* stream) {
  using namespace mcedm;
  int P = 0, blk = 0;
  int rc = mcedm_flat_geometry(H, W, &P, &blk);
  if (rc) return rc;
  MCEDM_REQUIRE(N == 64, "conv_flat: N=%d unsupported (64)", N);
  MCEDM_REQUIRE(res_mode >= 0 && res_mode <= 3 && (res_mode == 0 || res != nullptr), "conv_flat: bad residual mode");
  FlatParams p;
  memset(&p, 0, sizeof(p));
  p.H = H;
  p.W = W;
  p.P = P;
  p.tiles_per_img = blk / 128;
  p.total_tiles = (long long)B * p.tiles_per_img;
  p.bias = bias;
  p.out = out;
  p.res = res;
  p.res_mode = res_mode;
  p.stats = stats_partial;
  p.fmt = op_fmt ? 1 : 0;
  p.err = watchdog_ptr();
  MCEDM_REQUIRE(p.err != nullptr, "conv_flat: cannot allocate the watchdog word");
  MCEDM_REQUIRE((long long)B * blk < (1LL << 31), "conv_flat: tensor too large for 32-bit TMA coordinates");
  using Cfg = FlatCfg<64>;
  const int fixed = 1024 + 9 * Cfg::W_SEG_BYTES + Cfg::STAGE_BYTES + 512;
  int slots = (232448 - fixed) / kChunkBytes - 2;
  if (slots > 6) slots = 6;
  MCEDM_REQUIRE(slots >= 3, "conv_flat: shared memory budget");
  p.n_slots = slots;
  const int smem = fixed + (slots + 2) * kChunkBytes;
  CUtensorMap tm_w, tm_a;
  rc = make_tmap_rows64_bf16(&tm_w, w_packed, 9LL * N, N);
  if (rc) return rc;
  rc = make_tmap_rows64_bf16(&tm_a, src_flat, (long long)B * blk, 128);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MCEDM_CUDA(cudaFuncSetAttribute(conv_flat_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  long long grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  conv_flat_kernel<64><<<(unsigned)grid, Cfg::THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tm_w, tm_a, p);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}
