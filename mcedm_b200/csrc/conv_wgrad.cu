// K1w — weight gradient of the 3x3 / 1x1 convolutions on tcgen05 (training backward of
// models/adm_blocks.py:65-81 Conv2d.forward as autograd differentiates it):
//
//     dW[co][ci][ky][kx] = sum_{b,y,x} dy[b,y,x,co] * a[b, y+ky-1, x+kx-1, ci]          (a zero outside the image)
//
// i.e. a [64 x 64] GEMM per filter tap whose contraction index is the PIXEL.  In NHWC both operands have
// the channel contiguous and the pixel strided, so both are MN-major UMMA operands: one pixel = one
// 128-byte shared-memory row, a K=16 MMA step = 16 consecutive pixels of one image row.
//
//   * B operand  = a   : image row y' as one TMA box (64 ch, W+2 px from x=-1): the kx shift is a UMMA
//                        descriptor start address advanced by kx*128 B (SWIZZLE_128B is a function of the
//                        absolute shared-memory address, same trick as conv_rows.cu), zero columns come
//                        from TMA out-of-bounds fill (dense layout) or stored zeros (padded-flat layout);
//   * A operand  = dy  : the ky shift is moved onto dy:  dW[ky] = sum_{y'} dy[y'-(ky-1)]^T a[y'].  dy rows
//                        live in a ring of image-row slots with two mirror slots (rows y'-1, y', y'+1 are
//                        always contiguous), so TWO taps are stacked into one M=128 MMA: the two
//                        64-channel M atoms are one slot (= LBO) apart.  Pair A = (ky=1 | ky=0) starts at
//                        row y', pair B = (ky=2 | duplicate) starts at row y'-1.
//   * 6 accumulators (pair x kx) of [128 x 64] fp32 stay in TMEM for the CTA's whole row range; one
//     epilogue at the end writes the CTA's partial dW, folded in fixed order by wgrad_reduce_kernel
//     (deterministic, no atomics).
//
// Layout codes for dy / a:  0 = dense NHWC [B,H,W,C];  1 = the padded-flat layout of conv_flat.cu
// (position(b,y,x) = b*blk + (y+1)*pitch + x, padding zeroed once by the buffer's owner).
#include "ptx.cuh"
#include "runtime.cuh"
#include "../../include/mcedm_b200.h"

#include <cuda_bf16.h>
#include <cstdlib>
#include <cstdio>

namespace mcedm {

struct WgradParams {
  int H, W;
  long long total_rows;    // B * H
  int taps;                // 9 or 1
  int dy_row_off, a_row_off;   // stored row = image row + off (1 for the padded-flat layout)
  int dy_coff, a_coff;     // first channel of the 64-channel block inside the tensor
  int n_slots;             // dy ring depth S (physical S + 2)
  int n_aslots;
  int dy_slot_bytes;       // W * 128
  int a_slot_bytes;        // (W + 2) * 128 rounded up to 1024
  float* partial;          // [gridDim.x][taps][64][64]
  uint32_t idesc;          // M = 128, N = 64, A and B both MN-major, operand format of the launch
  uint32_t idesc192;       // the same with N = 192 (kx-stacked B operand)
  const float* coef;       // NULL, or fp32 [B][128] = (a | b): `a` is a RAW activation and the operand is act(a*x + b),
                           // applied to each row in shared memory by the (otherwise idle) epilogue warps
  int act;                 // 1 SiLU, 0 identity
  int fmt;                 // 0 bf16, 1 fp16
  unsigned int* err;
  int dbg;                 // bring-up (MCEDM_WG_DBG): 1 skip the epilogue's stores, 2 skip the MMAs
};

// 352 threads: warp 0 TMA, warps 1 and 10 MMA issuers (alternate rows, issue order handed over as in conv_rows.cu), warps
// 2-5 and 6-9 two transform / epilogue SETS (alternate a rows; pair A / pair B)
__global__ void __launch_bounds__(352, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_a,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.n_slots, SA = p.n_aslots;
  uint8_t* dy_smem = smem;                                           // (S + 2) slots
  uint8_t* a_smem = dy_smem + (S + 2) * p.dy_slot_bytes;             // SA slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + SA * p.a_slot_bytes);
  uint64_t* acc_full = bars;
  uint64_t* dy_full = bars + 1;
  uint64_t* dy_empty = dy_full + S;
  uint64_t* a_full = dy_empty + S;
  uint64_t* a_empty = a_full + SA;
  uint64_t* a_ready = a_empty + SA;            // row transformed and visible to the async proxy (p.coef only)
  uint64_t* turn = a_ready + SA;                // 2: issue-order hand-over between the two MMA warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r_begin = p.total_rows * blockIdx.x / gridDim.x;
  const long long r_end = p.total_rows * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_dy);
    prefetch_tmap(&tm_a);
    mbar_init(acc_full, 1);
    for (int i = 0; i < S; ++i) {
      mbar_init(&dy_full[i], 1);
      mbar_init(&dy_empty[i], 1);
    }
    for (int i = 0; i < SA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&a_ready[i], 4);
    }
    mbar_init(&turn[0], 1);
    mbar_init(&turn[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t hl = 0, al = 0;
      long long r = r_begin;
      while (r < r_end) {
        const int b = (int)(r / p.H);
        const int y0 = (int)(r - (long long)b * p.H);
        const int R = (int)((r_end - r) < (long long)(p.H - y0) ? (r_end - r) : (long long)(p.H - y0));
        for (int k = 0; k < R + 2; ++k) {
          // dy row y0 - 1 + k (rows -1 and H are zero: out-of-bounds fill or stored padding)
          const uint32_t slot = hl % (uint32_t)S, ph = (hl / (uint32_t)S) & 1u;
          mbar_wait(&dy_empty[slot], ph ^ 1u, p.err, 0x4100 + slot);
          const bool mirror = slot < 2 && hl >= (uint32_t)S;
          mbar_expect_tx(&dy_full[slot], (uint32_t)(mirror ? 2 * p.dy_slot_bytes : p.dy_slot_bytes));
          tma_load_4d(dy_smem + slot * p.dy_slot_bytes, &tm_dy, &dy_full[slot], p.dy_coff, 0,
                      y0 - 1 + k + p.dy_row_off, b);
          if (mirror)
            tma_load_4d(dy_smem + (S + slot) * p.dy_slot_bytes, &tm_dy, &dy_full[slot], p.dy_coff, 0,
                        y0 - 1 + k + p.dy_row_off, b);
          ++hl;
          if (k >= 2) {
            // a row y0 + k - 2 with a one-pixel halo on both sides
            const uint32_t as = al % (uint32_t)SA, aph = (al / (uint32_t)SA) & 1u;
            mbar_wait(&a_empty[as], aph ^ 1u, p.err, 0x4200 + as);
            mbar_expect_tx(&a_full[as], (uint32_t)((p.W + 2) * 128));
            tma_load_4d(a_smem + as * p.a_slot_bytes, &tm_a, &a_full[as], p.a_coff, -1, y0 + k - 2 + p.a_row_off, b);
            ++al;
          }
        }
        r += R;
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ====================================== MMA issuer ======================================
    // Two issuing warps run the same control flow (every barrier phase is observed by both) and issue alternate rows:
    // tcgen05.mma queues only ~2 instructions deep, so one issuer's waits, commits and loop control left the tensor pipe
    // idle ~1000 cycles per row.  All rows accumulate into the same TMEM columns: the issue ORDER is handed over through
    // turn[] right after a row's last MMA (the pipe executes in issue order), and commits - which track only their own
    // thread's MMAs - therefore also cover the other warp's earlier rows.
    const uint32_t my_par = (warp == 1) ? 0u : 1u;
    uint32_t my_n = 0, last_par = 0;
    const uint32_t idesc = p.idesc;
    const uint32_t dy_base = smem_u32(dy_smem);
    const uint32_t a_base = smem_u32(a_smem);
    const uint32_t lbo = (uint32_t)p.dy_slot_bytes;
    const int kblocks = p.W >> 4;
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO = 1024 | version 1 | SWIZZLE_128B
    const uint32_t loA = ((lbo >> 4) & 0x3FFFu) << 16;                         // LBO of the dy operand: one slot
    const uint32_t loB = ((8192u >> 4) & 0x3FFFu) << 16;                       // LBO of the a operand (unused: one MN block)
    auto mk = [](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
    uint32_t hbase = 0, waited = 0, ac = 0;
    uint32_t first = 1;
    long long r = r_begin;
    while (r < r_end) {
      const int b = (int)(r / p.H);
      const int y0 = (int)(r - (long long)b * p.H);
      const int R = (int)((r_end - r) < (long long)(p.H - y0) ? (r_end - r) : (long long)(p.H - y0));
      for (int j = 0; j < R; ++j) {
        while (waited < hbase + j + 3) {
          const uint32_t slot = waited % (uint32_t)S, ph = (waited / (uint32_t)S) & 1u;
          mbar_wait(&dy_full[slot], ph, p.err, 0x4300 + slot);
          ++waited;
        }
        const uint32_t as = ac % (uint32_t)SA, aph = (ac / (uint32_t)SA) & 1u;
        mbar_wait(p.coef ? &a_ready[as] : &a_full[as], aph, p.err, 0x4400 + as);
        tc_fence_after();
        // window slots w, w+1, w+2 hold dy rows y'-1, y', y'+1 (mirrors keep them contiguous)
        const uint32_t w0 = dy_base + ((hbase + j) % (uint32_t)S) * lbo;
        const uint32_t arow = a_base + as * (uint32_t)p.a_slot_bytes;
        // The issuing thread is this kernel's pacer (ncu: the MMA warp never waits, the transform warps wait 40 % of the
        // time for their TMA, the tensor pipe is active 29 %): ONE election per row, and every descriptor is (a word
        // formed once per row) + (a small multiple of the loop counters) - the per-tap elections, 64-bit descriptor
        // builds and reconvergence barriers of the first version cost ~3800 cycles per row for 1536 cycles of MMAs.
        const uint32_t a1 = loA | (((w0 + lbo) & 0x3FFFFu) >> 4);    // (ky=1 | ky=0), K block 0
        const uint32_t a2 = loA | ((w0 & 0x3FFFFu) >> 4);            // (ky=2 | duplicate)
        const uint32_t b0 = loB | ((arow & 0x3FFFFu) >> 4);          // a row, pixel 0 (= kx 0), K block 0
        const bool mine = (ac & 1u) == my_par;
        last_par = ac & 1u;
        if (mine) {
          if (my_par == 1u) mbar_wait(&turn[1], my_n & 1u, p.err, 0x4a01);
          else if (my_n > 0) mbar_wait(&turn[0], (my_n - 1u) & 1u, p.err, 0x4a00);
        }
        if (mine && elect_one() && !(p.dbg & 2)) {
          uint32_t acc = first ? 0u : 1u;
          if (p.taps == 9) {
            if (p.dbg & 4) {        // MCEDM_WG_DBG=4: one N = 64 MMA per (pair, kx) - the first version, A/B switch
#pragma unroll 2
              for (int kb = 0; kb < kblocks; ++kb) {
                const uint64_t adA = mk(a1 + (uint32_t)kb * 128u), adB = mk(a2 + (uint32_t)kb * 128u);   // + kb * 2048 B
                const uint32_t bk = b0 + (uint32_t)kb * 128u;                                              // + 16 pixels
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  const uint64_t bd = mk(bk + (uint32_t)kx * 8u);                                          // + kx pixels
                  umma_f16(tmem_base + kx * 64, adA, bd, idesc, acc);
                  umma_f16(tmem_base + (3 + kx) * 64, adB, bd, idesc, acc);
                }
                acc = 1u;
              }
            } else {
              // kx-STACKED B operand: N = 192 = three 64-channel MN blocks whose leading-dimension offset is ONE PIXEL
              // (128 B): block n is the same a row shifted by n pixels, i.e. the tap kx = n (SWIZZLE_128B is a function
              // of the absolute shared-memory address, so a block that starts 128 B later is the row-shifted view the
              // single-tap descriptors already used).  Its 192 accumulator columns are the three 64-column accumulators
              // (kx = 0, 1, 2) of a pair, which already sit next to each other.  2 MMAs per K block instead of 6: an
              // N = 64 MMA fetched 4 KB of dy for 2 KB of a (48 shared-memory wavefronts for 32 cycles of math).
              const uint32_t bst = b0 - loB + ((128u >> 4) << 16);
#pragma unroll 2
              for (int kb = 0; kb < kblocks; ++kb) {
                const uint64_t bd = mk(bst + (uint32_t)kb * 128u);
                umma_f16(tmem_base, mk(a1 + (uint32_t)kb * 128u), bd, p.idesc192, acc);
                umma_f16(tmem_base + 192, mk(a2 + (uint32_t)kb * 128u), bd, p.idesc192, acc);
                acc = 1u;
              }
            }
          } else {
            for (int kb = 0; kb < kblocks; ++kb) {
              umma_f16(tmem_base + 64, mk(a1 + (uint32_t)kb * 128u), mk(b0 + (uint32_t)kb * 128u + 8u), idesc, acc);
              acc = 1u;
            }
          }
        }
        first = 0;
        if (mine && elect_one()) {
          mbar_arrive(&turn[my_par ^ 1u]);         // row issued: the other warp may issue the next one
          umma_commit(&a_empty[as]);
          umma_commit(&dy_empty[(hbase + j) % (uint32_t)S]);
          if (j == R - 1) {
            umma_commit(&dy_empty[(hbase + j + 1) % (uint32_t)S]);
            umma_commit(&dy_empty[(hbase + j + 2) % (uint32_t)S]);
          }
        }
        __syncwarp();
        if (mine) ++my_n;
        ++ac;
      }
      hbase += R + 2;
      r += R;
    }
    if ((ac == 0 ? my_par == 0u : last_par == my_par) && elect_one()) umma_commit(acc_full);   // by the last row's issuer
    __syncwarp();
  } else {
    // ============== GroupNorm + SiLU transform of the `a` rows (raw activations), then the epilogue ==============
    // thread t owns the logical 16-byte chunk j = t & 7 (channels 8j .. 8j+7) of pixels 1 + (t >> 3) + 16 i of every
    // row (pixel 0 and pixel W+1 are the zero halo and stay zero); physical chunk = j ^ (pixel & 7) (SWIZZLE_128B,
    // slots are 1 KB aligned) - the scheme of conv_rows.cu's transform warps.
    if (p.coef != nullptr) {
      const int xset = ((int)threadIdx.x - 64) >> 7;          // transform set: a rows ac = xset (mod 2)
      const int t = ((int)threadIdx.x - 64) & 127;
      const int j = t & 7, prow = t >> 3;
      uint32_t ac = 0;
      long long r = r_begin;
      while (r < r_end) {
        const int b = (int)(r / p.H);
        const int y0 = (int)(r - (long long)b * p.H);
        const int R = (int)((r_end - r) < (long long)(p.H - y0) ? (r_end - r) : (long long)(p.H - y0));
        float ca[8], cb[8];
        {
          const float4* cf = reinterpret_cast<const float4*>(p.coef + (long long)b * 128 + j * 8);
          const float4 a0 = __ldg(cf), a1 = __ldg(cf + 1), b0 = __ldg(cf + 16), b1 = __ldg(cf + 17);
          ca[0] = a0.x; ca[1] = a0.y; ca[2] = a0.z; ca[3] = a0.w; ca[4] = a1.x; ca[5] = a1.y; ca[6] = a1.z; ca[7] = a1.w;
          cb[0] = b0.x; cb[1] = b0.y; cb[2] = b0.z; cb[3] = b0.w; cb[4] = b1.x; cb[5] = b1.y; cb[6] = b1.z; cb[7] = b1.w;
          if (p.act) {     // silu(u) = h (1 + tanh h), h = u / 2: the coefficients are halved once
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              ca[e] *= 0.5f;
              cb[e] *= 0.5f;
            }
          }
        }
        for (int jr = 0; jr < R; ++jr, ++ac) {
          if ((int)(ac & 1u) != xset) continue;
          const uint32_t as = ac % (uint32_t)SA, aph = (ac / (uint32_t)SA) & 1u;
          mbar_wait(&a_full[as], aph, p.err, 0x4600 + as);
          const uint32_t base = smem_u32(a_smem) + as * (uint32_t)p.a_slot_bytes;
          for (int px0 = 1 + prow; px0 <= p.W; px0 += 64) {
            uint4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int px = px0 + 16 * i;
              if (px <= p.W) v[i] = lds128(base + px * 128 + ((j ^ (px & 7)) << 4));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int px = px0 + 16 * i;
              if (px <= p.W) {
                const uint32_t in[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float x0, x1;
                  if (p.fmt) {
                    const float2 f = unpack_f16x2(in[e]);
                    x0 = f.x;
                    x1 = f.y;
                  } else {
                    x0 = bf16_lo(in[e]);
                    x1 = bf16_hi(in[e]);
                  }
                  float y0v = fmaf(x0, ca[2 * e], cb[2 * e]), y1v = fmaf(x1, ca[2 * e + 1], cb[2 * e + 1]);
                  if (p.act) {
                    y0v = silu_from_half_arg(y0v);
                    y1v = silu_from_half_arg(y1v);
                  }
                  o[e] = pack_op2(y0v, y1v, p.fmt);
                }
                sts128(base + px * 128 + ((j ^ (px & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
              }
            }
          }
          fence_proxy_async_smem();
          mbar_arrive_warp(&a_ready[as]);
        }
        r += R;
      }
    }
    // ======================================= epilogue =======================================
    const int q = warp & 3;
    const int m = q * 32 + lane;            // accumulator row: co = m & 63, upper half = second tap of the pair
    const int co = m & 63;
    mbar_wait(acc_full, 0, p.err, 0x4500);
    tc_fence_after();
    float* base = p.partial + (long long)blockIdx.x * p.taps * 4096;
    const bool has_work = r_end > r_begin;
    const int eset = (warp - 2) >> 2;          // epilogue set: accumulators of pair A (kx 0..2) / pair B
    for (int i = 3 * eset; i < 3 * eset + 3; ++i) {
      const int pair = i / 3, kx = i - pair * 3;
      int tap;
      if (p.taps == 9) {
        const int ky = (pair == 0) ? (m < 64 ? 1 : 0) : (m < 64 ? 2 : -1);
        tap = ky < 0 ? -1 : ky * 3 + kx;
      } else {
        tap = (i == 1 && m < 64) ? 0 : -1;
      }
      if (p.taps == 1 && i != 1) continue;       // warp-uniform
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + i * 64 + c * 32, v);
        tmem_wait_ld();
        if (tap >= 0 && !(p.dbg & 1)) {
          float4* dst = reinterpret_cast<float4*>(base + ((long long)tap * 64 + co) * 64 + c * 32);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 o;
            o.x = has_work ? __uint_as_float(v[4 * u + 0]) : 0.f;
            o.y = has_work ? __uint_as_float(v[4 * u + 1]) : 0.f;
            o.z = has_work ? __uint_as_float(v[4 * u + 2]) : 0.f;
            o.w = has_work ? __uint_as_float(v[4 * u + 3]) : 0.f;
            dst[u] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dw[(co*co_mul + co_add)][ci_off + ci][tap] (+)= sum_cta partial[cta][tap][co][ci]   (fp64 ordered sum)
// dw is the reference weight layout [Cout][cin_total][k][k] flattened, k*k = taps.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int n_ctas, int taps, float* __restrict__ dw, int cin_total,
                    int ci_off, int co_mul, int co_add, int co_count, int ci_count, int accumulate) {
  const int idx = blockIdx.x * 256 + threadIdx.x;        // (tap, co, ci), ci fastest
  if (idx >= taps * 4096) return;
  const int ci = idx & 63, co = (idx >> 6) & 63, tap = idx >> 12;
  if (co >= co_count || ci >= ci_count) return;
  double t = 0.0;
  int c = 0;
  for (; c + 4 <= n_ctas; c += 4) {
    const float v0 = partial[(long long)(c + 0) * taps * 4096 + idx];
    const float v1 = partial[(long long)(c + 1) * taps * 4096 + idx];
    const float v2 = partial[(long long)(c + 2) * taps * 4096 + idx];
    const float v3 = partial[(long long)(c + 3) * taps * 4096 + idx];
    t += (double)v0;
    t += (double)v1;
    t += (double)v2;
    t += (double)v3;
  }
  for (; c < n_ctas; ++c) t += (double)partial[(long long)c * taps * 4096 + idx];
  float* o = dw + ((long long)(co * co_mul + co_add) * cin_total + ci_off + ci) * taps + tap;
  *o = accumulate ? *o + (float)t : (float)t;
}

// Every weight-gradient fold of a training step in ONE launch (grid.y = job): 33 folds of 148 partials each are
// latency-bound (a dependent fp64 chain per thread), ~15 us per launch on their own.
__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(const mcedm_wgrad_job* __restrict__ jobs) {
  const mcedm_wgrad_job jb = jobs[blockIdx.y];
  const int idx = blockIdx.x * 256 + threadIdx.x;        // (tap, co, ci), ci fastest
  if (idx >= jb.taps * 4096) return;
  const int ci = idx & 63, co = (idx >> 6) & 63, tap = idx >> 12;
  if (co >= jb.co_count || ci >= jb.ci_count) return;
  const long long step = (long long)jb.taps * 4096;
  const float* p = jb.partial + idx;
  double t = 0.0;
  int c = 0;
  for (; c + 8 <= jb.n_ctas; c += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[(long long)(c + k) * step];
#pragma unroll
    for (int k = 0; k < 8; ++k) t += (double)v[k];
  }
  for (; c < jb.n_ctas; ++c) t += (double)p[(long long)c * step];
  float* o = jb.dw + ((long long)(co * jb.co_mul + jb.co_add) * jb.cin_total + jb.ci_off + ci) * jb.taps + tap;
  *o = (float)t;
}

static int wgrad_grid(int B, int H, int W) {
  // pixels per CTA >= 256 (was 512: at 32x32 and 32 samples only 64 of the 148 SMs had work)
  static int div = 0;
  if (!div) {
    const char* e = getenv("MCEDM_WG_DIV");
    div = e ? atoi(e) : 256;
    if (div < 128) div = 128;
  }
  long long g = (long long)B * H * W / div;
  if (g < 1) g = 1;
  if (g > num_sms()) g = num_sms();
  if (g > (long long)B * H) g = (long long)B * H;
  return (int)g;
}

static int make_pix_tmap(CUtensorMap* tm, const void* ptr, int layout, int c_total, int B, int H, int W, int box_w,
                         int* row_off) {
  if (layout == 0) {
    *row_off = 0;
    return make_tmap_pix_bf16(tm, ptr, c_total, W, H, B, (long long)H * W, box_w);
  }
  int P = 0, blk = 0;
  int rc = mcedm_flat_geometry(H, W, &P, &blk);
  if (rc) return rc;
  *row_off = 1;
  return make_tmap_pix_bf16(tm, ptr, c_total, P, H + 2, B, blk, box_w);
}

}  // namespace mcedm

extern "C" int mcedm_wgrad_ctas(int B, int H, int W) { return mcedm::wgrad_grid(B, H, W); }

extern "C" int mcedm_conv_wgrad(const void* dy, int dy_layout, int dy_ctotal, int dy_coff, const void* a, int a_layout,
                                int a_ctotal, int a_coff, int B, int H, int W, int taps, float* partial,
                                void* stream) {
  return mcedm_conv_wgrad16(dy, dy_layout, dy_ctotal, dy_coff, a, a_layout, a_ctotal, a_coff, B, H, W, taps, partial, 0,
                            stream);
}

extern "C" int mcedm_conv_wgrad16(const void* dy, int dy_layout, int dy_ctotal, int dy_coff, const void* a, int a_layout,
                                  int a_ctotal, int a_coff, int B, int H, int W, int taps, float* partial, int op_fmt,
                                  void* stream) {
  return mcedm_conv_wgrad16_fused(dy, dy_layout, dy_ctotal, dy_coff, a, a_layout, a_ctotal, a_coff, nullptr, 0, B, H, W,
                                  taps, partial, op_fmt, stream);
}

extern "C" int mcedm_conv_wgrad16_fused(const void* dy, int dy_layout, int dy_ctotal, int dy_coff, const void* a,
                                        int a_layout, int a_ctotal, int a_coff, const float* a_coef, int a_act, int B,
                                        int H, int W, int taps, float* partial, int op_fmt, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(a_coef == nullptr || a_ctotal == 64, "conv_wgrad: the in-kernel transform takes a 64-channel tensor");
  MCEDM_REQUIRE(B >= 1 && H >= 1 && W >= 16 && W <= 128 && W % 16 == 0, "conv_wgrad: unsupported W=%d", W);
  MCEDM_REQUIRE(taps == 9 || taps == 1, "conv_wgrad: taps=%d (9 or 1)", taps);
  MCEDM_REQUIRE(dy_ctotal % 64 == 0 && a_ctotal % 64 == 0 && dy_coff % 64 == 0 && a_coff % 64 == 0 &&
                    dy_coff < dy_ctotal && a_coff < a_ctotal,
                "conv_wgrad: bad channel block");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.H = H;
  p.W = W;
  p.total_rows = (long long)B * H;
  p.taps = taps;
  p.dy_coff = dy_coff;
  p.a_coff = a_coff;
  p.dy_slot_bytes = W * 128;
  p.a_slot_bytes = ((W + 2) * 128 + 1023) / 1024 * 1024;
  // ring depths: ~64 KB of dy rows and ~48 KB of a rows in flight (narrow levels have 2-4 KB rows: the
  // producer has to run many rows ahead of the MMA issuer to hide the TMA round trip)
  // (with 4 dy slots - 3 of them the MMA window - and 3 a slots at W = 128 the kernel ran at one global-memory latency per
  // row: 60 of its 85 us at 32 samples were spent with the MMAs AND the epilogue switched off.  MCEDM_WG_SLOTS="S,SA")
  p.n_slots = 98304 / p.dy_slot_bytes;
  if (p.n_slots < 4) p.n_slots = 4;
  if (p.n_slots > 16) p.n_slots = 16;
  p.n_aslots = 90112 / p.a_slot_bytes;
  if (p.n_aslots < 3) p.n_aslots = 3;
  if (p.n_aslots > 12) p.n_aslots = 12;
  if (const char* e = getenv("MCEDM_WG_SLOTS")) {
    int s_ = 0, sa_ = 0;
    if (sscanf(e, "%d,%d", &s_, &sa_) == 2 && s_ >= 4 && sa_ >= 3) {
      p.n_slots = s_;
      p.n_aslots = sa_;
    }
  }
  p.partial = partial;
  p.idesc = umma_idesc_16(128, 64, 1, 1, op_fmt ? 1 : 0);
  p.idesc192 = umma_idesc_16(128, 192, 1, 1, op_fmt ? 1 : 0);
  p.coef = a_coef;
  p.act = a_act ? 1 : 0;
  p.fmt = op_fmt ? 1 : 0;
  p.err = watchdog_ptr();
  if (const char* e = getenv("MCEDM_WG_DBG")) p.dbg = atoi(e);
  MCEDM_REQUIRE(p.err != nullptr, "conv_wgrad: cannot allocate the watchdog word");
  CUtensorMap tm_dy, tm_a;
  int rc = make_pix_tmap(&tm_dy, dy, dy_layout, dy_ctotal, B, H, W, W, &p.dy_row_off);
  if (rc) return rc;
  rc = make_pix_tmap(&tm_a, a, a_layout, a_ctotal, B, H, W, W + 2, &p.a_row_off);
  if (rc) return rc;
  const int smem = 1024 + (p.n_slots + 2) * p.dy_slot_bytes + p.n_aslots * p.a_slot_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MCEDM_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int grid = wgrad_grid(B, H, W);
  conv_wgrad_kernel<<<grid, 352, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tm_dy, tm_a, p);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_wgrad_reduce(const float* partial, int n_ctas, int taps, float* dw, int cin_total, int ci_off,
                                  int co_mul, int co_add, int co_count, int ci_count, int accumulate, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(n_ctas >= 1 && (taps == 9 || taps == 1), "wgrad_reduce: bad sizes");
  MCEDM_REQUIRE(co_count >= 1 && co_count <= 64 && ci_count >= 1 && ci_count <= 64, "wgrad_reduce: bad channel counts");
  const int n = taps * 4096;
  wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, n_ctas, taps, dw, cin_total, ci_off, co_mul, co_add, co_count, ci_count, accumulate);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mcedm_wgrad_reduce_batched(const mcedm_wgrad_job* jobs_dev, int n_jobs, void* stream) {
  using namespace mcedm;
  MCEDM_REQUIRE(n_jobs >= 1 && n_jobs <= 65535, "wgrad_reduce_batched: bad sizes");
  dim3 grid(9 * 4096 / 256, n_jobs);
  wgrad_reduce_batched_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(jobs_dev);
  MCEDM_CUDA(cudaGetLastError());
  return 0;
}
