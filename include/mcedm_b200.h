/*
 * mcedm_b200 — C ABI of the B200-native m-cedm hot path (libmcedm_b200.so).
 *
 * The reference (katehai/m-cedm) is pure Python/PyTorch: it has no FFI, and every numerical step of
 * its hot path is a library call made from models/adm_blocks.py, models/mcedm.py and
 * models/losses.py.  Each entry point below replaces one such call site (cited as file:line,
 * relative to the reference root) and is what a maintainer binds with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named host_*;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream;
 *   - return 0 on success; non-zero on failure with a message in mcedm_last_error();
 *   - activations inside the U-Net are NHWC ("channels last"), 64 channels per tensor:
 *       fp32 NHWC  = residual stream / conv outputs,
 *       bf16 NHWC  = tensor-core operands (normalised + activated tensors);
 *   - the public tensors of the reference API (x, cond, mask, D_x, sampler state) stay NCHW,
 *     fp32 (sampler state fp64), exactly as models/mcedm.py passes them.
 *   - there is NO CPU implementation behind any of these: without an sm_100 device they fail.
 */
#ifndef MCEDM_B200_H
#define MCEDM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MCEDM_ABI_VERSION 1
#if defined(__GNUC__)
#define MCEDM_API __attribute__((visibility("default")))
#else
#define MCEDM_API
#endif

/* -------------------------------------------------------------------------------------------- */
/* runtime                                                                                      */
/* -------------------------------------------------------------------------------------------- */
MCEDM_API int mcedm_abi_version(void);
MCEDM_API const char* mcedm_last_error(void);
/* Synchronises `stream`, returns non-zero if any kernel's bounded mbarrier wait timed out. */
MCEDM_API int mcedm_check_watchdog(void* stream);

/* -------------------------------------------------------------------------------------------- */
/* K1  convolution as implicit GEMM on tcgen05  (models/adm_blocks.py:65-81 Conv2d.forward;       */
/*     fused residual add :171,:179; fused decoder concat :401; fused 1x1 skip conv :150-151)     */
/* -------------------------------------------------------------------------------------------- */
/*
 * out[b,y,x,:N] = bias + sum_s  src[seg_src[s]][b, y+seg_dy[s], x+seg_dx[s], 0:64] . w_packed[s][:N][0:64]
 *                 (+ residual)
 *   src[i]      : n_src (1..4) bf16 NHWC tensors [B,H,W,64]; reads outside the image are zero (padding)
 *   seg_*       : HOST int arrays of length n_seg (<= 20); dy,dx in {-1,0,1}
 *   w_packed    : bf16 [n_seg][N][64]   (see mcedm_b200.packing for the reference-weight permutation)
 *   bias        : fp32 [N] or NULL
 *   N           : 16, 64, 128 or 192 output channels
 *   out         : [B,H,W,N] fp32 (out_bf16 = 0) or bf16 (out_bf16 = 1)
 *   res,res_mode: 0 none | 1 res is fp32 [B,H,W,N] | 2 res is [B,H/2,W/2,N], nearest x2 upsampled
 *                 (Conv2d up, adm_blocks.py:73-74) | 3 res is [B,2H,2W,N], 2x2 mean (down, :75-77)
 *   stats_partial: NULL, or fp32 [B*H*W/128][N/4][2] receiving per-128-pixel-tile (sum, sum of squares)
 *                 of the stored values per 4-channel GroupNorm group (feeds mcedm_gn_apply).
 * Requires W | 128, 128 | H*W.
 */
MCEDM_API int mcedm_conv_igemm(const void* const* src, int n_src, const int* seg_src, const int* seg_dy, const int* seg_dx,
                     int n_seg, const void* w_packed, const float* bias, int B, int H, int W, int N, void* out,
                     int out_bf16, const float* res, int res_mode, float* stats_partial, void* stream);

/* -------------------------------------------------------------------------------------------- */
/* bring-up / checker kernels (tests only; not on the product path)                              */
/* -------------------------------------------------------------------------------------------- */
MCEDM_API int mcedm_probe_umma(const void* a, int a_rows, const void* bm, int row_shift, int base_offset, int b_mn_major,
                     float* out, void* stream);
/* seg_dev: DEVICE int array [n_seg][3] = (src, dy, dx). Same math as mcedm_conv_igemm on CUDA cores. */
MCEDM_API int mcedm_conv_direct_ref(const void* const* src, int n_src, const int* seg_dev, int n_seg, const void* w_packed,
                          const float* bias, int B, int H, int W, int N, float* out, const float* res, int res_mode,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCEDM_B200_H */
