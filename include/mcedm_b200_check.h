/* Checker and bring-up kernels of mcedm_b200 (libmcedm_b200_check.so): test infrastructure, NOT part of the product
 * library - tests/ and scripts/ load it through mcedm_b200._lib.check_lib().  Same conventions as mcedm_b200.h
 * (return 0 on success; device pointers; `stream` = cudaStream_t). */
#ifndef MCEDM_B200_CHECK_H
#define MCEDM_B200_CHECK_H
#include "mcedm_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* tensor-pipe + shared-memory operand-fetch ceiling: every SM issues n_tiles x 36 tcgen05.mma (M=128, N, K=16, the conv
 * kernels' descriptor pattern) and nothing else; cycles_per_cta[sm] = clock64 ticks (DEVICE int64 [#SMs]). */
MCEDM_API int mcedm_probe_mma_rate(int N, int n_tiles, long long* cycles_per_cta, void* stream);
/* issue-queue depth of tcgen05.mma: per iteration n_mma x (M=128, N=192, K=16) + one commit, then `idle` cycles of
 * nothing on the issuing warp; out3_per_cta[sm] = {total, issue, commit} clock64 ticks (DEVICE int64 [#SMs][3]). */
MCEDM_API int mcedm_probe_mma_queue(int iters, int n_mma, int idle, long long* out3_per_cta, void* stream);
/* one M=128 x N=64 x K=64 UMMA whose A descriptor starts row_shift rows (128 B each) into a SWIZZLE_128B tile: pins
 * "a row-shifted view of a swizzled tile is a valid operand" (conv_rows.cu, conv_wgrad.cu) on silicon */
MCEDM_API int mcedm_probe_umma(const void* a, int a_rows, const void* bm, int row_shift, int base_offset, int b_mn_major,
                     float* out, void* stream);
/* seg_dev: DEVICE int array [n_seg][3] = (src, dy, dx). Same math as mcedm_conv_igemm on CUDA cores. */
MCEDM_API int mcedm_conv_direct_ref(const void* const* src, int n_src, const int* seg_dev, int n_seg, const void* w_packed,
                          const float* bias, int B, int H, int W, int N, float* out, const float* res, int res_mode,
                          void* stream);
/* fp32 CUDA-core attention on the same bf16 qkv (out fp32 [B,L,64]) */
MCEDM_API int mcedm_attention_ref(const void* qkv_bf16, int B, int L, float* out_f32, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCEDM_B200_CHECK_H */
