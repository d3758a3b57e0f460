// Host-side runtime shared by the launchers: error reporting, the device watchdog word,
// SM count and TMA tensor-map construction (cuTensorMapEncodeTiled through the runtime's
// driver entry point, so the library does not link libcuda directly).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mcedm {

// Records `msg` as the library's last error (thread-local) and returns `code`.
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define MCEDM_CUDA(x)                                                   \
  do {                                                                  \
    cudaError_t _e = (x);                                               \
    if (_e != cudaSuccess) return ::mcedm::cuda_fail(_e, #x);           \
  } while (0)
#define MCEDM_REQUIRE(cond, ...)                                        \
  do {                                                                  \
    if (!(cond)) return ::mcedm::fail(-1, __VA_ARGS__);                 \
  } while (0)

// Device word the bounded mbarrier waits write on timeout (see ptx.cuh). Allocated on first use.
unsigned int* watchdog_ptr();
int num_sms();

// NHWC bf16 activation tensor [B, H, W, C] viewed by TMA as (C, W, H, B), box (64, box_w, box_h, 1),
// SWIZZLE_128B, zero fill outside the tensor (this is what implements the conv's zero padding).
int make_tmap_nhwc_bf16(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int box_w, int box_h);
// Pixel-sequence view used by the backward kernels: bf16 [B][rows][pitch][c_total] with images img_stride
// pixels apart (dense NHWC: pitch = W, rows = H, img_stride = H*W; padded-flat: pitch = P, rows = H + 2,
// img_stride = blk).  TMA dims (c_total, pitch, rows, B), box (64, box_w, 1, 1), SWIZZLE_128B, zero fill.
int make_tmap_pix_bf16(CUtensorMap* out, const void* ptr, int c_total, int pitch, int rows, int B,
                       long long img_stride, int box_w);
// Row-major bf16 matrix [rows, 64] (packed weights), box (64, box_rows), SWIZZLE_128B.
int make_tmap_rows64_bf16(CUtensorMap* out, const void* ptr, long long rows, int box_rows);

// Programmatic dependent launch: the kernel may become resident before its predecessor in the stream has finished and
// MUST execute pdl_wait() (ptx.cuh) in every CTA before it reads anything an earlier kernel wrote.  Captured into CUDA
// graphs as programmatic edges.  OFF by default (MCEDM_PDL=1 enables it): on the graph-replayed U-Net evaluation the
// programmatic edges measured 1 % SLOWER at 256 samples and 8 % slower at 32 (1.38 -> 1.49 ms, ~1.1 us per launch): the
// persistent convolution CTAs fill an SM's shared memory, so a dependent CTA cannot become resident before its
// predecessor's CTA on that SM has exited anyway, and the plain kernel->kernel edge of a graph is the cheaper one.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mcedm
