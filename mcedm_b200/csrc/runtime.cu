// Runtime support for libmcedm_b200.so (see runtime.cuh). No torch, no libcuda link dependency.
#include <cstdlib>
#include "runtime.cuh"
#include "../../include/mcedm_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace mcedm {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

static std::mutex g_mu;
static unsigned int* g_watchdog[64] = {nullptr};
static int g_sms[64] = {0};

unsigned int* watchdog_ptr() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_watchdog[dev]) {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, 256) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, 256);
    g_watchdog[dev] = p;
  }
  return g_watchdog[dev];
}

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!g_sms[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sms[dev] = n;
  }
  return g_sms[dev];
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MCEDM_PDL");
    on = (e && e[0] == '1') ? 1 : 0;      // default OFF: measured slower (see runtime.cuh)
  }
  return on != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_nhwc_bf16(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int box_w, int box_h) {
  EncodeTiledFn fn = encode_fn();
  MCEDM_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MCEDM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA source must be 16-byte aligned");
  MCEDM_REQUIRE(C % 64 == 0 && box_w >= 1 && box_w <= 256 && box_h >= 1 && box_h <= 256,
                "bad NHWC tensor-map box (C=%d box_w=%d box_h=%d)", C, box_w, box_h);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MCEDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(NHWC %dx%dx%dx%d) failed: %d", B, H, W, C, (int)r);
  return 0;
}

int make_tmap_pix_bf16(CUtensorMap* out, const void* ptr, int c_total, int pitch, int rows, int B,
                       long long img_stride, int box_w) {
  EncodeTiledFn fn = encode_fn();
  MCEDM_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MCEDM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA source must be 16-byte aligned");
  MCEDM_REQUIRE(c_total % 64 == 0 && box_w >= 1 && box_w <= 256, "bad pixel tensor-map (C=%d box_w=%d)", c_total, box_w);
  cuuint64_t dims[4] = {(cuuint64_t)c_total, (cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)c_total * 2, (cuuint64_t)pitch * c_total * 2, (cuuint64_t)img_stride * c_total * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MCEDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(pix %dx%dx%dx%d) failed: %d", B, rows, pitch, c_total, (int)r);
  return 0;
}

int make_tmap_rows64_bf16(CUtensorMap* out, const void* ptr, long long rows, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  MCEDM_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MCEDM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA source must be 16-byte aligned");
  MCEDM_REQUIRE(box_rows >= 1 && box_rows <= 256, "bad weight tensor-map box rows %d", box_rows);
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MCEDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rows=%lld) failed: %d", rows, (int)r);
  return 0;
}

}  // namespace mcedm

extern "C" {

const char* mcedm_last_error(void) { return mcedm::g_err; }

int mcedm_abi_version(void) { return MCEDM_ABI_VERSION; }

int mcedm_check_watchdog(void* stream) {
  unsigned int* p = mcedm::watchdog_ptr();
  if (!p) return mcedm::fail(-2, "watchdog word unavailable");
  unsigned int h = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MCEDM_CUDA(cudaMemcpyAsync(&h, p, sizeof(h), cudaMemcpyDeviceToHost, s));
  MCEDM_CUDA(cudaStreamSynchronize(s));
  if (h != 0) {
    cudaMemsetAsync(p, 0, sizeof(h), s);
    cudaStreamSynchronize(s);
    return mcedm::fail(-3, "device watchdog fired: tag 0x%08x (an mbarrier wait timed out)", h);
  }
  return 0;
}

// Saturation audit (MCEDM_DBG=4, see ptx.cuh sat_audit): number of activation values beyond fp16's range that the fused
// 16-bit convolution epilogues clamped since the last reset.
int mcedm_saturation_count(long long* host_out, int reset, void* stream) {
  unsigned int* p = mcedm::watchdog_ptr();
  if (!p) return mcedm::fail(-2, "watchdog word unavailable");
  unsigned int h = 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MCEDM_CUDA(cudaMemcpyAsync(&h, p + 8, sizeof(h), cudaMemcpyDeviceToHost, s));
  MCEDM_CUDA(cudaStreamSynchronize(s));
  if (host_out) *host_out = (long long)h;
  if (reset) {
    MCEDM_CUDA(cudaMemsetAsync(p + 8, 0, sizeof(h), s));
    MCEDM_CUDA(cudaStreamSynchronize(s));
  }
  return 0;
}

}  // extern "C"
