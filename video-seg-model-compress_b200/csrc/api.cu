// api.cu — error plumbing and the small utility entry points of libdrnb200.so.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdarg.h>
#include <string.h>

namespace drnb200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return DRNB200_E_CUDA;
}

__global__ void labels_to_i64_kernel(const uint8_t* __restrict__ in, int64_t n,
                                     int64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (int64_t)in[i];
}

}  // namespace drnb200

extern "C" {

int drnb200_version(void) { return DRNB200_VERSION; }

const char* drnb200_last_error(void) { return drnb200::g_err; }

int drnb200_labels_to_i64(const uint8_t* labels, int64_t n, int64_t* out, void* stream) {
  DRN_REQUIRE(labels && out && n >= 0, "labels_to_i64: null pointer or negative size");
  if (n == 0) return DRNB200_OK;
  int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  drnb200::labels_to_i64_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(labels, n, out);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

int drnb200_ingest_lut(const float* mean, const float* std, int act_dtype, uint16_t* lut) {
  DRN_REQUIRE(mean && std && lut, "ingest_lut: null pointer");
  DRN_REQUIRE(act_dtype == DRNB200_BF16 || act_dtype == DRNB200_F16, "ingest_lut: bad act_dtype");
  for (int c = 0; c < 3; ++c) {
    DRN_REQUIRE(std[c] != 0.f, "ingest_lut: std[%d] is zero", c);
    for (int b = 0; b < 256; ++b) {
      // ToTensorVideoImage: img.float().div(255) (data_transforms.py:277); Normalize: t.sub_(m).div_(s)
      // (data_transforms.py:119-120) -- three correctly rounded fp32 operations, then ONE rounding to act_dtype
      volatile float v = (float)b / 255.0f;
      v = v - mean[c];
      v = v / std[c];
      if (act_dtype == DRNB200_F16) {
        const __half h = __float2half_rn(v);
        lut[c * 256 + b] = *reinterpret_cast<const uint16_t*>(&h);
      } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        lut[c * 256 + b] = *reinterpret_cast<const uint16_t*>(&h);
      }
    }
  }
  return DRNB200_OK;
}

}  // extern "C"
