// hist.cu — confusion-matrix accumulation on the device: fast_hist (semantic_seg.py:293-296).
//   k = (label >= 0) & (label < n);  hist[n*label[k] + pred[k]] += 1   (rows = label, cols = pred)
// HBM-bound: reads 1 byte of prediction + 1 (or 8) bytes of label per pixel.  Counters are kept in
// shared memory per CTA (n*n <= 1024 32-bit counters) and flushed once with 64-bit global atomics.
#include "common.cuh"

namespace drnb200 {

template <bool LABEL_I64>
__global__ void __launch_bounds__(256)
confusion_kernel(const uint8_t* __restrict__ pred, const void* __restrict__ label, int64_t n_px,
                 int classes, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_hist[1024];
  const int nn = classes * classes;
  for (int i = threadIdx.x; i < nn; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += stride) {
    long long l = LABEL_I64 ? reinterpret_cast<const long long*>(label)[i]
                            : (long long)reinterpret_cast<const uint8_t*>(label)[i];
    const int p = pred[i];
    if (l >= 0 && l < classes && p < classes) atomicAdd(&s_hist[(int)l * classes + p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nn; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

}  // namespace drnb200

using namespace drnb200;

extern "C" int drnb200_confusion(const uint8_t* pred, const void* label, int label_is_i64,
                                 int64_t n_px, int classes, int64_t* hist, void* stream) {
  DRN_REQUIRE(pred && label && hist, "confusion: null pointer");
  DRN_REQUIRE(classes > 0 && classes <= 32, "confusion: classes must be in [1,32] (got %d)", classes);
  DRN_REQUIRE(n_px >= 0, "confusion: negative pixel count");
  if (n_px == 0) return DRNB200_OK;
  // each CTA handles <= 2^24 pixels so the 32-bit shared counters cannot overflow
  int64_t blocks = (n_px + 256 * 64 - 1) / (256 * 64);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (n_px / blocks > (1 << 24) * 256ll) blocks = n_px / ((1 << 24) * 256ll) + 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (label_is_i64)
    confusion_kernel<true><<<(int)blocks, 256, 0, st>>>(pred, label, n_px, classes,
                                                       reinterpret_cast<unsigned long long*>(hist));
  else
    confusion_kernel<false><<<(int)blocks, 256, 0, st>>>(pred, label, n_px, classes,
                                                        reinterpret_cast<unsigned long long*>(hist));
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}
