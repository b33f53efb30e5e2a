// frameio.cu — the step right after the hot path in both callers: label map -> colour image,
// `CITYSCAPE_PALETTE[pred]` (semantic_seg.py:52-72, :101-112; seg_video.py:168), optionally alpha-blended over the
// input frame (seg_video.py:200-203 draws the colour image with alpha=0.6 over the frame).  HBM-bound byte work:
// 1 B read + 3 B written per pixel (+3 B read with the overlay).
#include "common.cuh"

namespace drnb200 {

// thread = 4 adjacent pixels: one 32-bit label load, three 32-bit stores (12 bytes RGBRGBRGBRGB)
template <bool BLEND>
__global__ void __launch_bounds__(256) palette_kernel(const uint8_t* __restrict__ labels, int64_t n_quads,
                                                      const uint8_t* __restrict__ palette, int n_colors,
                                                      const uint8_t* __restrict__ frames, float alpha,
                                                      uint8_t* __restrict__ out) {
  __shared__ uint8_t pal[256 * 3];
  for (int i = threadIdx.x; i < 256 * 3; i += blockDim.x) {
    const int c = i / 3;
    // labels >= n_colors (e.g. the ignore label 255) take the LAST palette row (black in CITYSCAPE_PALETTE)
    pal[i] = palette[(c < n_colors ? c : n_colors - 1) * 3 + (i - c * 3)];
  }
  __syncthreads();
  const float beta = __fsub_rn(1.0f, alpha);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads;
       q += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t l4 = __ldg(reinterpret_cast<const uint32_t*>(labels) + q);
    uint8_t b[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t l = (l4 >> (8 * j)) & 0xFFu;
      b[3 * j] = pal[3 * l]; b[3 * j + 1] = pal[3 * l + 1]; b[3 * j + 2] = pal[3 * l + 2];
    }
    if (BLEND) {
      const uint32_t* fp = reinterpret_cast<const uint32_t*>(frames) + 3 * q;
      const uint32_t f[3] = {__ldg(fp), __ldg(fp + 1), __ldg(fp + 2)};
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const float fv = (float)((f[j >> 2] >> (8 * (j & 3))) & 0xFFu);
        // round-half-even of alpha*colour + (1-alpha)*frame, every step a separately rounded fp32 operation
        b[j] = (uint8_t)__float2uint_rn(__fadd_rn(__fmul_rn(alpha, (float)b[j]), __fmul_rn(beta, fv)));
      }
    }
    uint32_t* op = reinterpret_cast<uint32_t*>(out) + 3 * q;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      op[k] = (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) |
              ((uint32_t)b[4 * k + 3] << 24);
  }
}

}  // namespace drnb200

extern "C" int drnb200_colorize(const uint8_t* labels, int64_t n_px, const uint8_t* palette, int n_colors,
                                const uint8_t* frames_or_null, float alpha, uint8_t* out_rgb, void* stream) {
  DRN_REQUIRE(labels && palette && out_rgb, "colorize: null pointer");
  DRN_REQUIRE(n_px >= 0 && n_px % 4 == 0, "colorize: the pixel count must be a multiple of 4 (got %lld)", (long long)n_px);
  DRN_REQUIRE(n_colors >= 1 && n_colors <= 256, "colorize: bad palette size %d", n_colors);
  DRN_REQUIRE(alpha >= 0.f && alpha <= 1.f, "colorize: alpha must be in [0,1]");
  if (n_px == 0) return DRNB200_OK;
  const int64_t quads = n_px / 4;
  int64_t blocks = (quads + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (frames_or_null)
    drnb200::palette_kernel<true><<<(int)blocks, 256, 0, st>>>(labels, quads, palette, n_colors, frames_or_null, alpha, out_rgb);
  else
    drnb200::palette_kernel<false><<<(int)blocks, 256, 0, st>>>(labels, quads, palette, n_colors, nullptr, alpha, out_rgb);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}
