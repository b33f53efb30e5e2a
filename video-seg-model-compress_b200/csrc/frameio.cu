// frameio.cu — the step right after the hot path in both callers: label map -> colour image,
// `CITYSCAPE_PALETTE[pred]` (semantic_seg.py:52-72, :101-112; seg_video.py:168), optionally alpha-blended over the
// input frame (seg_video.py:200-203 draws the colour image with alpha=0.6 over the frame).  HBM-bound byte work:
// 1 B read + 3 B written per pixel (+3 B read with the overlay).
#include "common.cuh"
#include <algorithm>

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

// ------------------------------------------------------------------------------------------------ frame resize
// The step right BEFORE the ingest in the video caller: seg_video_old.py:125-128 resizes every decoded frame with
// torchvision T.Resize on a PIL image = Pillow's 8-bit BILINEAR resampler (src/libImaging/Resample.c):
//   per pass  ss = 1 << 21;  ss += pixel * k_int[t]  over the taps in order;  out = clip8(ss >> 22)
// with k_int = (int)(k * 2^22 +- 0.5) of the same double coefficients the multi-scale path uses (host tables).
// Horizontal pass into a uint8 temporary, then the vertical pass, exactly like ImagingResampleInner.
// One pass kernel, axis-agnostic: `out[n][o][i][c] = clip8(sum_t in[n][i or o ...])`; thread = one output byte triple
// position (pixel), the three channels of a pixel together.  HBM-bound byte work of a few MB per frame.
constexpr int kResizeBits = 32 - 8 - 2;

// HORIZ: out [N][L][O][3] from in [N][L][I][3]   (L = rows, axis of length I -> O is the pixel axis of a row)
// else:  out [N][O][L][3] from in [N][I][L][3]   (axis of length I -> O is the row axis)
template <bool HORIZ>
__global__ void __launch_bounds__(256) resize_u8_pass_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                             int N, int L, int I, int O, const int32_t* __restrict__ kmin,
                                                             const int32_t* __restrict__ kcnt,
                                                             const int32_t* __restrict__ kk, int ksize) {
  const int64_t total = (int64_t)N * L * O;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int n, l, o;
    if (HORIZ) { o = (int)(idx % O); l = (int)((idx / O) % L); n = (int)(idx / ((int64_t)O * L)); }
    else       { l = (int)(idx % L); o = (int)((idx / L) % O); n = (int)(idx / ((int64_t)O * L)); }
    const int lo = __ldg(kmin + o), cnt = __ldg(kcnt + o);
    const int32_t* k = kk + (int64_t)o * ksize;
    int s0 = 1 << (kResizeBits - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < cnt; ++t) {
      const int i = lo + t;
      const uint8_t* px = HORIZ ? in + (((int64_t)n * L + l) * I + i) * 3 : in + (((int64_t)n * I + i) * L + l) * 3;
      const int w = __ldg(k + t);
      s0 += (int)__ldg(px) * w; s1 += (int)__ldg(px + 1) * w; s2 += (int)__ldg(px + 2) * w;
    }
    uint8_t* op = HORIZ ? out + (((int64_t)n * L + l) * O + o) * 3 : out + (((int64_t)n * O + o) * L + l) * 3;
    op[0] = (uint8_t)min(max(s0 >> kResizeBits, 0), 255);
    op[1] = (uint8_t)min(max(s1 >> kResizeBits, 0), 255);
    op[2] = (uint8_t)min(max(s2 >> kResizeBits, 0), 255);
  }
}

}  // namespace drnb200

extern "C" int drnb200_resize_u8(const uint8_t* src, int N, int Hs, int Ws, uint8_t* dst, int H, int W,
                                 const int32_t* xmin, const int32_t* xcnt, const int32_t* xk, int kx,
                                 const int32_t* ymin, const int32_t* ycnt, const int32_t* yk, int ky,
                                 uint8_t* tmp, void* stream) {
  DRN_REQUIRE(src && dst && N >= 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "resize_u8: null pointer or bad shape");
  DRN_REQUIRE(kx > 0 || Ws == W, "resize_u8: kx == 0 (no horizontal pass) needs Ws == W");
  DRN_REQUIRE(ky > 0 || Hs == H, "resize_u8: ky == 0 (no vertical pass) needs Hs == H");
  DRN_REQUIRE(kx == 0 || (xmin && xcnt && xk), "resize_u8: horizontal tables missing");
  DRN_REQUIRE(ky == 0 || (ymin && ycnt && yk), "resize_u8: vertical tables missing");
  DRN_REQUIRE(!(kx > 0 && ky > 0) || tmp, "resize_u8: a two-pass resize needs the [N, Hs, W, 3] scratch");
  if (N == 0) return DRNB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  auto grid = [](int64_t total) { return (int)std::min<int64_t>((total + 255) / 256, 148 * 16); };
  if (kx == 0 && ky == 0) {
    DRN_CUDA(cudaMemcpyAsync(dst, src, (size_t)N * H * W * 3, cudaMemcpyDeviceToDevice, st));
    return DRNB200_OK;
  }
  const uint8_t* vin = src;
  if (kx > 0) {
    uint8_t* hout = ky > 0 ? tmp : dst;
    drnb200::resize_u8_pass_kernel<true><<<grid((int64_t)N * Hs * W), 256, 0, st>>>(src, hout, N, Hs, Ws, W, xmin, xcnt, xk, kx);
    DRN_CUDA(cudaGetLastError());
    vin = hout;
  }
  if (ky > 0) {
    drnb200::resize_u8_pass_kernel<false><<<grid((int64_t)N * H * W), 256, 0, st>>>(vin, dst, N, W, Hs, H, ymin, ycnt, yk, ky);
    DRN_CUDA(cudaGetLastError());
  }
  return DRNB200_OK;
}

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
