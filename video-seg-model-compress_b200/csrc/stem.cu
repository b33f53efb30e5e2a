// stem.cu — DRN layer0: Conv2d(3,16,7,s1,p3) -> BatchNorm2d(eval) -> ReLU  (drn.py:132-137).
// Input is the float32 NCHW frame batch exactly as the reference callers pass it to DRNSeg.forward
// (semantic_seg.py:440-444); output is the NHWC 16-bit activation the tcgen05 convs consume, so the
// layout change and the down-conversion are fused into this kernel.
// CUDA-core direct kernel: each thread produces 4 horizontally adjacent pixels x 16 channels; a row of
// 10 input values per (ci, ky) is reused across the 7 kx taps of the 4 pixels (448 FMA per 38 smem loads).
#include "common.cuh"

namespace drnb200 {

constexpr int ST_TW = 64;   // output tile width  (16 threads x 4 pixels)
constexpr int ST_TH = 8;    // output tile height
constexpr int ST_IW = ST_TW + 6;
constexpr int ST_IH = ST_TH + 6;
constexpr int ST_C0 = 16;

template <int DT>
__global__ void __launch_bounds__(128) stem_kernel(const float* __restrict__ x,
                                                   const float* __restrict__ w,
                                                   const float* __restrict__ scale,
                                                   const float* __restrict__ shift, int N, int H,
                                                   int W, uint16_t* __restrict__ y) {
  __shared__ float s_in[3][ST_IH][ST_IW + 2];
  __shared__ __align__(16) float s_w[147][ST_C0];  // [ci*49 + ky*7 + kx][co]
  const int n = blockIdx.z;
  const int ox0 = blockIdx.x * ST_TW, oy0 = blockIdx.y * ST_TH;
  const int tid = threadIdx.x;

  for (int i = tid; i < 147 * ST_C0; i += 128) {
    const int co = i / 147, k = i - co * 147;  // w is OIHW: [co][ci][ky][kx]
    s_w[k][co] = __ldg(w + i);
  }
  for (int i = tid; i < 3 * ST_IH * ST_IW; i += 128) {
    const int ci = i / (ST_IH * ST_IW);
    const int r = (i / ST_IW) % ST_IH, c = i % ST_IW;
    const int iy = oy0 + r - 3, ix = ox0 + c - 3;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = __ldg(x + (((size_t)n * 3 + ci) * H + iy) * W + ix);
    s_in[ci][r][c] = v;
  }
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;  // 16 x 8 threads
  float acc[4][ST_C0];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < ST_C0; ++c) acc[p][c] = 0.f;

  for (int ci = 0; ci < 3; ++ci) {
    for (int ky = 0; ky < 7; ++ky) {
      float xin[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) xin[i] = s_in[ci][ty + ky][tx * 4 + i];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float4* wp = reinterpret_cast<const float4*>(&s_w[ci * 49 + ky * 7 + kx][0]);
        float wv[ST_C0];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = wp[q];
          wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int c = 0; c < ST_C0; ++c) acc[p][c] = fmaf(xin[p + kx], wv[c], acc[p][c]);
      }
    }
  }

  const int oy = oy0 + ty;
  if (oy >= H) return;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int ox = ox0 + tx * 4 + p;
    if (ox >= W) continue;
    uint32_t pk[8];
#pragma unroll
    for (int c = 0; c < ST_C0; c += 2) {
      const float a = fmaxf(fmaf(acc[p][c], __ldg(scale + c), __ldg(shift + c)), 0.f);
      const float b = fmaxf(fmaf(acc[p][c + 1], __ldg(scale + c + 1), __ldg(shift + c + 1)), 0.f);
      pk[c >> 1] = (uint32_t)Act<DT>::from_f32(a) | ((uint32_t)Act<DT>::from_f32(b) << 16);
    }
    uint4* yp = reinterpret_cast<uint4*>(y + (((size_t)n * H + oy) * W + ox) * ST_C0);
    yp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    yp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

}  // namespace drnb200

using namespace drnb200;

extern "C" int drnb200_stem_forward(const float* x_nchw, const float* w_oihw, const float* bn_scale,
                                    const float* bn_shift, int N, int H, int W, int C0,
                                    int act_dtype, void* y_nhwc, void* stream) {
  DRN_REQUIRE(x_nchw && w_oihw && bn_scale && bn_shift && y_nhwc, "stem_forward: null pointer");
  DRN_REQUIRE(C0 == ST_C0, "stem_forward: C0 must be 16 (got %d)", C0);
  DRN_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, "stem_forward: bad shape N=%d H=%d W=%d", N, H, W);
  DRN_REQUIRE(act_dtype == DRNB200_BF16 || act_dtype == DRNB200_F16, "stem_forward: bad act_dtype");
  dim3 grid((W + ST_TW - 1) / ST_TW, (H + ST_TH - 1) / ST_TH, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == DRNB200_BF16)
    stem_kernel<DRNB200_BF16><<<grid, 128, 0, st>>>(x_nchw, w_oihw, bn_scale, bn_shift, N, H, W,
                                                   reinterpret_cast<uint16_t*>(y_nhwc));
  else
    stem_kernel<DRNB200_F16><<<grid, 128, 0, st>>>(x_nchw, w_oihw, bn_scale, bn_shift, N, H, W,
                                                  reinterpret_cast<uint16_t*>(y_nhwc));
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}
