// head.cu — (c) the segmentation head: seg 1x1 classifier -> fixed-bilinear x8 transposed conv ->
// [log-softmax] -> argmax.   Reference: semantic_seg.py:137-158 (seg, up, softmax), :115-124
// (fill_up_weights), :445 (torch.max(final, 1)).
//
// Round-1 structure: the classifier GEMM runs on the tcgen05 conv kernel (MODE_P: M = 128 pixels,
// N = 32 = classes padded, K = C) into a small float32 NHWC scratch [N,h,w,32] that stays L2-resident
// (4 MB per 1024x2048 frame); up_argmax_kernel then reads it and writes the label map.  The 19 x 1024 x 2048
// float32 logits (159 MB/frame) the reference materialises three times are never written on the fast path.
//
// Transposed-conv rule (k=16, s=8, p=4, depthwise, zero padded):
//   out[y] = sum_i in[i] * wk[y + 4 - 8 i],   wk[k] = 1 - |2k - 15| / 16,  0 <= k < 16
//   => exactly two candidate rows: i0 = (y+4)>>3 with k0 = (y+4)&7, and i0-1 with k0+8; rows outside
//      [0,h) are dropped without renormalisation (this is what differs from F.interpolate at borders).
#include "conv_internal.cuh"
#include <new>

struct drnb200_head_plan {
  int N, h, w, C, classes, act_dtype;
  drnb200_conv_plan* conv;
  int32_t* d_row_ptr;
  int32_t* d_kblk;
  uint16_t* d_wpacked;
  float* d_wpad;     // [32, C] fp32 staging of the zero-padded classifier
  float* d_scale;    // [32] ones
  float* d_shift;    // [32] bias, zero padded
  float* d_logits;   // [N,h,w,32] fp32 scratch
};

namespace drnb200 {

constexpr int HEAD_CP = 32;  // classes padded to the UMMA N granule used by MODE_P

__device__ __forceinline__ float up_w(int k) {  // fill_up_weights, f = 8, c = 15/16
  return 1.0f - fabsf((float)(2 * k - 15)) * (1.0f / 16.0f);
}

__global__ void head_pad_kernel(const float* __restrict__ seg_w, const float* __restrict__ seg_b,
                                int C, int classes, float* __restrict__ wpad,
                                float* __restrict__ scale, float* __restrict__ shift,
                                int32_t* __restrict__ row_ptr, int32_t* __restrict__ kblk, int n_kb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < HEAD_CP * C) {
    const int r = i / C;
    wpad[i] = r < classes ? __ldg(seg_w + i) : 0.f;
  }
  if (i < HEAD_CP) {
    scale[i] = 1.f;
    shift[i] = i < classes ? __ldg(seg_b + i) : 0.f;
  }
  if (i < n_kb) kblk[i] = i;
  if (i == 0) { row_ptr[0] = 0; row_ptr[1] = n_kb; }
}

// One CTA = a 64 x 16 block of full-resolution pixels; its 10 x 4 low-resolution neighbourhood of class
// logits (<= 5 KB) is staged in shared memory once (the first version let every thread re-read its four
// neighbours from global memory: 1.3 GB of L2->SM traffic per batch, 0.24 ms).  One thread = 4 horizontally
// adjacent pixels (they share the same 2x2 low-res neighbourhood because x0 % 4 == 0).  Label bytes are
// packed into one 32-bit store, log-probs into float4 stores: every warp-level store instruction writes
// 128 / 512 contiguous bytes.
constexpr int UP_BW = 64, UP_BH = 16;               // output block
constexpr int UP_LW = UP_BW / 8 + 2, UP_LH = UP_BH / 8 + 2;   // low-res tile incl. the +-1 apron

template <int CLS_MAX>
__global__ void __launch_bounds__(256)
up_argmax_kernel(const float* __restrict__ L, int N, int h, int w, int classes,
                 uint8_t* __restrict__ labels, float* __restrict__ logprob) {
  constexpr int CP = (CLS_MAX + 3) & ~3;             // padded class pitch (float4 reads)
  __shared__ __align__(16) float s_l[UP_LH][UP_LW][CP];
  const int H = 8 * h, W = 8 * w;
  const int n = blockIdx.z;
  const int bx0 = blockIdx.x * UP_BW, by0 = blockIdx.y * UP_BH;
  const int lx0 = bx0 / 8 - 1, ly0 = by0 / 8 - 1;    // low-res origin of the staged tile
  for (int idx = threadIdx.x; idx < UP_LH * UP_LW * (CP / 4); idx += blockDim.x) {
    const int q4 = idx % (CP / 4);
    const int c = (idx / (CP / 4)) % UP_LW, r = idx / ((CP / 4) * UP_LW);
    const int ly = ly0 + r, lx = lx0 + c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);      // rows / columns outside the map contribute zero
    if (ly >= 0 && ly < h && lx >= 0 && lx < w)
      v = __ldg(reinterpret_cast<const float4*>(L + (((size_t)n * h + ly) * w + lx) * HEAD_CP) + q4);
    *reinterpret_cast<float4*>(&s_l[r][c][q4 * 4]) = v;
  }
  __syncthreads();
  const int x0 = bx0 + (threadIdx.x & 15) * 4, y = by0 + (threadIdx.x >> 4);
  if (x0 >= W || y >= H) return;

  const int i0 = (y + 4) >> 3, ky = (y + 4) & 7;
  const float wy0 = up_w(ky), wy1 = up_w(ky + 8);    // row i0 and row i0-1 (zeros were staged if outside)
  const int j0 = (x0 + 4) >> 3, kx = (x0 + 4) & 7;
  const int ra = i0 - ly0, ca = j0 - lx0;            // tile coordinates of (i0, j0); (i0-1, j0-1) = (ra-1, ca-1)

  // vertical pass: V0 = column j0, V1 = column j0-1
  float V0[CLS_MAX], V1[CLS_MAX];
#pragma unroll
  for (int q = 0; q < CP / 4; ++q) {
    const float4 a0 = *reinterpret_cast<const float4*>(&s_l[ra][ca][4 * q]);
    const float4 b0 = *reinterpret_cast<const float4*>(&s_l[ra - 1][ca][4 * q]);
    const float4 a1 = *reinterpret_cast<const float4*>(&s_l[ra][ca - 1][4 * q]);
    const float4 b1 = *reinterpret_cast<const float4*>(&s_l[ra - 1][ca - 1][4 * q]);
    const float a0v[4] = {a0.x, a0.y, a0.z, a0.w}, b0v[4] = {b0.x, b0.y, b0.z, b0.w};
    const float a1v[4] = {a1.x, a1.y, a1.z, a1.w}, b1v[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * q + e < CLS_MAX) {
        V0[4 * q + e] = wy0 * a0v[e] + wy1 * b0v[e];
        V1[4 * q + e] = wy0 * a1v[e] + wy1 * b1v[e];
      }
  }

  uint32_t packed = 0;
  float lse[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float wx0 = up_w(kx + t), wx1 = up_w(kx + t + 8);
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < CLS_MAX; ++c) {
      if (c < classes) {
        const float v = wx0 * V0[c] + wx1 * V1[c];
        if (v > best) { best = v; arg = c; }  // strict > : first maximum wins (torch.max)
      }
    }
    packed |= (uint32_t)arg << (8 * t);
    if (logprob != nullptr) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CLS_MAX; ++c)
        if (c < classes) s += expf(wx0 * V0[c] + wx1 * V1[c] - best);
      lse[t] = best + logf(s);
    }
  }
  const size_t pix = ((size_t)n * H + y) * W + x0;
  if (labels != nullptr) *reinterpret_cast<uint32_t*>(labels + pix) = packed;
  if (logprob != nullptr) {
#pragma unroll
    for (int c = 0; c < CLS_MAX; ++c) {
      if (c < classes) {
        float4 o;
        o.x = up_w(kx + 0) * V0[c] + up_w(kx + 8) * V1[c] - lse[0];
        o.y = up_w(kx + 1) * V0[c] + up_w(kx + 9) * V1[c] - lse[1];
        o.z = up_w(kx + 2) * V0[c] + up_w(kx + 10) * V1[c] - lse[2];
        o.w = up_w(kx + 3) * V0[c] + up_w(kx + 11) * V1[c] - lse[3];
        *reinterpret_cast<float4*>(logprob + (((size_t)n * classes + c) * H + y) * W + x0) = o;
      }
    }
  }
}

// [N,h,w,32] float32 -> [N,classes,h,w] float32  (DRNSeg.forward()[1])
__global__ void logits_to_nchw_kernel(const float* __restrict__ L, int N, int h, int w, int classes,
                                      float* __restrict__ out) {
  const int64_t total = (int64_t)N * classes * h * w;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % w);
  const int y = (int)((i / w) % h);
  const int c = (int)((i / ((int64_t)w * h)) % classes);
  const int n = (int)(i / ((int64_t)w * h * classes));
  out[i] = __ldg(L + (((size_t)n * h + y) * w + x) * HEAD_CP + c);
}

}  // namespace drnb200

using namespace drnb200;

extern "C" void drnb200_head_plan_destroy(drnb200_head_plan* plan) {
  if (!plan) return;
  if (plan->conv) drnb200_conv_plan_destroy(plan->conv);
  cudaFree(plan->d_row_ptr); cudaFree(plan->d_kblk); cudaFree(plan->d_wpacked);
  cudaFree(plan->d_wpad); cudaFree(plan->d_scale); cudaFree(plan->d_shift); cudaFree(plan->d_logits);
  delete plan;
}

extern "C" int drnb200_head_plan_create(drnb200_head_plan** out, int N, int h, int w, int C,
                                        int classes, int act_dtype, const float* seg_w,
                                        const float* seg_b, void* stream) {
  DRN_REQUIRE(out && seg_w && seg_b, "head_plan_create: null pointer");
  DRN_REQUIRE(N > 0 && N <= 65535 && h > 0 && w > 0 && C > 0 && C % 16 == 0, "head_plan_create: bad shape");
  DRN_REQUIRE(classes > 0 && classes <= HEAD_CP, "head_plan_create: classes must be in [1,32]");
  DRN_REQUIRE(act_dtype == DRNB200_BF16 || act_dtype == DRNB200_F16, "head_plan_create: bad act_dtype");
  drnb200_head_plan* p = new (std::nothrow) drnb200_head_plan();
  if (!p) { set_error("head_plan_create: out of host memory"); return DRNB200_E_NOMEM; }
  *p = drnb200_head_plan{};
  p->N = N; p->h = h; p->w = w; p->C = C; p->classes = classes; p->act_dtype = act_dtype;
  const int tile_ci = (C % 64 == 0) ? 64 : (C % 32 == 0 ? 32 : 16);
  const int n_kb = C / tile_ci;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes); };
  alloc((void**)&p->d_row_ptr, 2 * sizeof(int32_t));
  alloc((void**)&p->d_kblk, n_kb * sizeof(int32_t));
  alloc((void**)&p->d_wpacked, (size_t)HEAD_CP * C * 2);
  alloc((void**)&p->d_wpad, (size_t)HEAD_CP * C * sizeof(float));
  alloc((void**)&p->d_scale, HEAD_CP * sizeof(float));
  alloc((void**)&p->d_shift, HEAD_CP * sizeof(float));
  alloc((void**)&p->d_logits, (size_t)N * h * w * HEAD_CP * sizeof(float));
  if (e != cudaSuccess) { drnb200_head_plan_destroy(p); return cuda_fail(e, "cudaMalloc(head plan)"); }
  const int total = HEAD_CP * C;
  head_pad_kernel<<<(total + 255) / 256, 256, 0, st>>>(seg_w, seg_b, C, classes, p->d_wpad, p->d_scale,
                                                      p->d_shift, p->d_row_ptr, p->d_kblk, n_kb);
  int rc = drnb200_pack_weights(p->d_wpad, nullptr, HEAD_CP, C, 1, 1, HEAD_CP, tile_ci, p->d_row_ptr,
                                p->d_kblk, act_dtype, p->d_wpacked, stream);
  if (rc == DRNB200_OK) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize(head plan)");
  }
  if (rc == DRNB200_OK) {
    drnb200_conv_desc d{};
    d.N = N; d.H = h; d.W = w; d.Cin = C; d.Cout = HEAD_CP; d.ksize = 1; d.stride = 1; d.dilation = 1;
    d.relu = 0; d.has_residual = 0; d.act_dtype = act_dtype; d.out_f32 = 1;
    d.tile_o = HEAD_CP; d.tile_ci = tile_ci; d.impl = DRNB200_IMPL_AUTO;
    rc = drnb200_conv_plan_create(&p->conv, &d, p->d_row_ptr, p->d_kblk, p->d_wpacked, p->d_scale,
                                  p->d_shift);
  }
  if (rc != DRNB200_OK) { drnb200_head_plan_destroy(p); return rc; }
  *out = p;
  return DRNB200_OK;
}

extern "C" int drnb200_head_forward(drnb200_head_plan* plan, const void* x_nhwc, uint8_t* labels,
                                    float* seg_logits, float* logprob, void* stream) {
  DRN_REQUIRE(plan && x_nhwc, "head_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = drnb200_conv_forward(plan->conv, x_nhwc, nullptr, plan->d_logits, stream);
  if (rc) return rc;
  if (labels || logprob) {
    const dim3 blocks((8 * plan->w + UP_BW - 1) / UP_BW, (8 * plan->h + UP_BH - 1) / UP_BH, plan->N);
    if (plan->classes <= 19)
      up_argmax_kernel<19><<<blocks, 256, 0, st>>>(plan->d_logits, plan->N, plan->h, plan->w,
                                                   plan->classes, labels, logprob);
    else
      up_argmax_kernel<32><<<blocks, 256, 0, st>>>(plan->d_logits, plan->N, plan->h, plan->w,
                                                   plan->classes, labels, logprob);
    DRN_CUDA(cudaGetLastError());
  }
  if (seg_logits) {
    const int64_t total = (int64_t)plan->N * plan->classes * plan->h * plan->w;
    logits_to_nchw_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(plan->d_logits, plan->N, plan->h,
                                                                    plan->w, plan->classes, seg_logits);
    DRN_CUDA(cudaGetLastError());
  }
  return DRNB200_OK;
}
