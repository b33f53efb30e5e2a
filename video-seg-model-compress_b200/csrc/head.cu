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
#include <algorithm>
#include <cstdlib>
#include <cudaTypedefs.h>
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
  int fused_ok;      // labels-only calls run head_fused_kernel
  int up_mode;       // 0 = fixed-bilinear ConvTranspose2d (default), 1 = UpsamplingBilinear2d (align_corners=True)
  const void* fmap_ptr;
  CUtensorMap fmap;  // [N,h,w,C] activations, box {64, 16, 8, 1}
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

// ---- upsample core shared by up_argmax_kernel and head_fused_kernel.  Every product/sum is an explicitly rounded
//      operation so that the two kernels produce bit-identical values (predict() == argmax of forward()).
// vertical pass at low-res tile position `at` = &tile[ra][ca][0] (rows `row_pitch` floats apart, class pitch CP):
// V0 = column j0, V1 = column j0-1, each = wy0 * row i0 + wy1 * row i0-1 (zeros were staged outside the map)
template <int CLS_MAX>
__device__ __forceinline__ void up_vertical(const float* at, int row_pitch, int ky, float (&V0)[CLS_MAX],
                                            float (&V1)[CLS_MAX]) {
  constexpr int CP = (CLS_MAX + 3) & ~3;
  const float wy0 = up_w(ky), wy1 = up_w(ky + 8);
#pragma unroll
  for (int q = 0; q < CP / 4; ++q) {
    const float4 a0 = *reinterpret_cast<const float4*>(at + 4 * q);
    const float4 b0 = *reinterpret_cast<const float4*>(at - row_pitch + 4 * q);
    const float4 a1 = *reinterpret_cast<const float4*>(at - CP + 4 * q);
    const float4 b1 = *reinterpret_cast<const float4*>(at - row_pitch - CP + 4 * q);
    const float a0v[4] = {a0.x, a0.y, a0.z, a0.w}, b0v[4] = {b0.x, b0.y, b0.z, b0.w};
    const float a1v[4] = {a1.x, a1.y, a1.z, a1.w}, b1v[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * q + e < CLS_MAX) {
        V0[4 * q + e] = __fmaf_rn(wy1, b0v[e], __fmul_rn(wy0, a0v[e]));
        V1[4 * q + e] = __fmaf_rn(wy1, b1v[e], __fmul_rn(wy0, a1v[e]));
      }
  }
}
// horizontal pass: the two taps of an output column always sum to one (w[k] + w[k+8] = 1 for k = 0..7), so
//   out = w0*V0 + w1*V1 = V1 + w0*(V0 - V1)
// is evaluated as ONE fma per (pixel, class) on D = V0 - V1, computed once per class for all pixels of the thread
// (zero padding is already in V0/V1: rows and columns outside the map were staged as zeros).  Every kernel of this
// file goes through up_h(), so predict() and forward() see bit-identical values.
__device__ __forceinline__ float up_h(int k, float d, float v1) { return __fmaf_rn(up_w(k), d, v1); }

// horizontal pass + argmax of 4 adjacent pixels (same 2x2 low-res neighbourhood): labels packed into 4 bytes;
// strict > : the first maximum wins like torch.max
template <int CLS_MAX>
__device__ __forceinline__ uint32_t up_argmax4(const float (&V0)[CLS_MAX], const float (&V1)[CLS_MAX], int kx,
                                               int classes, float (&best4)[4]) {
  uint32_t packed = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < CLS_MAX; ++c) {
      if (c < classes) {
        const float v = up_h(kx + t, __fsub_rn(V0[c], V1[c]), V1[c]);
        if (v > best) { best = v; arg = c; }
      }
    }
    best4[t] = best;
    packed |= (uint32_t)arg << (8 * t);
  }
  return packed;
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
  const int j0 = (x0 + 4) >> 3, kx = (x0 + 4) & 7;
  const int ra = i0 - ly0, ca = j0 - lx0;            // tile coordinates of (i0, j0); (i0-1, j0-1) = (ra-1, ca-1)
  float V0[CLS_MAX], V1[CLS_MAX];
  up_vertical<CLS_MAX>(&s_l[ra][ca][0], UP_LW * CP, ky, V0, V1);
  float best4[4];
  const uint32_t packed = up_argmax4<CLS_MAX>(V0, V1, kx, classes, best4);
  float lse[4];
  if (logprob != nullptr) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CLS_MAX; ++c)
        if (c < classes) s += expf(up_h(kx + t, __fsub_rn(V0[c], V1[c]), V1[c]) - best4[t]);
      lse[t] = best4[t] + logf(s);
    }
  }
  const size_t pix = ((size_t)n * H + y) * W + x0;
  if (labels != nullptr) *reinterpret_cast<uint32_t*>(labels + pix) = packed;
  if (logprob != nullptr) {
#pragma unroll
    for (int c = 0; c < CLS_MAX; ++c) {
      if (c < classes) {
        float4 o;
        const float d = __fsub_rn(V0[c], V1[c]);
        o.x = up_h(kx + 0, d, V1[c]) - lse[0];
        o.y = up_h(kx + 1, d, V1[c]) - lse[1];
        o.z = up_h(kx + 2, d, V1[c]) - lse[2];
        o.w = up_h(kx + 3, d, V1[c]) - lse[3];
        *reinterpret_cast<float4*>(logprob + (((size_t)n * classes + c) * H + y) * W + x0) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// head_fused_kernel — (c) as ONE kernel: classifier GEMM (tcgen05, TMEM) -> low-res class logits in shared memory
// -> fixed-bilinear x8 upsample -> argmax -> packed label stores.  Nothing but the label map is written.
// Persistent CTAs.  The transposed conv makes output pixel x depend on the low-res columns j0-1 and j0 with
// j0 = (x+4)>>3, so the natural unit is the CELL of 8x8 output pixels x = 8*j0-4 .. 8*j0+3 (same for y) that share one
// 2x2 low-res neighbourhood.  Tile = 15 x 7 cells = 120 x 56 output pixels, whose neighbourhoods are exactly the
// 16 x 8 low-res pixels = the 128 rows of one UMMA M tile (round 1 tiled the OUTPUT on multiples of 8 instead: 14 x 6
// cells per 16 x 8 low-res tile, 25 % more tiles for the same frame).
//   warp 0      TMA: per tile C/64 K-blocks, box {64 ch, 16, 8} (out-of-map pixels arrive as zeros)
//   warp 1      MMA: D[128 px][32 classes] += A[128 x 64] * W[32 x 64]^T, classifier weights resident in smem
//   warps 2-5   TMEM -> + bias (0 for pixels outside the map: the transposed conv pads with zeros, not with the
//               bias) -> shared-memory logits [128 px][20]
//   warps 6-19  upsample + argmax.  One thread = the 8 pixels of one cell in one output row: the vertical
//               interpolation of the two low-res columns (2 x 19 values) is computed ONCE per 8 pixels (round 1:
//               once per 4), the eight horizontal weights are compile-time constants, and the labels leave as two
//               32-bit stores.  Arithmetic per value is the same explicitly rounded sequence as up_argmax_kernel.
constexpr int HF_LW = 16, HF_LH = 8;                  // low-res tile (UMMA M = 128 pixels)
constexpr int HF_CX = HF_LW - 1, HF_CY = HF_LH - 1;   // 15 x 7 cells per tile
constexpr int HF_STAGES = 6, HF_ACC = 4, HF_LBUF = 2;
constexpr int HF_UP_WARPS = 14;
constexpr int HF_THREADS = (6 + HF_UP_WARPS) * 32;
constexpr int HF_CP = 20;                             // class pitch of the staged logits (19 classes, float4 reads)
constexpr int HF_MAX_KB = 16;                         // C <= 1024

// horizontal pass + argmax for the 8 pixels of one cell (kx = 0..7 compile-time): same operations, in the same order,
// as up_argmax4 performs for its four pixels -> bit-identical values, first maximum wins
template <int CLS_MAX>
__device__ __forceinline__ void up_argmax8(const float (&V0)[CLS_MAX], const float (&V1)[CLS_MAX], int classes,
                                           uint32_t cand, uint32_t& lo, uint32_t& hi) {
  float best[8];
  uint32_t arg[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { best[t] = -INFINITY; arg[t] = 0u; }
#pragma unroll
  for (int c = 0; c < CLS_MAX; ++c) {
    if (c < classes && ((cand >> c) & 1u)) {     // classes that cannot win anywhere in this cell are skipped (cell_candidates)
      const float d = __fsub_rn(V0[c], V1[c]);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float v = up_h(t, d, V1[c]);
        if (v > best[t]) { best[t] = v; arg[t] = (uint32_t)c; }
      }
    }
  }
  lo = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
  hi = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
}

// Candidate classes of one cell = the 2x2 low-res neighbourhood p00 | p01 / p10 | p11 (class pitch 1, HF_CP floats per
// pixel).  Every output pixel of the cell is a convex combination (weights >= 0, sum <= 1; corners outside the map
// are zero for ALL classes) of the four corner logits, so a class k with  L_x[j] - L_x[k] > delta  at all four
// corners x for some class j loses to j at every pixel of the cell by more than delta * (sum of weights) — with
// delta = 1e-3 * max(1, max |L|) that is three orders of magnitude above the rounding error of the interpolation
// (a few ulps of max |L|), so k can be neither the argmax nor tie with it and skipping it leaves the labels
// bit-identical to the full evaluation (forward()'s log-probs + torch.max).  j = the class with the best worst
// corner.  Border cells (a zero-padded corner: differences are 0) and NaNs keep all classes.
template <int CLS_MAX>
__device__ __forceinline__ uint32_t cell_candidates(const float* p00, const float* p01, const float* p10,
                                                    const float* p11, int classes) {
  float best_min = -INFINITY, mag = 0.f;
  int js = 0;
#pragma unroll
  for (int k = 0; k < CLS_MAX; ++k) {
    if (k < classes) {
      const float a = p00[k], b = p01[k], c = p10[k], d = p11[k];
      const float mn = fminf(fminf(a, b), fminf(c, d));
      mag = fmaxf(mag, fmaxf(fmaxf(fabsf(a), fabsf(b)), fmaxf(fabsf(c), fabsf(d))));
      if (mn > best_min) { best_min = mn; js = k; }
    }
  }
  const float delta = 1e-3f * fmaxf(1.f, mag);
  const float ja = p00[js], jb = p01[js], jc = p10[js], jd = p11[js];
  uint32_t cand = 0u;
#pragma unroll
  for (int k = 0; k < CLS_MAX; ++k) {
    if (k < classes) {
      const float g = fminf(fminf(ja - p00[k], jb - p01[k]), fminf(jc - p10[k], jd - p11[k]));
      if (!(g > delta)) cand |= 1u << k;          // keeps js itself (g = 0), ties, near-ties and NaNs
    }
  }
  return cand;
}

struct HeadFusedParams {
  const uint8_t* w_packed;     // C/64 tiles of 32 x 64 (128-byte rows, SWIZZLE_128B)
  const float* bias;           // [32], zero padded
  uint8_t* labels;             // [N, 8h, 8w]
  int N, h, w, n_kb, classes, tiles_x, tiles_y, total_tiles;
  uint32_t idesc;
  int prune;                   // 1: skip classes that cannot win in a cell (cell_candidates); 0: evaluate all (A/B knob)
};

struct __align__(16) HFSync {
  alignas(16) float logits[HF_LBUF][HF_LW * HF_LH][HF_CP];
  alignas(16) float bias[32];
  uint64_t full[HF_STAGES], empty[HF_STAGES], tfull[HF_ACC], tempty[HF_ACC], lfull[HF_LBUF], lempty[HF_LBUF], wfull;
  uint32_t tmem_base, pad;
  int chunk[HF_LBUF];            // next 32-item chunk of the tile staged in logits[lb] (claimed by the upsample warps)
  uint32_t cand[HF_LBUF][HF_CX * HF_CY];   // per cell: classes that can win somewhere in the cell (cell_candidates)
};

template <int DT, bool PRUNE>
__global__ void __launch_bounds__(HF_THREADS, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const HeadFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* ring = smem;                                   // HF_STAGES x 16 KB: [128 px][64 ch] 16-bit
  uint8_t* wsm = smem + HF_STAGES * 16384;                // n_kb x 4 KB
  __shared__ HFSync sync_s;
  HFSync* sync = &sync_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < HF_STAGES; ++i) { mbar_init(&sync->full[i], 1); mbar_init(&sync->empty[i], 1); }
    for (int i = 0; i < HF_ACC; ++i) { mbar_init(&sync->tfull[i], 1); mbar_init(&sync->tempty[i], 4); }
    for (int i = 0; i < HF_LBUF; ++i) { mbar_init(&sync->lfull[i], 4); mbar_init(&sync->lempty[i], HF_UP_WARPS); }
    mbar_init(&sync->wfull, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < HF_LBUF) sync->chunk[threadIdx.x] = 0;
  if (warp == 1) {
    tmem_alloc(&sync->tmem_base, HF_ACC * 32);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  if (threadIdx.x < 32) sync->bias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sync->tmem_base;
  const int tiles_per_frame = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================================================================= TMA producer (warp-uniform loop)
    if (elect_one()) {
      mbar_arrive_expect_tx(&sync->wfull, (uint32_t)p.n_kb * 4096u);
      bulk_load(p.w_packed, &sync->wfull, wsm, (uint32_t)p.n_kb * 4096u);
    }
    uint32_t s = 0, ph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int n = t / tiles_per_frame, r = t - n * tiles_per_frame;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int lx0 = tx * HF_CX - 1, ly0 = ty * HF_CY - 1;
      for (int kb = 0; kb < p.n_kb; ++kb) {
        mbar_wait_relaxed(&sync->empty[s], ph ^ 1u, 500);
        if (elect_one()) {
          mbar_arrive_expect_tx(&sync->full[s], 16384u);
          tma_load_4d(&tmap_x, &sync->full[s], ring + s * 16384, kb * 64, lx0, ly0, n);
        }
        __syncwarp();
        if (++s == HF_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (warp-uniform loop)
    mbar_wait(&sync->wfull, 0);
    const uint64_t d_hi = umma_smem_desc(0u, 128);
    const uint32_t w16 = smem_u32(wsm) >> 4, r16 = smem_u32(ring) >> 4;
    uint32_t s = 0, ph = 0;
    int i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const int acc = i % HF_ACC;
      mbar_wait_relaxed(&sync->tempty[acc], ((uint32_t)(i / HF_ACC) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 32u;
      for (int kb = 0; kb < p.n_kb; ++kb) {
        mbar_wait(&sync->full[s], ph);
        tc_fence_after();
        const uint32_t a16 = r16 + s * (16384u >> 4), b16 = w16 + (uint32_t)kb * (4096u >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(d_tmem, d_hi | (uint64_t)(a16 + 2u * ks), d_hi | (uint64_t)(b16 + 2u * ks), p.idesc,
                     (kb > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&sync->empty[s]);
        }
        __syncwarp();
        if (++s == HF_STAGES) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(&sync->tfull[acc]);
      __syncwarp();
    }
  } else if (warp < 6) {
    // ================================================================= logits: TMEM -> +bias -> shared memory
    const int q = warp & 3, m = q * 32 + lane;          // TMEM lane = low-res pixel (row m / 16, column m % 16)
    int i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const int acc = i % HF_ACC, lb = i % HF_LBUF;
      const int n = t / tiles_per_frame, r = t - n * tiles_per_frame;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int lx = tx * HF_CX - 1 + (m & (HF_LW - 1)), ly = ty * HF_CY - 1 + (m >> 4);
      const bool inside = lx >= 0 && lx < p.w && ly >= 0 && ly < p.h;
      mbar_wait_relaxed(&sync->tfull[acc], (uint32_t)(i / HF_ACC) & 1u, 64);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + (uint32_t)acc * 32u + ((uint32_t)(q * 32) << 16), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sync->tempty[acc]);
      mbar_wait_relaxed(&sync->lempty[lb], ((uint32_t)(i / HF_LBUF) & 1u) ^ 1u, 1000);   // a whole tile of slack
      if (m == 0) sync->chunk[lb] = 0;                    // every upsample warp has left this buffer
      float* dst = &sync->logits[lb][m][0];
#pragma unroll
      for (int c4 = 0; c4 < HF_CP / 4; ++c4) {
        float4 o;
        o.x = inside ? __fadd_rn(__uint_as_float(v[4 * c4]), sync->bias[4 * c4]) : 0.f;
        o.y = inside ? __fadd_rn(__uint_as_float(v[4 * c4 + 1]), sync->bias[4 * c4 + 1]) : 0.f;
        o.z = inside ? __fadd_rn(__uint_as_float(v[4 * c4 + 2]), sync->bias[4 * c4 + 2]) : 0.f;
        o.w = inside ? __fadd_rn(__uint_as_float(v[4 * c4 + 3]), sync->bias[4 * c4 + 3]) : 0.f;
        *reinterpret_cast<float4*>(dst + 4 * c4) = o;
      }
      if (PRUNE) {
        named_bar_sync(2, 128);                             // all 128 low-res pixels of the tile are staged
        if (m < HF_CX * HF_CY) {                            // one cell per thread
          const int cy = m / HF_CX, cx = m - cy * HF_CX;
          const float* p00 = &sync->logits[lb][cy * HF_LW + cx][0];
          sync->cand[lb][m] = cell_candidates<19>(p00, p00 + HF_CP, p00 + HF_LW * HF_CP, p00 + (HF_LW + 1) * HF_CP,
                                                  p.classes);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sync->lfull[lb]);
    }
  } else {
    // ================================================================= upsample + argmax
    const int H = 8 * p.h, W = 8 * p.w;
    constexpr int ITEMS = HF_CY * 8 * HF_CX;            // 56 output rows x 15 cells
    int i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const int lb = i % HF_LBUF;
      const int n = t / tiles_per_frame, r = t - n * tiles_per_frame;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int X0 = tx * (HF_CX * 8) - 4, Y0 = ty * (HF_CY * 8) - 4;      // first pixel of cell (0, 0)
      mbar_wait(&sync->lfull[lb], (uint32_t)(i / HF_LBUF) & 1u);
      const float* tile = &sync->logits[lb][0][0];
      // 32-item chunks are claimed dynamically: the four SM sub-partitions host different numbers of upsample warps
      // (and the polling warps of the other roles), a static split left the slowest one 10 % behind
      for (;;) {
        int ck = 0;
        if (lane == 0) ck = atomicAdd(&sync->chunk[lb], 1);
        ck = __shfl_sync(0xffffffffu, ck, 0);
        if (ck * 32 >= ITEMS) break;
        const int it = ck * 32 + lane;
        if (it < ITEMS) {
          const int yy = it / HF_CX, cx = it - yy * HF_CX;
          const int y = Y0 + yy, x0 = X0 + 8 * cx;
          if (y >= 0 && y < H && x0 < W) {
            // cell row yy>>3 reads low-res tile rows (yy>>3) and (yy>>3)+1, cell column cx reads tile columns cx, cx+1
            float V0[19], V1[19];
            up_vertical<19>(tile + (((yy >> 3) + 1) * HF_LW + cx + 1) * HF_CP, HF_LW * HF_CP, yy & 7, V0, V1);
            uint32_t lo, hi;
            up_argmax8<19>(V0, V1, p.classes, PRUNE ? sync->cand[lb][(yy >> 3) * HF_CX + cx] : 0xFFFFFFFFu, lo, hi);
            uint8_t* row = p.labels + ((size_t)n * H + y) * W;
            if (x0 >= 0) *reinterpret_cast<uint32_t*>(row + x0) = lo;
            if (x0 + 4 < W) *reinterpret_cast<uint32_t*>(row + x0 + 4) = hi;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sync->lempty[lb]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, HF_ACC * 32);
  }
}

// use_torch_up=True (semantic_seg.py:144-145): nn.UpsamplingBilinear2d(scale_factor=8) == bilinear interpolation with
// align_corners=True.  src = y * (h-1)/(H-1); the four neighbours are blended as torch's CPU kernel does it
// (rows first: w0*(a0*v00 + a1*v01) + w1*(a0*v10 + a1*v11), all fp32).  Thread = one output pixel; the low-res logits
// [N,h,w,32] are L2-resident.  Parity path only (labels, log-probs); the fused kernel serves the default `up`.
template <int CLS_MAX>
__global__ void __launch_bounds__(256)
up_aligned_kernel(const float* __restrict__ L, int N, int h, int w, int classes, uint8_t* __restrict__ labels,
                  float* __restrict__ logprob) {
  const int H = 8 * h, W = 8 * w;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (x >= W) return;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const float fy = sy * (float)y, fx = sx * (float)x;
  const int i0 = (int)fy, j0 = (int)fx;
  const int i1 = min(i0 + 1, h - 1), j1 = min(j0 + 1, w - 1);
  const float w1 = fy - (float)i0, w0 = 1.f - w1, a1 = fx - (float)j0, a0 = 1.f - a1;
  const float* r00 = L + (((size_t)n * h + i0) * w + j0) * HEAD_CP;
  const float* r01 = L + (((size_t)n * h + i0) * w + j1) * HEAD_CP;
  const float* r10 = L + (((size_t)n * h + i1) * w + j0) * HEAD_CP;
  const float* r11 = L + (((size_t)n * h + i1) * w + j1) * HEAD_CP;
  float v[CLS_MAX];
  float best = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int c = 0; c < CLS_MAX; ++c) {
    if (c < classes) {
      const float top = __fmaf_rn(a1, __ldg(r01 + c), __fmul_rn(a0, __ldg(r00 + c)));
      const float bot = __fmaf_rn(a1, __ldg(r11 + c), __fmul_rn(a0, __ldg(r10 + c)));
      v[c] = __fmaf_rn(w1, bot, __fmul_rn(w0, top));
      if (v[c] > best) { best = v[c]; arg = c; }
    }
  }
  const size_t pix = ((size_t)n * H + y) * W + x;
  if (labels != nullptr) labels[pix] = (uint8_t)arg;
  if (logprob != nullptr) {
    float ssum = 0.f;
#pragma unroll
    for (int c = 0; c < CLS_MAX; ++c)
      if (c < classes) ssum += expf(v[c] - best);
    const float lse = best + logf(ssum);
#pragma unroll
    for (int c = 0; c < CLS_MAX; ++c)
      if (c < classes) logprob[(((size_t)n * classes + c) * H + y) * W + x] = v[c] - lse;
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
  static const char* env_hf = getenv("DRNB200_HEAD");      // A/B knob: "split" keeps GEMM + upsample as two launches
  p->fused_ok = (tile_ci == 64 && n_kb <= HF_MAX_KB && classes <= 19 && !(env_hf && env_hf[0] == 's')) ? 1 : 0;
  p->fmap_ptr = nullptr;
  *out = p;
  return DRNB200_OK;
}

static int head_fused_launch(drnb200_head_plan* plan, const void* x, uint8_t* labels, cudaStream_t st) {
  if (plan->fmap_ptr != x) {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
      void* sym = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
          qres != cudaDriverEntryPointSuccess) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return DRNB200_E_CUDA;
      }
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
    }
    cuuint64_t gdim[4] = {(cuuint64_t)plan->C, (cuuint64_t)plan->w, (cuuint64_t)plan->h, (cuuint64_t)plan->N};
    cuuint64_t gstr[3] = {(cuuint64_t)plan->C * 2, (cuuint64_t)plan->w * plan->C * 2,
                          (cuuint64_t)plan->h * plan->w * plan->C * 2};
    cuuint32_t box[4] = {64, HF_LW, HF_LH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&plan->fmap, plan->act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                  : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(head) failed with CUresult %d", (int)r); return DRNB200_E_CUDA; }
    plan->fmap_ptr = x;
  }
  HeadFusedParams p{};
  p.w_packed = reinterpret_cast<const uint8_t*>(plan->d_wpacked);
  p.bias = plan->d_shift; p.labels = labels;
  p.N = plan->N; p.h = plan->h; p.w = plan->w; p.n_kb = plan->C / 64; p.classes = plan->classes;
  p.tiles_x = (plan->w + 1 + HF_CX - 1) / HF_CX;      // cells j0 = 0 .. w (the first and the last are half cells)
  p.tiles_y = (plan->h + 1 + HF_CY - 1) / HF_CY;
  p.total_tiles = plan->N * p.tiles_x * p.tiles_y;
  p.idesc = umma_idesc_f16(128, HEAD_CP, plan->act_dtype);
  // Opt-in (DRNB200_HEAD_PRUNE=1): skip classes that cannot win in a cell.  Labels stay bit-identical (135 GPU tests pass
  // with it on), but on the benchmark's white-noise frames + random-init network few classes are dominated and the
  // divergent class loop costs more than it saves: 0.110-0.112 ms without vs 0.115-0.120 ms with (same box, A/B).
  static const char* env_pr = getenv("DRNB200_HEAD_PRUNE");
  p.prune = (env_pr && env_pr[0] == '1') ? 1 : 0;
  const size_t smem = 1024 + HF_STAGES * 16384 + (size_t)p.n_kb * 4096;
  constexpr size_t kHfMaxSmem = 1024 + HF_STAGES * 16384 + (size_t)HF_MAX_KB * 4096;
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_tiles, sms);
  if (grid == 0) return DRNB200_OK;
  static std::atomic<unsigned long long> attr[4];
  void (*kern)(const CUtensorMap, const HeadFusedParams) =
      plan->act_dtype == DRNB200_BF16 ? (p.prune ? head_fused_kernel<DRNB200_BF16, true> : head_fused_kernel<DRNB200_BF16, false>)
                                      : (p.prune ? head_fused_kernel<DRNB200_F16, true> : head_fused_kernel<DRNB200_F16, false>);
  if (attr_needed_on_this_device(attr[(plan->act_dtype == DRNB200_BF16 ? 0 : 2) + (p.prune ? 1 : 0)]))
    DRN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHfMaxSmem));
  launch_chained(kern, grid, HF_THREADS, smem, st, plan->fmap, p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

extern "C" int drnb200_head_plan_fused(const drnb200_head_plan* plan) {
  return (plan && plan->fused_ok && plan->up_mode == 0) ? 1 : 0;
}

extern "C" int drnb200_head_plan_set_upsample(drnb200_head_plan* plan, int mode) {
  DRN_REQUIRE(plan, "head_plan_set_upsample: null plan");
  DRN_REQUIRE(mode == DRNB200_UP_TRANSPOSED || mode == DRNB200_UP_ALIGNED, "head_plan_set_upsample: unknown mode %d", mode);
  plan->up_mode = mode;
  return DRNB200_OK;
}

extern "C" int drnb200_head_forward(drnb200_head_plan* plan, const void* x_nhwc, uint8_t* labels,
                                    float* seg_logits, float* logprob, void* stream) {
  DRN_REQUIRE(plan && x_nhwc, "head_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (labels && !seg_logits && !logprob && plan->fused_ok && plan->up_mode == 0)   // the fast path: one kernel, labels only
    return head_fused_launch(plan, x_nhwc, labels, st);
  int rc = drnb200_conv_forward(plan->conv, x_nhwc, nullptr, plan->d_logits, stream);
  if (rc) return rc;
  if ((labels || logprob) && plan->up_mode == DRNB200_UP_ALIGNED) {
    const dim3 blocks((8 * plan->w + 255) / 256, 8 * plan->h, plan->N);
    if (plan->classes <= 19)
      up_aligned_kernel<19><<<blocks, 256, 0, st>>>(plan->d_logits, plan->N, plan->h, plan->w, plan->classes, labels, logprob);
    else
      up_aligned_kernel<32><<<blocks, 256, 0, st>>>(plan->d_logits, plan->N, plan->h, plan->w, plan->classes, labels, logprob);
    DRN_CUDA(cudaGetLastError());
  } else if (labels || logprob) {
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
