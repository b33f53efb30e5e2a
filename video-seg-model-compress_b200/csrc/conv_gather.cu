// conv_gather.cu — small-Cin convolutions on tcgen05 with an im2col tile built by the CTA's own threads.
//
// The TMA path of conv_tc.cu issues one box per filter tap; with 16 input channels a tap is a 32-byte row
// and the kernel becomes TMA-issue / L2-request bound (measured: 2.9 ms for DRN-D-22 layer1 at batch 8,
// 17x its HBM floor).  Here the 128-pixel x K im2col tile is gathered with plain 16-byte loads (the 9x
// tap re-use is served by L1), written straight into the UMMA SWIZZLE_32B K-major layout, and multiplied
// by weights that stay resident in shared memory for the whole persistent CTA:
//   MODE 3x3 : x NHWC 16-bit with exactly 16 channels, K-block = one tap (drn.py:201-211 layer1/layer2)
//   MODE stem: x NCHW float32 with 3 channels, 7x7, K = 147 padded to 160; the fp32 -> 16-bit conversion of
//              the frame is fused into the gather (drn.py:132-137; callers pass fp32 NCHW, semantic_seg.py:440)
// GEMM orientation: M = 128 output pixels (TMEM lanes), N = Cout (16 or 32 columns), K-step = 16.
// Roles (288 threads): warps 0-7 gather (256 threads: pixel m = tid & 127, chunk parity = tid >> 7),
// warps 0-3 also run the epilogue of the previous tile, warp 8 allocates TMEM and issues the MMAs.
// Two im2col buffers and two TMEM accumulators software-pipeline gather(i+1) | MMA(i) | epilogue(i-1).
#include "conv_internal.cuh"
#include <algorithm>
#include <new>

namespace drnb200 {

constexpr int G_THREADS = 288;
constexpr int G_MAX_KB = 10;            // 9 taps, or 160/16 stem K-blocks
constexpr int G_KB_BYTES = 128 * 32;    // one K-block of the im2col tile: 128 pixels x 16 elements
constexpr int G_ABUF_BYTES = G_MAX_KB * G_KB_BYTES;
constexpr uint32_t G_TMEM_COLS = 64;    // 2 accumulators x 32 columns

struct GatherParams {
  const void* x;
  void* y;
  const uint8_t* w_packed;   // n_kb tiles of Cout x 16, 32-byte rows, SWIZZLE_32B image (compact.cu)
  const int32_t* kblk;       // live K-block ids (3x3: tap; stem: k / 16)
  const float* scale;
  const float* shift;
  int n_kb;
  int N, H, W, OH, OW, Cout, stride, relu;
  int stem;                  // 0: 3x3 over NHWC 16-channel input, 1: 7x7 over NCHW fp32 3-channel input
  int TW, TH, tw_shift, tiles_x, tiles_y, total_tiles;
  uint32_t idesc;
};

struct __align__(8) GSync {
  uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2];
  uint32_t tmem_base, pad;
};

// k -> (ci, ky-3, kx-3) of the 7x7 stem, k = ci*49 + ky*7 + kx (OIHW flattening); k >= 147 is padding
__constant__ int c_stem_lut[160];

template <int DT>
__global__ void __launch_bounds__(G_THREADS, 1) conv_gather_kernel(const GatherParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* abuf = smem;                                   // 2 x G_ABUF_BYTES
  uint8_t* wsm = smem + 2 * G_ABUF_BYTES;                 // n_kb x Cout x 32 B
  GSync* sync = reinterpret_cast<GSync*>(wsm + G_MAX_KB * 32 * 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sync->a_full[b], 256);
      mbar_init(&sync->a_empty[b], 1);
      mbar_init(&sync->t_full[b], 1);
      mbar_init(&sync->t_empty[b], 4);
    }
    mbar_fence_init();
  }
  if (warp == 8) {
    tmem_alloc(&sync->tmem_base, G_TMEM_COLS);
    tmem_relinquish();
  }
  // resident weights: plain 16-byte copies (a few KB, once per CTA)
  {
    const int n16 = p.n_kb * p.Cout * 2;
    const uint4* src = reinterpret_cast<const uint4*>(p.w_packed);
    uint4* dst = reinterpret_cast<uint4*>(wsm);
    for (int i = tid; i < n16; i += G_THREADS) dst[i] = __ldg(src + i);
  }
  fence_proxy_async_smem();       // weights were written by the generic proxy, UMMA reads them (async)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sync->tmem_base;
  const int w_kb_bytes = p.Cout * 32;

  if (warp == 8) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      int i = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
        const int b = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        mbar_wait(&sync->a_full[b], par);
        mbar_wait(&sync->t_empty[b], par ^ 1u);
        tc_fence_after();
        const uint32_t a0 = smem_u32(abuf + b * G_ABUF_BYTES), w0 = smem_u32(wsm);
        for (int kb = 0; kb < p.n_kb; ++kb)
          umma_f16(tmem_base + b * 32u, umma_smem_desc(a0 + kb * G_KB_BYTES, 32),
                   umma_smem_desc(w0 + kb * w_kb_bytes, 32), p.idesc, kb > 0 ? 1u : 0u);
        umma_commit(&sync->a_empty[b]);
        umma_commit(&sync->t_full[b]);
      }
    }
    __syncwarp();
  } else {
    // ================================================================= gather (+ epilogue on warps 0-3)
    const int m = tid & 127, half = tid >> 7;
    const uint16_t* x16 = reinterpret_cast<const uint16_t*>(p.x);
    const float* x32 = reinterpret_cast<const float*>(p.x);
    uint16_t* y16 = reinterpret_cast<uint16_t*>(p.y);

    auto epilogue = [&](int i, int t) {
      if (warp >= 4) return;
      const int b = i & 1;
      mbar_wait(&sync->t_full[b], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      int pt = t;
      const int txi = pt % p.tiles_x; pt /= p.tiles_x;
      const int tyi = pt % p.tiles_y;
      const int n = pt / p.tiles_y;
      const int ox = txi * p.TW + (m & (p.TW - 1)), oy = tyi * p.TH + (m >> p.tw_shift);
      const bool valid = ox < p.OW && oy < p.OH;
      const uint32_t t_addr = tmem_base + b * 32u + ((uint32_t)(warp * 32) << 16);
      uint16_t* yp = y16 + (((size_t)n * p.OH + oy) * p.OW + ox) * p.Cout;
      for (int cb = 0; cb < p.Cout; cb += 16) {
        uint32_t v[16];
        tmem_ld16(t_addr + (uint32_t)cb, v);
        tmem_ld_wait();
        if (valid) {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = fmaf(__uint_as_float(v[2 * e]), __ldg(p.scale + cb + 2 * e), __ldg(p.shift + cb + 2 * e));
            float c = fmaf(__uint_as_float(v[2 * e + 1]), __ldg(p.scale + cb + 2 * e + 1),
                           __ldg(p.shift + cb + 2 * e + 1));
            if (p.relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
            w[e] = (uint32_t)Act<DT>::from_f32(a) | ((uint32_t)Act<DT>::from_f32(c) << 16);
          }
          uint4* o = reinterpret_cast<uint4*>(yp + cb);
          o[0] = make_uint4(w[0], w[1], w[2], w[3]);
          o[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sync->t_empty[b]);
    };

    int i = 0, t_prev = -1;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const int b = i & 1;
      mbar_wait(&sync->a_empty[b], ((uint32_t)(i >> 1) & 1u) ^ 1u);
      int pt = t;
      const int txi = pt % p.tiles_x; pt /= p.tiles_x;
      const int tyi = pt % p.tiles_y;
      const int n = pt / p.tiles_y;
      const int ox = txi * p.TW + (m & (p.TW - 1)), oy = tyi * p.TH + (m >> p.tw_shift);
      uint8_t* a = abuf + b * G_ABUF_BYTES;
      if (!p.stem) {
        // K-block = tap: 16 channels = two 16-byte chunks; this thread copies chunk `half`
        const int iy0 = oy * p.stride - 1, ix0 = ox * p.stride - 1;
        uint4 vals[9];
#pragma unroll
        for (int kb = 0; kb < 9; ++kb) {
          vals[kb] = make_uint4(0u, 0u, 0u, 0u);
          if (kb < p.n_kb) {
            const int tap = __ldg(p.kblk + kb);
            const int iy = iy0 + tap / 3, ix = ix0 + tap % 3;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
              vals[kb] = __ldg(reinterpret_cast<const uint4*>(
                                   x16 + (((size_t)n * p.H + iy) * p.W + ix) * 16) + half);
          }
        }
#pragma unroll
        for (int kb = 0; kb < 9; ++kb)
          if (kb < p.n_kb)
            *reinterpret_cast<uint4*>(a + kb * G_KB_BYTES + swz_offset((uint32_t)m, (uint32_t)half, 32)) =
                vals[kb];
      } else {
        // stem: chunk c holds k = 8c .. 8c+7 of this pixel; this thread builds chunks half, half+2, ...
        const float* xn = x32 + (size_t)n * 3 * p.H * p.W;
        for (int c = half; c < 2 * G_MAX_KB; c += 2) {
          uint32_t w[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            float f[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int lut = c_stem_lut[c * 8 + e2 * 2 + u];
              const int ci = lut & 3, iy = oy + ((lut >> 2) & 15) - 3, ix = ox + ((lut >> 6) & 15) - 3;
              f[u] = 0.f;
              if (ci < 3 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                f[u] = __ldg(xn + ((size_t)ci * p.H + iy) * p.W + ix);
            }
            w[e2] = (uint32_t)Act<DT>::from_f32(f[0]) | ((uint32_t)Act<DT>::from_f32(f[1]) << 16);
          }
          *reinterpret_cast<uint4*>(a + (c >> 1) * G_KB_BYTES + swz_offset((uint32_t)m, (uint32_t)(c & 1), 32)) =
              make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();              // im2col tile -> visible to the tensor core (async proxy)
      mbar_arrive(&sync->a_full[b]);
      if (t_prev >= 0) epilogue(i - 1, t_prev);
      t_prev = t;
    }
    if (t_prev >= 0) epilogue(i - 1, t_prev);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, G_TMEM_COLS);
  }
}

static int ilog2i(int v) { int r = 0; while ((1 << r) < v) ++r; return r; }

static int gather_launch(GatherParams& p, int act_dtype, cudaStream_t st) {
  static bool lut_done = false;
  if (!lut_done) {
    int lut[160];
    for (int k = 0; k < 160; ++k) {
      if (k < 147) { const int ci = k / 49, ky = (k % 49) / 7, kx = k % 7; lut[k] = ci | (ky << 2) | (kx << 6); }
      else lut[k] = 3;   // ci == 3: padding element
    }
    DRN_CUDA(cudaMemcpyToSymbol(c_stem_lut, lut, sizeof(lut)));
    lut_done = true;
  }
  p.TW = 32; p.TH = 4; p.tw_shift = ilog2i(p.TW);
  p.tiles_x = (p.OW + p.TW - 1) / p.TW;
  p.tiles_y = (p.OH + p.TH - 1) / p.TH;
  p.total_tiles = p.N * p.tiles_x * p.tiles_y;
  p.idesc = umma_idesc_f16(128, p.Cout, act_dtype);
  const size_t smem = 1024 + 2 * G_ABUF_BYTES + G_MAX_KB * 32 * 32 + sizeof(GSync);
  static bool attr_done[2] = {false, false};
  if (!attr_done[act_dtype]) {
    if (act_dtype == DRNB200_BF16)
      DRN_CUDA(cudaFuncSetAttribute(conv_gather_kernel<DRNB200_BF16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      DRN_CUDA(cudaFuncSetAttribute(conv_gather_kernel<DRNB200_F16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[act_dtype] = true;
  }
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_tiles, sms);
  if (grid == 0) return DRNB200_OK;
  if (act_dtype == DRNB200_BF16) conv_gather_kernel<DRNB200_BF16><<<grid, G_THREADS, smem, st>>>(p);
  else conv_gather_kernel<DRNB200_F16><<<grid, G_THREADS, smem, st>>>(p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

bool conv_gather_supported(const drnb200_conv_desc& d) {
  return d.ksize == 3 && d.Cin == 16 && d.tile_ci == 16 && d.tile_o == d.Cout &&
         (d.Cout == 16 || d.Cout == 32) && d.dilation == 1 && (d.stride == 1 || d.stride == 2) &&
         !d.has_residual && !d.out_f32;
}

int conv_gather_launch(drnb200_conv_plan* plan, cudaStream_t st) {
  const ConvParams& c = plan->p;
  GatherParams p{};
  p.x = c.x; p.y = c.y; p.w_packed = c.w_packed; p.kblk = c.kblk; p.scale = c.scale; p.shift = c.shift;
  p.n_kb = plan->h_row_ptr[1];
  p.N = c.N; p.H = c.H; p.W = c.W; p.OH = c.OH; p.OW = c.OW; p.Cout = c.Cout; p.stride = c.stride;
  p.relu = c.relu; p.stem = 0;
  if (p.n_kb == 0) {        // everything pruned: y = act(shift); the direct kernel handles that corner
    return conv_direct_launch(plan, st);
  }
  return gather_launch(p, plan->d.act_dtype, st);
}

// ----------------------------------------------------------------------------------------------- stem plan
__global__ void stem_pad_kernel(const float* __restrict__ w, float* __restrict__ wpad,
                                int32_t* __restrict__ row_ptr, int32_t* __restrict__ kblk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 16 * 160) {
    const int co = i / 160, k = i - co * 160;
    wpad[i] = k < 147 ? __ldg(w + co * 147 + k) : 0.f;
  }
  if (i < G_MAX_KB) kblk[i] = i;
  if (i == 0) { row_ptr[0] = 0; row_ptr[1] = G_MAX_KB; }
}

}  // namespace drnb200

using namespace drnb200;

struct drnb200_stem_plan {
  int N, H, W, act_dtype;
  float *d_wpad, *d_scale, *d_shift;
  int32_t *d_row_ptr, *d_kblk;
  uint16_t* d_wpacked;
};

extern "C" void drnb200_stem_plan_destroy(drnb200_stem_plan* p) {
  if (!p) return;
  cudaFree(p->d_wpad); cudaFree(p->d_scale); cudaFree(p->d_shift);
  cudaFree(p->d_row_ptr); cudaFree(p->d_kblk); cudaFree(p->d_wpacked);
  delete p;
}

extern "C" int drnb200_stem_plan_create(drnb200_stem_plan** out, const float* w_oihw,
                                        const float* bn_scale, const float* bn_shift, int N, int H,
                                        int W, int C0, int act_dtype, void* stream) {
  DRN_REQUIRE(out && w_oihw && bn_scale && bn_shift, "stem_plan_create: null pointer");
  DRN_REQUIRE(C0 == 16, "stem_plan_create: C0 must be 16 (got %d)", C0);
  DRN_REQUIRE(N > 0 && H > 0 && W > 0, "stem_plan_create: bad shape N=%d H=%d W=%d", N, H, W);
  DRN_REQUIRE(act_dtype == DRNB200_BF16 || act_dtype == DRNB200_F16, "stem_plan_create: bad act_dtype");
  drnb200_stem_plan* p = new (std::nothrow) drnb200_stem_plan();
  if (!p) { set_error("stem_plan_create: out of host memory"); return DRNB200_E_NOMEM; }
  *p = drnb200_stem_plan{};
  p->N = N; p->H = H; p->W = W; p->act_dtype = act_dtype;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes); };
  alloc((void**)&p->d_wpad, 16 * 160 * sizeof(float));
  alloc((void**)&p->d_scale, 16 * sizeof(float));
  alloc((void**)&p->d_shift, 16 * sizeof(float));
  alloc((void**)&p->d_row_ptr, 2 * sizeof(int32_t));
  alloc((void**)&p->d_kblk, G_MAX_KB * sizeof(int32_t));
  alloc((void**)&p->d_wpacked, 16 * 160 * 2);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_scale, bn_scale, 16 * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->d_shift, bn_shift, 16 * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) { drnb200_stem_plan_destroy(p); return cuda_fail(e, "stem plan setup"); }
  stem_pad_kernel<<<(16 * 160 + 255) / 256, 256, 0, st>>>(w_oihw, p->d_wpad, p->d_row_ptr, p->d_kblk);
  int rc = drnb200_pack_weights(p->d_wpad, nullptr, 16, 160, 1, 1, 16, 16, p->d_row_ptr, p->d_kblk,
                                act_dtype, p->d_wpacked, stream);
  if (rc == DRNB200_OK) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize(stem plan)");
  }
  if (rc != DRNB200_OK) { drnb200_stem_plan_destroy(p); return rc; }
  *out = p;
  return DRNB200_OK;
}

extern "C" int drnb200_stem_plan_forward(drnb200_stem_plan* plan, const float* x_nchw, void* y_nhwc,
                                         void* stream) {
  DRN_REQUIRE(plan && x_nchw && y_nhwc, "stem_plan_forward: null pointer");
  GatherParams p{};
  p.x = x_nchw; p.y = y_nhwc; p.w_packed = reinterpret_cast<const uint8_t*>(plan->d_wpacked);
  p.kblk = plan->d_kblk; p.scale = plan->d_scale; p.shift = plan->d_shift; p.n_kb = G_MAX_KB;
  p.N = plan->N; p.H = plan->H; p.W = plan->W; p.OH = plan->H; p.OW = plan->W; p.Cout = 16;
  p.stride = 1; p.relu = 1; p.stem = 1;
  return gather_launch(p, plan->act_dtype, (cudaStream_t)stream);
}
