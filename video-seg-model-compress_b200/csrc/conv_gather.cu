// conv_gather.cu — small-Cin convolutions on tcgen05 with an im2col tile built by the CTA's own threads.
//
// The TMA path of conv_tc.cu issues one box per filter tap; with 16 input channels a tap is a 32-byte row
// and the kernel becomes TMA-issue / L2-request bound (measured: 2.9 ms for DRN-D-22 layer1 at batch 8,
// 17x its HBM floor).  Here ONE TMA box per tile brings the input halo (tile + filter apron, zero-filled
// outside the image = the conv padding) into shared memory, the CTA's threads expand it into the
// 128-pixel x K im2col tile directly in the UMMA SWIZZLE_32B K-major layout (smem -> smem, no global
// re-reads), and the weights stay resident in shared memory for the whole persistent CTA:
//   MODE 3x3 : x NHWC 16-bit with exactly 16 channels, K-block = one tap (drn.py:201-211 layer1/layer2)
//   MODE stem: x NCHW float32 with 3 channels, 7x7, K = 147 padded to 160; the fp32 -> 16-bit conversion of
//              the frame is fused into the expansion (drn.py:132-137; callers pass fp32 NCHW, semantic_seg.py:440)
// GEMM orientation: M = 128 output pixels (TMEM lanes), N = Cout (16 or 32 columns), K-step = 16.
// Roles (832 threads): warps 0-15 expand (512 threads; the expansion is latency/issue bound, so it gets
// most of the CTA's warps), warp 16 allocates TMEM and issues the MMAs, warp 17 issues the halo TMA loads
// up to G_HRING tiles ahead, warps 18-25 run the epilogue (two groups of four warps on alternate tiles).
// A ring of halo buffers, two im2col buffers and two TMEM accumulators pipeline
//   TMA(i+k) | expand(i+1) | MMA(i) | epilogue(i-1)   with every stage on its own warps.
#include "conv_internal.cuh"
#include <algorithm>
#include <cudaTypedefs.h>
#include <new>

namespace drnb200 {

constexpr int G_EXP_WARPS = 16;         // expanding warps
constexpr int G_W_MMA = G_EXP_WARPS, G_W_TMA = G_EXP_WARPS + 1, G_W_EPI = G_EXP_WARPS + 2;
constexpr int G_THREADS = (G_EXP_WARPS + 2 + 8) * 32;
constexpr int G_MAX_KB = 10;            // 9 taps, or 160/16 stem K-blocks
constexpr int G_KB_BYTES = 128 * 32;    // one K-block of the im2col tile: 128 pixels x 16 elements
constexpr int G_MAX_ABUF = 4;            // im2col buffers: expand(i+k) | MMA(i) — two were latency-bound (measured)
constexpr int G_ACC = 4;                // TMEM accumulator stages (MMA -> epilogue round trip >> tile time)
constexpr uint32_t G_TMEM_COLS = 128;   // 4 accumulators x 32 columns
constexpr int G_STEM_TW = 32, G_STEM_TH = 4;                // stem output tile: 32 x 4 pixels
constexpr int G_STEM_HW = 40, G_STEM_HH = G_STEM_TH + 6;    // stem halo: 3 planes x 10 rows x 40 floats
constexpr int G_HRING = 6;              // max halo ring depth: TMA latency (~2 us) >> per-tile time (~0.3 us)

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
  int TW, TH, tw_shift;      // output tile (32x4; 16x8 for stride 2 so that the halo row fits one TMA box row)
  int tiles_x, tiles_y, total_tiles;
  uint32_t magic_x, magic_y; // ceil(2^32 / tiles_{x,y}) for the division-free tile decode
  int halo_w, halo_h;        // 3x3: halo box in pixels
  uint32_t halo_bytes;       // TMA transaction bytes of one halo box
  uint32_t halo_stride;      // halo buffer pitch in shared memory (halo_bytes rounded up to 1 KB)
  uint32_t abuf_bytes;       // one im2col buffer: n_kb x 4 KB
  int n_abuf, ring;          // buffers that fit in shared memory
  uint32_t idesc;
};

struct __align__(16) GSync {
  alignas(16) float scale[32];
  alignas(16) float shift[32];
  uint64_t h_full[G_HRING], h_empty[G_HRING], a_full[G_MAX_ABUF], a_empty[G_MAX_ABUF],
      t_full[G_ACC], t_empty[G_ACC];
  uint32_t tmem_base, pad;
};

struct GTile { int n, ox0, oy0; };
__device__ __forceinline__ GTile g_decode(const GatherParams& p, int t) {
  GTile c;
  const int q1 = p.tiles_x == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_x);      // t / tiles_x
  const int txi = t - q1 * p.tiles_x;
  c.n = p.tiles_y == 1 ? q1 : (int)__umulhi((uint32_t)q1, p.magic_y);             // q1 / tiles_y
  const int tyi = q1 - c.n * p.tiles_y;
  c.ox0 = txi * p.TW; c.oy0 = tyi * p.TH;
  return c;
}

// halo float offset of stem element k (k = ci*49 + ky*7 + kx); the halo origin is (ox0-4, oy0-3)
__host__ __device__ constexpr int stem_halo_off(int k) {
  return ((k / 49) * G_STEM_HH + (k % 49) / 7) * G_STEM_HW + (k % 7) + 1;
}

// chunks PART, PART+4, ... of one pixel's 160-element im2col row, offsets resolved at compile time
template <int DT, int PART>
__device__ __forceinline__ void stem_expand(const float* hp, uint8_t* a, int m) {
#pragma unroll
  for (int cc = 0; cc < G_MAX_KB / 2; ++cc) {
    const int c = 4 * cc + PART;
    uint32_t w[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      const int k0 = c * 8 + e2 * 2, k1 = k0 + 1;
      const float f0 = k0 < 147 ? hp[stem_halo_off(k0)] : 0.f;
      const float f1 = k1 < 147 ? hp[stem_halo_off(k1)] : 0.f;
      w[e2] = (uint32_t)Act<DT>::from_f32(f0) | ((uint32_t)Act<DT>::from_f32(f1) << 16);
    }
    *reinterpret_cast<uint4*>(a + (c >> 1) * G_KB_BYTES + swz_offset((uint32_t)m, (uint32_t)(c & 1), 32)) =
        make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <int DT>
__global__ void __launch_bounds__(G_THREADS, 1)
conv_gather_kernel(const __grid_constant__ CUtensorMap tmap_x, const GatherParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* abuf = smem;                                   // n_abuf x abuf_bytes
  uint8_t* halo = smem + (size_t)p.n_abuf * p.abuf_bytes; // ring x halo_stride
  uint8_t* wsm = halo + (size_t)p.ring * p.halo_stride;   // n_kb x Cout x 32 B
  GSync* sync = reinterpret_cast<GSync*>(wsm + G_MAX_KB * 32 * 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch();
  if (tid == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int b = 0; b < G_HRING; ++b) {
      mbar_init(&sync->h_full[b], 1);
      mbar_init(&sync->h_empty[b], G_EXP_WARPS);    // one arrive per expanding warp
    }
    for (int b = 0; b < G_MAX_ABUF; ++b) {
      mbar_init(&sync->a_full[b], G_EXP_WARPS);
      mbar_init(&sync->a_empty[b], 1);
    }
    for (int b = 0; b < G_ACC; ++b) {
      mbar_init(&sync->t_full[b], 1);
      mbar_init(&sync->t_empty[b], 4);
    }
    mbar_fence_init();
  }
  if (warp == G_W_MMA) {
    tmem_alloc(&sync->tmem_base, G_TMEM_COLS);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  if (tid < 32) {   // BN affine -> shared memory: with ~all of the SM's storage carved out as shared memory
                    // there is no L1 left, and 32 global loads per pixel thread per tile would go to L2
    sync->scale[tid] = tid < p.Cout ? __ldg(p.scale + tid) : 0.f;
    sync->shift[tid] = tid < p.Cout ? __ldg(p.shift + tid) : 0.f;
  }
  {  // resident weights: plain 16-byte copies (a few KB, once per CTA)
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

  if (warp == G_W_TMA) {
    // ================================================================= halo TMA producer (warp-uniform loop)
    {
      int b = 0;
      uint32_t bph = 0;                       // ring slot / phase kept incrementally (no div/mod per tile)
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const GTile c = g_decode(p, t);
        mbar_wait(&sync->h_empty[b], bph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&sync->h_full[b], p.halo_bytes);
          if (p.stem)   // tensor {W, H, 3, N} fp32, box {40, 10, 3, 1}; x origin ox0-4 keeps the box 16-byte aligned
            tma_load_4d(&tmap_x, &sync->h_full[b], halo + (size_t)b * p.halo_stride, c.ox0 - 4, c.oy0 - 3, 0, c.n);
          else          // tensor {W*4, H, N, 1} of 8-byte elements (a pixel = 16 ch x 2 B = 4 elements), box
                        // {halo_w*4, halo_h, 1, 1}: one box row per halo row
            tma_load_4d(&tmap_x, &sync->h_full[b], halo + (size_t)b * p.halo_stride, (c.ox0 * p.stride - 1) * 4,
                        c.oy0 * p.stride - 1, c.n, 0);
        }
        __syncwarp();
        if (++b == p.ring) { b = 0; bph ^= 1u; }
      }
    }
  } else if (warp == G_W_MMA) {
    // ================================================================= MMA issuer (warp-uniform loop)
    {
      const uint64_t d_hi = umma_smem_desc(0u, 32);
      const uint32_t w16 = smem_u32(wsm) >> 4, wk16 = (uint32_t)w_kb_bytes >> 4;
      const int n_kb = p.n_kb;
      int i = 0, b = 0;
      uint32_t bph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
        const int ta = i % G_ACC;
        mbar_wait(&sync->a_full[b], bph);
        mbar_wait(&sync->t_empty[ta], ((uint32_t)(i / G_ACC) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t a16 = smem_u32(abuf + (size_t)b * p.abuf_bytes) >> 4;
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < G_MAX_KB; ++kb)
            if (kb < n_kb)
              umma_f16(tmem_base + ta * 32u, d_hi | (uint64_t)(a16 + kb * (G_KB_BYTES >> 4)),
                       d_hi | (uint64_t)(w16 + kb * wk16), p.idesc, kb > 0 ? 1u : 0u);
          umma_commit(&sync->a_empty[b]);
          umma_commit(&sync->t_full[ta]);
        }
        __syncwarp();
        if (++b == p.n_abuf) { b = 0; bph ^= 1u; }
      }
    }
  } else if (warp >= G_W_EPI) {
    // ================================================================= epilogue (warps 10..13)
    uint16_t* y16 = reinterpret_cast<uint16_t*>(p.y);
    const int q = warp & 3;                                // TMEM lane quarter of this warp
    const int grp = (warp - G_W_EPI) >> 2;                 // epilogue group 0/1 takes alternate tiles
    const int m = q * 32 + lane;                           // TMEM lane = pixel of the tile
    for (int i = grp, t = blockIdx.x + grp * gridDim.x; t < p.total_tiles; t += 2 * gridDim.x, i += 2) {
      const int b = i % G_ACC;
      mbar_wait(&sync->t_full[b], (uint32_t)(i / G_ACC) & 1u);
      tc_fence_after();
      const GTile c = g_decode(p, t);
      const int ox = c.ox0 + (m & (p.TW - 1)), oy = c.oy0 + (m >> p.tw_shift);
      const bool valid = ox < p.OW && oy < p.OH;
      const uint32_t t_addr = tmem_base + b * 32u + ((uint32_t)(q * 32) << 16);
      uint16_t* yp = y16 + (((size_t)c.n * p.OH + oy) * p.OW + ox) * p.Cout;
      for (int cb = 0; cb < p.Cout; cb += 16) {
        uint32_t v[16];
        tmem_ld16(t_addr + (uint32_t)cb, v);
        tmem_ld_wait();
        if (valid) {
          uint32_t w[8];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 sc = *reinterpret_cast<const float4*>(&sync->scale[cb + 4 * e4]);
            const float4 sh = *reinterpret_cast<const float4*>(&sync->shift[cb + 4 * e4]);
            float a0 = fmaf(__uint_as_float(v[4 * e4]), sc.x, sh.x), a1 = fmaf(__uint_as_float(v[4 * e4 + 1]), sc.y, sh.y);
            float a2 = fmaf(__uint_as_float(v[4 * e4 + 2]), sc.z, sh.z), a3 = fmaf(__uint_as_float(v[4 * e4 + 3]), sc.w, sh.w);
            if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
            w[2 * e4] = pack2<DT>(a0, a1);
            w[2 * e4 + 1] = pack2<DT>(a2, a3);
          }
          stg256(yp + cb, w);                    // 16 channels = one full 32-byte sector per lane
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sync->t_empty[b]);
    }
  } else {
    // ================================================================= expand
    // 3x3: byte offset of every live tap inside the halo (constant for the whole kernel)
    int tap_off[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int kb = 2 * j + (tid >> 8);
      tap_off[j] = 0;
      if (!p.stem && kb < p.n_kb) {
        const int tap = __ldg(p.kblk + kb);
        tap_off[j] = ((tap / 3) * p.halo_w + (tap % 3)) * 32;
      }
    }
    int b = 0, hb = 0;
    uint32_t bph = 0, hph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(&sync->h_full[hb], hph);          // halo of this tile has landed
      mbar_wait(&sync->a_empty[b], bph ^ 1u);     // the MMAs that read this im2col buffer have retired
      const uint8_t* h = halo + (size_t)hb * p.halo_stride;
      uint8_t* a = abuf + (size_t)b * p.abuf_bytes;
      if (!p.stem) {
        // thread = (pixel m, 16-byte half of its 16 channels, tap parity); a K-block is one filter tap
        const int m = (tid & 255) >> 1, half = tid & 1, tg = tid >> 8;
        const int hx = (m & (p.TW - 1)) * p.stride, hy = (m >> p.tw_shift) * p.stride;
        const uint8_t* src0 = h + ((size_t)hy * p.halo_w + hx) * 32 + half * 16;
        uint8_t* dst0 = a + swz_offset((uint32_t)m, (uint32_t)half, 32);
        uint4 v[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int kb = 2 * j + tg;
          if (kb < p.n_kb) v[j] = *reinterpret_cast<const uint4*>(src0 + tap_off[j]);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int kb = 2 * j + tg;
          if (kb < p.n_kb) *reinterpret_cast<uint4*>(dst0 + kb * G_KB_BYTES) = v[j];
        }
      } else {
        // stem: thread = (pixel m, chunk parity); chunk c holds k = 8c .. 8c+7 (k = ci*49 + ky*7 + kx)
        const int m = tid & 127;
        const float* hp = reinterpret_cast<const float*>(h) + (m / G_STEM_TW) * G_STEM_HW + (m & (G_STEM_TW - 1));
        switch (tid >> 7) {                       // warp-uniform: 4 thread groups x 5 chunks each
          case 0: stem_expand<DT, 0>(hp, a, m); break;
          case 1: stem_expand<DT, 1>(hp, a, m); break;
          case 2: stem_expand<DT, 2>(hp, a, m); break;
          default: stem_expand<DT, 3>(hp, a, m); break;
        }
      }
      fence_proxy_async_smem();              // im2col tile -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sync->a_full[b]);
        mbar_arrive(&sync->h_empty[hb]);     // halo buffer may be refilled
      }
      if (++b == p.n_abuf) { b = 0; bph ^= 1u; }
      if (++hb == p.ring) { hb = 0; hph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == G_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, G_TMEM_COLS);
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
  return fn;
}

// cached halo tensor map per (pointer, geometry)
struct GMapCache {
  const void* ptr = nullptr;
  int N = 0, H = 0, W = 0, stride = 0, stem = -1, dt = -1;
  CUtensorMap map;
};

static int gather_launch(GatherParams& p, int act_dtype, GMapCache& cache, cudaStream_t st) {
  if (p.stem || p.stride == 1) { p.TW = 32; p.TH = 4; p.tw_shift = 5; }
  else { p.TW = 16; p.TH = 8; p.tw_shift = 4; }
  p.tiles_x = (p.OW + p.TW - 1) / p.TW;
  p.tiles_y = (p.OH + p.TH - 1) / p.TH;
  p.total_tiles = p.N * p.tiles_x * p.tiles_y;
  if ((uint64_t)p.total_tiles * (uint64_t)std::max(p.tiles_x, p.tiles_y) >= (1ull << 32)) {
    set_error("conv_gather: problem too large for the 32-bit tile decode");
    return DRNB200_E_ARG;
  }
  p.magic_x = p.tiles_x == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_x - 1) / p.tiles_x);
  p.magic_y = p.tiles_y == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_y - 1) / p.tiles_y);
  p.idesc = umma_idesc_f16(128, p.Cout, act_dtype);
  if (p.stem) {
    p.halo_w = G_STEM_HW; p.halo_h = G_STEM_HH;
    p.halo_bytes = 3 * G_STEM_HH * G_STEM_HW * 4;
  } else {
    p.halo_w = (p.TW - 1) * p.stride + 3; p.halo_h = (p.TH - 1) * p.stride + 3;
    p.halo_bytes = (uint32_t)p.halo_w * p.halo_h * 32;
  }
  if (cache.ptr != p.x || cache.N != p.N || cache.H != p.H || cache.W != p.W || cache.stride != p.stride ||
      cache.stem != p.stem || cache.dt != act_dtype) {
    auto fn = g_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    CUresult r;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (p.stem) {
      cuuint64_t gdim[4] = {(cuuint64_t)p.W, (cuuint64_t)p.H, 3, (cuuint64_t)p.N};
      cuuint64_t gstr[3] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.W * p.H * 4, (cuuint64_t)p.W * p.H * 12};
      cuuint32_t box[4] = {(cuuint32_t)G_STEM_HW, (cuuint32_t)G_STEM_HH, 3, 1};
      r = fn(&cache.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(p.x), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gdim[4] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H, (cuuint64_t)p.N, 1};
      cuuint64_t gstr[3] = {(cuuint64_t)p.W * 32, (cuuint64_t)p.W * p.H * 32, (cuuint64_t)p.W * p.H * 32 * p.N};
      cuuint32_t box[4] = {(cuuint32_t)p.halo_w * 4, (cuuint32_t)p.halo_h, 1, 1};
      r = fn(&cache.map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(p.x), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(halo) failed with CUresult %d (stem=%d W=%d H=%d N=%d)", (int)r, p.stem,
                p.W, p.H, p.N);
      return DRNB200_E_CUDA;
    }
    cache.ptr = p.x; cache.N = p.N; cache.H = p.H; cache.W = p.W; cache.stride = p.stride; cache.stem = p.stem;
    cache.dt = act_dtype;
  }
  p.halo_stride = (p.halo_bytes + 1023u) & ~1023u;
  p.abuf_bytes = (uint32_t)p.n_kb * G_KB_BYTES;
  const size_t kMaxSmem = 232448;
  const size_t fixed = 1024 + G_MAX_KB * 32 * 32 + sizeof(GSync);
  static const char* env_ab = getenv("DRNB200_G_ABUF");      // A/B knob: number of im2col buffers (2..4)
  p.n_abuf = env_ab ? std::max(2, std::min(G_MAX_ABUF, atoi(env_ab))) : 3;
  p.ring = (int)std::min<size_t>(G_HRING, (kMaxSmem - fixed - (size_t)p.n_abuf * p.abuf_bytes) / p.halo_stride);
  if (p.ring < 2) { set_error("conv_gather: buffers do not fit shared memory"); return DRNB200_E_ARG; }
  const size_t smem = fixed + (size_t)p.n_abuf * p.abuf_bytes + (size_t)p.ring * p.halo_stride;
  static std::atomic<unsigned long long> attr_done[2];
  if (attr_needed_on_this_device(attr_done[act_dtype])) {
    if (act_dtype == DRNB200_BF16)
      DRN_CUDA(cudaFuncSetAttribute(conv_gather_kernel<DRNB200_BF16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    else
      DRN_CUDA(cudaFuncSetAttribute(conv_gather_kernel<DRNB200_F16>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
  }
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_tiles, sms);
  if (grid == 0) return DRNB200_OK;
  if (act_dtype == DRNB200_BF16) launch_chained(conv_gather_kernel<DRNB200_BF16>, grid, G_THREADS, smem, st, cache.map, p);
  else launch_chained(conv_gather_kernel<DRNB200_F16>, grid, G_THREADS, smem, st, cache.map, p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

bool conv_gather_supported(const drnb200_conv_desc& d) {
  static const char* env = getenv("DRNB200_GATHER");   // A/B knob: 0 = route 16-channel layers elsewhere
  if (env && env[0] == '0') return false;
  return d.ksize == 3 && d.Cin == 16 && d.tile_ci == 16 && d.tile_o == d.Cout &&
         (d.Cout == 16 || d.Cout == 32) && d.dilation == 1 && (d.stride == 1 || d.stride == 2) &&
         !d.has_residual && !d.out_f32 && (d.x_cpitch == 0 || d.x_cpitch == d.Cin) && d.relu_n == 0;
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
  static_assert(sizeof(GMapCache) <= sizeof(plan->gather_cache), "gather cache storage too small");
  GMapCache* cache = reinterpret_cast<GMapCache*>(plan->gather_cache);
  if (!plan->gather_cache_init) { new (cache) GMapCache(); plan->gather_cache_init = true; }
  return gather_launch(p, plan->d.act_dtype, *cache, st);
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
  drnb200::GMapCache cache;
  drnb200::StemTxState* tx;      // Toeplitz-weight kernel (default); null -> im2col-gather kernel
  int N, H, W, act_dtype;
  float *d_wpad, *d_scale, *d_shift;
  int32_t *d_row_ptr, *d_kblk;
  uint16_t* d_wpacked;
};

extern "C" void drnb200_stem_plan_destroy(drnb200_stem_plan* p) {
  if (!p) return;
  cudaFree(p->d_wpad); cudaFree(p->d_scale); cudaFree(p->d_shift);
  cudaFree(p->d_row_ptr); cudaFree(p->d_kblk); cudaFree(p->d_wpacked);
  stem_tx_destroy(p->tx);
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
  static const char* env = getenv("DRNB200_STEM");      // A/B knob: "gather" keeps the im2col-gather stem
  p->tx = nullptr;
  if (rc == DRNB200_OK && !(env && env[0] == 'g')) rc = stem_tx_create(&p->tx, w_oihw, act_dtype, st);
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
  if (plan->tx)
    return stem_tx_forward(plan->tx, x_nchw, 0, nullptr, 0, y_nhwc, plan->d_scale, plan->d_shift, plan->N, plan->H,
                           plan->W, plan->act_dtype, (cudaStream_t)stream);
  GatherParams p{};
  p.x = x_nchw; p.y = y_nhwc; p.w_packed = reinterpret_cast<const uint8_t*>(plan->d_wpacked);
  p.kblk = plan->d_kblk; p.scale = plan->d_scale; p.shift = plan->d_shift; p.n_kb = G_MAX_KB;
  p.N = plan->N; p.H = plan->H; p.W = plan->W; p.OH = plan->H; p.OW = plan->W; p.Cout = 16;
  p.stride = 1; p.relu = 1; p.stem = 1;
  return gather_launch(p, plan->act_dtype, plan->cache, (cudaStream_t)stream);
}

extern "C" int drnb200_stem_plan_forward_u8(drnb200_stem_plan* plan, const uint8_t* frames_nhwc,
                                            const uint16_t* lut, int bgr, void* y_nhwc, void* stream) {
  DRN_REQUIRE(plan && frames_nhwc && lut && y_nhwc, "stem_plan_forward_u8: null pointer");
  DRN_REQUIRE(plan->tx, "stem_plan_forward_u8: needs the Toeplitz stem (unset DRNB200_STEM=gather)");
  return stem_tx_forward(plan->tx, frames_nhwc, 1, lut, bgr ? 1 : 0, y_nhwc, plan->d_scale, plan->d_shift, plan->N,
                         plan->H, plan->W, plan->act_dtype, (cudaStream_t)stream);
}
