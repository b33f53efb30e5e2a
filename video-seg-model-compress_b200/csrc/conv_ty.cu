// conv_ty.cu — 3x3 stride-1 convolution 16 -> 16 channels (DRN layer1, full resolution) with the y direction of
// the filter folded into the WEIGHT operand ("Toeplitz along y") and the input read as PIXEL PAIRS.
//
// Why: conv_halo.cu runs this layer as 9 shifted-window MMAs per 128 pixels, i.e. every pixel's 32 bytes are read
// from shared memory nine times as the A operand (ncu, profiles/r02_ncu_front_kernels.txt: 42.5 M operand wavefronts
// = 5.4 GB per 8 frames, the tensor-core read port at 73 % of its peak, 0.225 ms against an HBM floor of 0.166 ms).
// Version 1 of this file kept one pixel per operand row (32-byte rows, SWIZZLE_32B) and folded only the filter rows:
// 0.209-0.214 ms — half the operand reads, but a TMA box of 32-byte rows is bound by the number of rows it moves
// (1360 rows per 43 KB halo), and one MMA of N <= 48 per 128 pixels and tap column is bound by the MMA count.
// Here the input is the tensor {32, W/2, H, N} of pixel pairs (64-byte rows, SWIZZLE_64B, as in conv_s2.cu):
//   a tile is 128 pairs = 256 pixels of a row (UMMA M = pair r) x 8 output rows; accumulator columns are
//   (output row yo, pixel-of-pair s, cout) = 8 x 32, so that a TMEM lane holds the 64 contiguous output bytes of a pair;
//   output pixel 2r + s, tap kx reads input pixel 2r + s + kx - 1, which is one of FOUR operand windows
//     a = 0: pair r - 1, odd pixel   -> (s=0, kx=0)               a = 1: pair r, even pixel     -> (s=0, kx=1), (s=1, kx=0)
//     a = 2: pair r, odd pixel       -> (s=0, kx=2), (s=1, kx=1)  a = 3: pair r + 1, even pixel -> (s=1, kx=2)
//   (a descriptor start of +32 bytes inside the 64-byte row selects the odd pixel);
//   input row e of the halo (rows y0-1 .. y0+8) feeds the up to three output rows yo = e - ky, so the B operand of
//   window a is the stack over (ky = 2, 1, 0) x (s = 0, 1) of the 16 x 16 tap tiles named above (zeros where a window
//   has no tap for that s), 96 x 16, or a 32/64-row slice of it at the top and bottom of the tile.
// 47 MMAs (M = 128, N = 32/64/96, K = 16) per 2048 pixels instead of 74 of N <= 48, and half as many TMA rows.
// An output row's column block is first written by its own ky = 0 MMA of window 0 (accumulate off; its zero s = 1 half
// clears the other pixel), which is why that MMA is issued separately from the ky = 1, 2 slice of the same input row.
// Roles (352 threads): warp 0 halo TMA producer, warps 1-2 MMA issue on alternate tiles (warp 1 allocates TMEM),
// warps 3-10 epilogue (two groups on alternate tiles; thread = pixel pair, BN affine + ReLU, 2 x 32-byte stores per row).
// Barriers: tile i of a CTA uses halo slot i % 2 and accumulator i % 2, but barrier i % 8 of each kind.  With one
// barrier per slot, an MMA warp could poll "slot 0, next fill" while the CURRENT fill of slot 0 was still in flight
// (another warp had not consumed it yet and TMA boxes complete out of order once the input comes from HBM):
// mbarrier.try_wait.parity on a barrier that is two phases behind answers "done".  Measured with version 1 (four slots,
// three MMA warps): 1-2 % of the launches of two 1024x2048 frames faulted that way.  More barriers than resources keep
// every waiter within one phase of its barrier whatever the interleaving (the look-ahead is bounded by the slots), and
// every barrier's successive phases are awaited by the same warp.
#include "conv_internal.cuh"
#include <algorithm>
#include <cudaTypedefs.h>
#include <new>

namespace drnb200 {

constexpr int TY_W = 256, TY_H = 8;            // output tile (pixels x rows); 128 pixel pairs = UMMA M
constexpr int TY_HR = TY_H + 2;                // halo rows
constexpr int TY_HP = 136;                     // pixel pairs per halo row: 130 needed (pair -1 .. pair 128), rounded up so
                                               // that a row is a multiple of the 512-byte SWIZZLE_64B period
constexpr uint32_t TY_PAIR = 64;               // bytes per pixel pair (2 x 16 channels x 16 bit)
constexpr uint32_t TY_ROWB = TY_HP * TY_PAIR;  // 8704 = 17 x 512
constexpr uint32_t TY_HALO_TX = TY_HR * TY_ROWB;
constexpr uint32_t TY_SLOT = (TY_HALO_TX + 1023u) & ~1023u;
constexpr uint32_t TY_WTAP = 16 * 32;          // one tap: 16 couts x 16 cin x 16 bit
constexpr uint32_t TY_WWIN = 6 * TY_WTAP;      // per window: [ky=2: s0, s1; ky=1: s0, s1; ky=0: s0, s1]
constexpr uint32_t TY_WBYTES = (4 * TY_WWIN + 1023u) & ~1023u;
constexpr int TY_RING = 2;
constexpr int TY_MMA_WARPS = 2;
constexpr int TY_ACC = 2;                      // TMEM accumulators of 256 columns (8 rows x 2 pixels x 16 couts)
constexpr int TY_NBAR = 8;
constexpr int TY_EPI_GROUPS = 2;
constexpr int TY_W_EPI = 1 + TY_MMA_WARPS;
constexpr int TY_THREADS = (TY_W_EPI + 4 * TY_EPI_GROUPS) * 32;
static_assert(TY_ACC % TY_MMA_WARPS == 0 && TY_RING % TY_MMA_WARPS == 0 && TY_ACC % TY_EPI_GROUPS == 0 &&
              TY_NBAR % TY_ACC == 0 && TY_NBAR % TY_RING == 0, "slot / accumulator / barrier indices are masks of the tile index");

struct TyParams {
  const void* x;
  void* y;
  const uint8_t* w_packed;     // live taps only, 512 bytes each (pack_weights, tile 16 x 16, SWIZZLE_32B rows)
  const int32_t* kblk;         // tap index ky*3+kx of every packed tile
  const float* scale;
  const float* shift;
  int n_kb, N, H, W, relu_n;
  int tiles_x, tiles_y, total_tiles;
  uint32_t magic_x, magic_y;
  uint32_t idesc[3];           // N = 32, 64, 96
};

struct __align__(16) TySync {
  uint64_t h_full[TY_NBAR], h_empty[TY_NBAR], t_full[TY_NBAR], t_empty[TY_NBAR];
  uint32_t tmem_base, pad[3];
  alignas(16) float scale[32];   // (s, cout): the 16 BN factors twice, in accumulator-column order
  alignas(16) float shift[32];
};

struct TyTile { int n, ox0, oy0; };
__device__ __forceinline__ TyTile ty_decode(const TyParams& p, int t) {
  TyTile c;
  const int q1 = p.tiles_x == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_x);
  const int txi = t - q1 * p.tiles_x;
  c.n = p.tiles_y == 1 ? q1 : (int)__umulhi((uint32_t)q1, p.magic_y);
  const int tyi = q1 - c.n * p.tiles_y;
  c.ox0 = txi * TY_W; c.oy0 = tyi * TY_H;
  return c;
}

// K-major operand descriptor without the start address: 8-row groups `sbo` bytes apart, layout 4 = SWIZZLE_64B,
// 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t ty_desc_hi(uint32_t sbo, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// tap column kx that window a contributes to pixel-of-pair s (-1: none)
__host__ __device__ constexpr int ty_kx(int a, int s) { return s == 0 ? (a <= 2 ? a : -1) : (a >= 1 ? a - 1 : -1); }

template <int DT>
__global__ void __launch_bounds__(TY_THREADS, 1)
conv_ty_kernel(const __grid_constant__ CUtensorMap tmap_x, const TyParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* halo = smem;                                    // TY_RING x TY_SLOT
  uint8_t* wsm = smem + (size_t)TY_RING * TY_SLOT;         // 4 windows x [96 rows][32 B]
  TySync* sync = reinterpret_cast<TySync*>(wsm + TY_WBYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch();
  if (tid == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int b = 0; b < TY_NBAR; ++b) {
      mbar_init(&sync->h_full[b], 1); mbar_init(&sync->h_empty[b], 1);
      mbar_init(&sync->t_full[b], 1); mbar_init(&sync->t_empty[b], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sync->tmem_base, TY_ACC * 256);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  // resident weights: block (a, j = 2 - ky, s) <- the packed tile of tap ky*3 + ty_kx(a, s); zeros when the window has
  // no tap for s or the tap is pruned
  for (int it = tid; it < 24 * 32; it += TY_THREADS) {
    const int blk = it >> 5, part = it & 31;
    const int a = blk / 6, js = blk - a * 6, ky = 2 - (js >> 1), s = js & 1;
    const int kx = ty_kx(a, s);
    int idx = -1;
    if (kx >= 0)
      for (int k = 0; k < p.n_kb; ++k)
        if (__ldg(p.kblk + k) == ky * 3 + kx) idx = k;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (idx >= 0) v = __ldg(reinterpret_cast<const uint4*>(p.w_packed + (size_t)idx * TY_WTAP) + part);
    reinterpret_cast<uint4*>(wsm + (size_t)blk * TY_WTAP)[part] = v;
  }
  if (tid < 32) { sync->scale[tid] = __ldg(p.scale + (tid & 15)); sync->shift[tid] = __ldg(p.shift + (tid & 15)); }
  fence_proxy_async_smem();       // written by the generic proxy, read by UMMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sync->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform loop)
    int i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const TyTile c = ty_decode(p, t);
      // slot i % 2 was read by the MMAs of tile i - 2, which commit to barrier (i - 2) % 8
      if (i >= TY_RING) mbar_wait(&sync->h_empty[(i - TY_RING) & (TY_NBAR - 1)], (uint32_t)((i - TY_RING) / TY_NBAR) & 1u);
      if (elect_one()) {
        uint64_t* full = &sync->h_full[i & (TY_NBAR - 1)];
        mbar_arrive_expect_tx(full, TY_HALO_TX);
        // tensor {32, W/2, H, N}; box {32, 136, 10, 1}; zero fill outside the image = the conv padding
        tma_load_4d(&tmap_x, full, halo + (size_t)(i & (TY_RING - 1)) * TY_SLOT, 0, c.ox0 / 2 - 1, c.oy0 - 1, c.n);
      }
      __syncwarp();
    }
  } else if (warp < TY_W_EPI) {
    // ===================================================================== MMA issuers (warp-uniform loop)
    const uint64_t a_hi = ty_desc_hi(8u * TY_PAIR, 4);     // A: 64-byte rows, SWIZZLE_64B
    const uint64_t b_hi = ty_desc_hi(8u * 32u, 6);         // B: 32-byte rows, SWIZZLE_32B
    const uint32_t w16 = smem_u32(wsm) >> 4;
    const int mw = warp - 1;
    int i = mw;
    for (int t = blockIdx.x + mw * gridDim.x; t < p.total_tiles; t += TY_MMA_WARPS * gridDim.x, i += TY_MMA_WARPS) {
      const int acc = i & (TY_ACC - 1), b = i & (TY_RING - 1), bar = i & (TY_NBAR - 1);
      mbar_wait(&sync->h_full[bar], (uint32_t)(i / TY_NBAR) & 1u);
      // the accumulator was read out by the epilogue of tile i - 2, which arrives on barrier (i - 2) % 8
      if (i >= TY_ACC) mbar_wait(&sync->t_empty[(i - TY_ACC) & (TY_NBAR - 1)], (uint32_t)((i - TY_ACC) / TY_NBAR) & 1u);
      tc_fence_after();
      const uint32_t h16 = smem_u32(halo + (size_t)b * TY_SLOT) >> 4;
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
      if (elect_one()) {
        // input row e through window a -> output rows yo_lo .. yo_lo + nblk - 1 (32 columns each) through the rows
        // of window a's weight stack that start at block row j0 (stack of ky = 2, 1, 0: output row yo takes ky = e - yo)
        auto mma = [&](int e, int a, int yo_lo, int nblk, int j0, uint32_t accumulate) {
          umma_f16(d_tmem + (uint32_t)(yo_lo * 32),
                   a_hi | (uint64_t)(h16 + ((uint32_t)e * TY_ROWB + 32u * (uint32_t)(a + 1)) / 16u),
                   b_hi | (uint64_t)(w16 + ((uint32_t)a * TY_WWIN + (uint32_t)j0 * 2u * TY_WTAP) / 16u),
                   p.idesc[nblk - 1], accumulate);
        };
#pragma unroll
        for (int e = 0; e < TY_HR; ++e) {
          const int yo_lo = e - 2 > 0 ? e - 2 : 0, yo_hi = e < TY_H - 1 ? e : TY_H - 1;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            if (a == 0 && e < TY_H) {
              mma(e, 0, e, 1, 2, 0u);                                       // ky = 0: first write of output row e
              if (e > yo_lo) mma(e, 0, yo_lo, e - yo_lo, 2 - (e - yo_lo), 1u);
            } else {
              mma(e, a, yo_lo, yo_hi - yo_lo + 1, 2 - (e - yo_lo), 1u);
            }
          }
        }
        umma_commit(&sync->h_empty[bar]);
        umma_commit(&sync->t_full[bar]);
      }
      __syncwarp();
    }
  } else {
    // ===================================================================== epilogue: thread = pixel pair of the row tile
    const int q = warp & 3;
    const int grp = (warp - TY_W_EPI) >> 2;
    const int m = q * 32 + lane;
    uint16_t* y16 = reinterpret_cast<uint16_t*>(p.y);
    const bool relu_all = p.relu_n >= 16;
    for (int i = grp, t = blockIdx.x + grp * gridDim.x; t < p.total_tiles;
         t += TY_EPI_GROUPS * gridDim.x, i += TY_EPI_GROUPS) {
      const int acc = i & (TY_ACC - 1), bar = i & (TY_NBAR - 1);
      const TyTile c = ty_decode(p, t);
      const int ox = c.ox0 + 2 * m;
      const bool xok = ox < p.W;                 // W is even: both pixels of a pair are inside or outside
      uint16_t* yrow = y16 + (((size_t)c.n * p.H + c.oy0) * p.W + ox) * 16;
      mbar_wait(&sync->t_full[bar], (uint32_t)(i / TY_NBAR) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16);
      uint32_t v[2][32];
      tmem_ld32(t_addr, v[0]);
#pragma unroll
      for (int yo = 0; yo < TY_H; ++yo) {
        tmem_ld_wait();
        if (yo < TY_H - 1) {
          tmem_ld32(t_addr + (uint32_t)(32 * (yo + 1)), v[(yo + 1) & 1]);
        } else {                               // accumulator read out: hand it back before the last row's math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sync->t_empty[bar]);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          uint32_t w[8];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 sc = *reinterpret_cast<const float4*>(&sync->scale[16 * s + 4 * e4]);
            const float4 sh = *reinterpret_cast<const float4*>(&sync->shift[16 * s + 4 * e4]);
            float a0 = fmaf(__uint_as_float(v[yo & 1][16 * s + 4 * e4]), sc.x, sh.x);
            float a1 = fmaf(__uint_as_float(v[yo & 1][16 * s + 4 * e4 + 1]), sc.y, sh.y);
            float a2 = fmaf(__uint_as_float(v[yo & 1][16 * s + 4 * e4 + 2]), sc.z, sh.z);
            float a3 = fmaf(__uint_as_float(v[yo & 1][16 * s + 4 * e4 + 3]), sc.w, sh.w);
            const int ch = 4 * e4;
            if (relu_all || ch < p.relu_n) a0 = fmaxf(a0, 0.f);
            if (relu_all || ch + 1 < p.relu_n) a1 = fmaxf(a1, 0.f);
            if (relu_all || ch + 2 < p.relu_n) a2 = fmaxf(a2, 0.f);
            if (relu_all || ch + 3 < p.relu_n) a3 = fmaxf(a3, 0.f);
            w[2 * e4] = pack2<DT>(a0, a1);
            w[2 * e4 + 1] = pack2<DT>(a2, a3);
          }
          if (xok && c.oy0 + yo < p.H) stg256(yrow + (size_t)yo * p.W * 16 + 16 * s, w);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TY_ACC * 256);
  }
}

struct TyMapCache {
  const void* ptr = nullptr;
  CUtensorMap map;
};

static PFN_cuTensorMapEncodeTiled_v12000 ty_encode_fn() {
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

bool conv_ty_supported(const drnb200_conv_desc& d) {
  static const char* env = getenv("DRNB200_TY");            // A/B knob: "0" keeps conv_halo for this layer
  if (env && env[0] == '0') return false;
  // W even: the pixel-pair view of a row must not straddle two rows (odd widths stay on conv_halo)
  return d.ksize == 3 && d.stride == 1 && d.dilation == 1 && d.Cin == 16 && d.tile_ci == 16 && d.Cout == 16 &&
         d.tile_o == 16 && !d.has_residual && !d.out_f32 && (d.x_cpitch == 0 || d.x_cpitch == 16) && d.W % 2 == 0 &&
         d.W >= 8;
}

int conv_ty_launch(drnb200_conv_plan* plan, cudaStream_t st) {
  const ConvParams& c = plan->p;
  const drnb200_conv_desc& d = plan->d;
  TyParams p{};
  p.x = c.x; p.y = c.y; p.w_packed = c.w_packed; p.kblk = c.kblk; p.scale = c.scale; p.shift = c.shift;
  p.n_kb = plan->h_row_ptr[1];
  if (p.n_kb == 0) return conv_direct_launch(plan, st);     // everything pruned: y = act(shift)
  p.N = c.N; p.H = c.H; p.W = c.W; p.relu_n = c.relu_n;
  p.tiles_x = (c.W + TY_W - 1) / TY_W;
  p.tiles_y = (c.H + TY_H - 1) / TY_H;
  p.total_tiles = c.N * p.tiles_x * p.tiles_y;
  if ((uint64_t)p.total_tiles * (uint64_t)std::max(p.tiles_x, p.tiles_y) >= (1ull << 32)) {
    set_error("conv_ty: problem too large for the 32-bit tile decode");
    return DRNB200_E_ARG;
  }
  p.magic_x = p.tiles_x == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_x - 1) / p.tiles_x);
  p.magic_y = p.tiles_y == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_y - 1) / p.tiles_y);
  for (int n = 0; n < 3; ++n) p.idesc[n] = umma_idesc_f16(128, 32 * (n + 1), d.act_dtype);
  const size_t smem = 1024 + (size_t)TY_RING * TY_SLOT + TY_WBYTES + sizeof(TySync);

  static_assert(sizeof(TyMapCache) <= sizeof(plan->gather_cache), "tensor-map cache storage too small");
  TyMapCache* cache = reinterpret_cast<TyMapCache*>(plan->gather_cache);
  if (!plan->gather_cache_init) { new (cache) TyMapCache(); plan->gather_cache_init = true; }
  if (cache->ptr != p.x) {
    auto fn = ty_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    cuuint64_t gdim[4] = {32, (cuuint64_t)(c.W / 2), (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {64, (cuuint64_t)c.W * 32, (cuuint64_t)c.H * c.W * 32};
    cuuint32_t box[4] = {32, (cuuint32_t)TY_HP, (cuuint32_t)TY_HR, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&cache->map, d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    4, const_cast<void*>(p.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(conv_ty) failed with CUresult %d (W=%d H=%d N=%d)", (int)r, c.W, c.H, c.N);
      return DRNB200_E_CUDA;
    }
    cache->ptr = p.x;
  }
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_tiles, sms);
  if (grid == 0) return DRNB200_OK;
  static std::atomic<unsigned long long> attr[2];
  const bool bf = d.act_dtype == DRNB200_BF16;
  if (attr_needed_on_this_device(attr[bf ? 1 : 0])) {
    if (bf) DRN_CUDA(cudaFuncSetAttribute(conv_ty_kernel<DRNB200_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else DRN_CUDA(cudaFuncSetAttribute(conv_ty_kernel<DRNB200_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (bf) launch_chained(conv_ty_kernel<DRNB200_BF16>, grid, TY_THREADS, smem, st, cache->map, p);
  else launch_chained(conv_ty_kernel<DRNB200_F16>, grid, TY_THREADS, smem, st, cache->map, p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

}  // namespace drnb200
