// conv_ty.cu — 3x3 stride-1 convolution 16 -> 16 channels (DRN layer1, full resolution) with the y direction of
// the filter folded into the WEIGHT operand ("Toeplitz along y").
//
// Why: conv_halo.cu runs this layer as 9 shifted-window MMAs per 128 pixels, i.e. every pixel's 32 bytes are read
// from shared memory nine times as the A operand (ncu, profiles/r02_ncu_front_kernels.txt: 42.5 M operand wavefronts
// = 5.4 GB per 8 frames, the tensor-core read port at 73 % of its peak, tensor pipe 16 % busy, 0.225 ms against an
// HBM floor of 0.166 ms).  Here a tile is 128 pixels of a row (UMMA M) x 8 output rows and the accumulator columns
// are (output row yo, cout): input row e of the halo (rows y0-1 .. y0+8), shifted by kx pixels, is ONE A operand
// that feeds the up to three output rows yo = e - ky it contributes to — the B operand is the stack
// [w(ky=2,kx); w(ky=1,kx); w(ky=0,kx)] (48 x 16) or a 16/32-row window of it at the top and bottom of the tile.
// 37 MMAs (M=128, N=16/32/48, K=16) per 1024 pixels instead of 72 of N=16: 152 KB instead of 324 KB of operand reads.
// An output row's column block is first written by its own ky=0 MMA (accumulate off), which is why that MMA is issued
// separately from the ky=1,2 window of the same input row.
// Roles (384 threads): warp 0 halo TMA producer, warps 1-3 MMA issue on alternate tiles (warp 1 allocates TMEM; two of
// the three are used: measured 2 % faster than three, DRNB200_TY_NMMA), warps 4-11 epilogue (two groups on alternate
// tiles; thread = pixel, BN affine + ReLU, 32-byte stores).
// Barriers: tile i of a CTA uses halo slot i % 4 and accumulator i % 4, but barrier i % 8 of each kind.  With one
// barrier per slot, MMA warp A (tiles 1, 4, ...) can poll "slot 0, second fill" while the FIRST fill of slot 0 is
// still in flight (warp B has not consumed tile 0 yet and TMA boxes complete out of order once the input comes from
// HBM): mbarrier.try_wait.parity on a barrier that is two phases behind answers "done".  Measured: 1-2 % of the
// launches of two 1024x2048 frames faulted that way.  Twice as many barriers as resources keeps every waiter within
// one phase of its barrier whatever the interleaving of the warps (the look-ahead is bounded by the four slots).
#include "conv_internal.cuh"
#include <algorithm>
#include <cudaTypedefs.h>
#include <new>

namespace drnb200 {

constexpr int TY_W = 128, TY_H = 8;            // output tile (pixels x rows)
constexpr int TY_HR = TY_H + 2;                // halo rows
constexpr int TY_HP = 136;                     // halo pixels per row: 130 needed, rounded up so that a row is a
                                               // multiple of the 256-byte SWIZZLE_32B period
constexpr uint32_t TY_PITCH = 32;              // bytes per pixel (16 channels x 16 bit)
constexpr uint32_t TY_ROWB = TY_HP * TY_PITCH; // 4352
constexpr uint32_t TY_HALO_TX = TY_HR * TY_ROWB;
constexpr uint32_t TY_SLOT = (TY_HALO_TX + 1023u) & ~1023u;
constexpr uint32_t TY_WTAP = 16 * TY_PITCH;    // one tap: 16 couts x 16 cin
constexpr uint32_t TY_WKX = 3 * TY_WTAP;       // per kx: [ky=2; ky=1; ky=0]
constexpr uint32_t TY_WBYTES = (3 * TY_WKX + 1023u) & ~1023u;
constexpr int TY_RING = 4;                     // halo slots (4 x 43 KB)
constexpr int TY_MMA_WARPS = 3;
constexpr int TY_ACC = 4;                      // TMEM accumulators of 128 columns
constexpr int TY_NBAR = 2 * TY_RING;           // barriers per kind (see the header)
static_assert(TY_ACC == TY_RING && (TY_RING & (TY_RING - 1)) == 0, "slot / accumulator / barrier indices are masks of i");
constexpr int TY_EPI_GROUPS = 2;
constexpr int TY_W_EPI = 1 + TY_MMA_WARPS;
constexpr int TY_THREADS = (TY_W_EPI + 4 * TY_EPI_GROUPS) * 32;

struct TyParams {
  const void* x;
  void* y;
  const uint8_t* w_packed;     // live taps only, 512 bytes each (pack_weights, tile 16 x 16, SWIZZLE_32B rows)
  const int32_t* kblk;         // tap index ky*3+kx of every packed tile
  const float* scale;
  const float* shift;
  int n_kb, N, H, W, relu_n;
  int tiles_x, tiles_y, total_tiles, n_mma;
  uint32_t magic_x, magic_y;
  uint32_t idesc[3];           // N = 16, 32, 48
};

struct __align__(16) TySync {
  uint64_t h_full[TY_NBAR], h_empty[TY_NBAR], t_full[TY_NBAR], t_empty[TY_NBAR];
  uint32_t tmem_base;
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

// K-major SWIZZLE_32B operand descriptor without the start address: 8-row groups `sbo` bytes apart
__device__ __forceinline__ uint64_t ty_desc_hi(uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;        // SWIZZLE_32B
  return d;
}

template <int DT>
__global__ void __launch_bounds__(TY_THREADS, 1)
conv_ty_kernel(const __grid_constant__ CUtensorMap tmap_x, const TyParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* halo = smem;                                    // ring x TY_SLOT
  uint8_t* wsm = smem + (size_t)TY_RING * TY_SLOT;         // 3 x [48 rows][32 B]
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
    tmem_alloc(&sync->tmem_base, TY_ACC * 128);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  // resident weights: slot (kx, 2 - ky) <- the packed tile of tap ky*3+kx, zeros when the tap is pruned
  if (tid < 9 * 32) {
    const int slot = tid >> 5, part = tid & 31;
    const int kx = slot / 3, ky = 2 - (slot - kx * 3);
    int idx = -1;
    for (int k = 0; k < p.n_kb; ++k)
      if (__ldg(p.kblk + k) == ky * 3 + kx) idx = k;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (idx >= 0) v = __ldg(reinterpret_cast<const uint4*>(p.w_packed + (size_t)idx * TY_WTAP) + part);
    reinterpret_cast<uint4*>(wsm + (size_t)slot * TY_WTAP)[part] = v;
  }
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
      // slot i % 4 was read by the MMAs of tile i - 4, which commit to barrier (i - 4) % 8
      if (i >= TY_RING) mbar_wait(&sync->h_empty[(i - TY_RING) & (TY_NBAR - 1)], (uint32_t)((i - TY_RING) / TY_NBAR) & 1u);
      if (elect_one()) {
        uint64_t* full = &sync->h_full[i & (TY_NBAR - 1)];
        mbar_arrive_expect_tx(full, TY_HALO_TX);
        // tensor {16, W, H, N}; box {16, 136, 10, 1}; zero fill outside the image = the conv padding
        tma_load_4d(&tmap_x, full, halo + (size_t)(i & (TY_RING - 1)) * TY_SLOT, 0, c.ox0 - 1, c.oy0 - 1, c.n);
      }
      __syncwarp();
    }
  } else if (warp < TY_W_EPI) {
    // ===================================================================== MMA issuers (warp-uniform loop)
    const uint64_t d_hi = ty_desc_hi(8u * TY_PITCH);       // A and B: 8-row groups are contiguous (256 B)
    const uint32_t w16 = smem_u32(wsm) >> 4;
    const int mw = warp - 1;
    int i = mw;
    for (int t = blockIdx.x + mw * gridDim.x; mw < p.n_mma && t < p.total_tiles;
         t += p.n_mma * gridDim.x, i += p.n_mma) {
      const int acc = i & (TY_ACC - 1), b = i & (TY_RING - 1), bar = i & (TY_NBAR - 1);
      mbar_wait(&sync->h_full[bar], (uint32_t)(i / TY_NBAR) & 1u);
      // the accumulator was read out by the epilogue of tile i - 4, which arrives on barrier (i - 4) % 8
      if (i >= TY_ACC) mbar_wait(&sync->t_empty[(i - TY_ACC) & (TY_NBAR - 1)], (uint32_t)((i - TY_ACC) / TY_NBAR) & 1u);
      tc_fence_after();
      const uint32_t h16 = smem_u32(halo + (size_t)b * TY_SLOT) >> 4;
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 128u;
      if (elect_one()) {
        // input row e, shifted kx pixels -> output rows yo_lo .. yo_lo + nblk - 1 through B rows brow .. (stack of
        // ky = 2, 1, 0 per kx: output row yo takes ky = e - yo)
        auto mma = [&](int e, int kx, int yo_lo, int nblk, int brow, uint32_t accumulate) {
          umma_f16(d_tmem + (uint32_t)(yo_lo * 16),
                   d_hi | (uint64_t)(h16 + ((uint32_t)e * TY_ROWB + (uint32_t)kx * TY_PITCH) / 16u),
                   d_hi | (uint64_t)(w16 + ((uint32_t)kx * TY_WKX + (uint32_t)brow * TY_PITCH) / 16u),
                   p.idesc[nblk - 1], accumulate);
        };
#pragma unroll
        for (int e = 0; e < TY_HR; ++e) {
          const int yo_lo = e - 2 > 0 ? e - 2 : 0, yo_hi = e < TY_H - 1 ? e : TY_H - 1;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (kx == 0 && e < TY_H) {
              mma(e, 0, e, 1, 32, 0u);                                      // ky = 0: first write of output row e
              if (e > yo_lo) mma(e, 0, yo_lo, e - yo_lo, (2 - (e - yo_lo)) * 16, 1u);
            } else {
              mma(e, kx, yo_lo, yo_hi - yo_lo + 1, (2 - (e - yo_lo)) * 16, 1u);
            }
          }
        }
        umma_commit(&sync->h_empty[bar]);
        umma_commit(&sync->t_full[bar]);
      }
      __syncwarp();
    }
  } else {
    // ===================================================================== epilogue: thread = pixel of the row tile
    const int q = warp & 3;
    const int grp = (warp - TY_W_EPI) >> 2;
    const int m = q * 32 + lane;
    uint16_t* y16 = reinterpret_cast<uint16_t*>(p.y);
    float sc[16], sh[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) { sc[e] = __ldg(p.scale + e); sh[e] = __ldg(p.shift + e); }
    const bool relu_all = p.relu_n >= 16;
    for (int i = grp, t = blockIdx.x + grp * gridDim.x; t < p.total_tiles;
         t += TY_EPI_GROUPS * gridDim.x, i += TY_EPI_GROUPS) {
      const int acc = i & (TY_ACC - 1), bar = i & (TY_NBAR - 1);
      const TyTile c = ty_decode(p, t);
      const int ox = c.ox0 + m;
      const bool xok = ox < p.W;
      uint16_t* yrow = y16 + (((size_t)c.n * p.H + c.oy0) * p.W + ox) * 16;
      mbar_wait(&sync->t_full[bar], (uint32_t)(i / TY_NBAR) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * 128u + ((uint32_t)(q * 32) << 16);
      uint32_t v[2][32];
      tmem_ld32(t_addr, v[0]);
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        tmem_ld_wait();
        if (blk < 3) {
          tmem_ld32(t_addr + (uint32_t)(32 * (blk + 1)), v[(blk + 1) & 1]);
        } else {                               // accumulator read out: hand it back before the last two rows' math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sync->t_empty[bar]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int yo = 2 * blk + u;
          float f[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            f[e] = fmaf(__uint_as_float(v[blk & 1][16 * u + e]), sc[e], sh[e]);
            if (relu_all || e < p.relu_n) f[e] = fmaxf(f[e], 0.f);
          }
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = pack2<DT>(f[2 * e], f[2 * e + 1]);
          if (xok && c.oy0 + yo < p.H) stg256(yrow + (size_t)yo * p.W * 16, w);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TY_ACC * 128);
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
  return d.ksize == 3 && d.stride == 1 && d.dilation == 1 && d.Cin == 16 && d.tile_ci == 16 && d.Cout == 16 &&
         d.tile_o == 16 && !d.has_residual && !d.out_f32 && (d.x_cpitch == 0 || d.x_cpitch == 16) && d.W >= 8;
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
  for (int n = 0; n < 3; ++n) p.idesc[n] = umma_idesc_f16(128, 16 * (n + 1), d.act_dtype);
  static const char* env_nmma = getenv("DRNB200_TY_NMMA");  // A/B knob: MMA-issuing warps in use (1..3), same results
  p.n_mma = env_nmma ? std::max(1, std::min(TY_MMA_WARPS, atoi(env_nmma))) : 2;
  const size_t smem = 1024 + (size_t)TY_RING * TY_SLOT + TY_WBYTES + sizeof(TySync);

  static_assert(sizeof(TyMapCache) <= sizeof(plan->gather_cache), "tensor-map cache storage too small");
  TyMapCache* cache = reinterpret_cast<TyMapCache*>(plan->gather_cache);
  if (!plan->gather_cache_init) { new (cache) TyMapCache(); plan->gather_cache_init = true; }
  if (cache->ptr != p.x) {
    auto fn = ty_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    cuuint64_t gdim[4] = {16, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {32, (cuuint64_t)c.W * 32, (cuuint64_t)c.H * c.W * 32};
    cuuint32_t box[4] = {16, (cuuint32_t)TY_HP, (cuuint32_t)TY_HR, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&cache->map, d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    4, const_cast<void*>(p.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
