// conv_y2.cu — 3x3 STRIDE-2 convolution 32 -> 128 channels (DRN block 3.0: conv1 with the 1x1 downsample fused as 64
// extra output channels; 1/2 resolution in, 1/4 out): conv_ys.cu's row streaming over conv_s2.cu's pixel-pair rows.
//
// Why: the per-tap pipeline of conv_tc.cu loads this layer's input nine times with TMA element strides, one 64-byte
// row per pixel and tap: 2304 TMA rows per 256 output pixels, and the TMA row rate (~2.5 cycles per row) alone puts
// 0.09 ms under the launch (measured 0.116 ms; HBM floor 0.083 ms).  Here the input is the tensor {64, W/2, H, N} of
// pixel pairs (128-byte rows, SWIZZLE_128B); a work item is 128 output pixels of a row (UMMA M) x 4 output rows, its
// nine input rows 2*y0-1 .. 2*y0+7 pass once through a ring of single-row slots, and tap kx of output pixel m is a
// K-slice of pair m-1 (odd pixel) / pair m (even) / pair m (odd).  Even input rows feed one output row (ky = 1, N = 128;
// issued first: the accumulate-off MMA), odd rows two (ky = 2 of row j-1 and ky = 0 of row j: N = 256 against the stack
// [w(ky=2,kx); w(ky=0,kx)]): 54 MMAs per 512 output pixels, the four output rows of an item own the four 128-column
// slots of TMEM, output row j is committed as soon as input row 2j+1 has been issued.  Weights (72 KB) resident.
// Roles (320 threads): warp 0 row TMA producer, warp 1 TMEM allocation + MMA issue, warps 2-9 epilogue (two groups on
// alternate output rows; thread = output pixel: BN affine + ReLU on the first relu_n channels, the finished row — 128
// pixels x 256 B — staged as two SWIZZLE_128B halves and stored by two TMA stores).
// Barriers: one per row slot and per accumulator slot, each with ONE waiting warp (group) that meets its phases in order.
#include "conv_internal.cuh"
#include <algorithm>
#include <cudaTypedefs.h>
#include <new>

namespace drnb200 {

constexpr int Y2_W = 128, Y2_SEG = 4;          // work item: output pixels of a row (UMMA M) x output rows
constexpr int Y2_ROWS = 2 * Y2_SEG + 1;        // input rows per item (halo rows h = 0 .. 8 <-> input row 2*oy0 - 1 + h)
constexpr int Y2_HP = 136;                     // pixel pairs per row slot: 129 needed, rounded up so that a slot is a
                                               // multiple of the 1024-byte SWIZZLE_128B period
constexpr uint32_t Y2_PAIR = 128;              // bytes per pixel pair (2 x 32 channels x 16 bit)
constexpr uint32_t Y2_SLOT = Y2_HP * Y2_PAIR;  // 17408 = 17 x 1024
constexpr int Y2_RING = 5;
constexpr uint32_t Y2_WTAP = 128 * 64;         // one tap: 128 couts x 32 cin x 16 bit (64-byte rows, SWIZZLE_64B)
constexpr uint32_t Y2_WKX = 3 * Y2_WTAP;       // per kx: [ky=2; ky=0; ky=1]
constexpr uint32_t Y2_WBYTES = 3 * Y2_WKX;     // 72 KB
constexpr uint32_t Y2_HALF = Y2_W * 128;       // half of a finished output row: 128 pixels x 64 couts x 16 bit
constexpr uint32_t Y2_STAGE = 2 * Y2_HALF;     // a finished output row: 128 pixels x 256 B
constexpr int Y2_EPI_GROUPS = 2;
constexpr int Y2_W_EPI = 2;
constexpr int Y2_THREADS = (Y2_W_EPI + 4 * Y2_EPI_GROUPS) * 32;
static_assert(Y2_SEG * 128 == 512, "the output rows of an item own the whole TMEM: slot == output row");

struct Y2Params {
  const uint8_t* w_packed;     // live taps only, 8 KB each (pack_weights, tile 128 x 32, SWIZZLE_64B rows)
  const int32_t* kblk;         // tap index ky*3+kx of every packed tile
  const float* scale;
  const float* shift;
  int n_kb, N, OH, OW, relu_n;
  int tiles_x, tiles_y, total_tiles;
  uint32_t magic_x, magic_y;
  uint32_t idesc128, idesc256;
};

struct __align__(16) Y2Sync {
  uint64_t h_full[8], h_empty[8], t_full[Y2_SEG], t_empty[Y2_SEG], w_full;
  uint32_t tmem_base, pad;
  alignas(16) float scale[128];
  alignas(16) float shift[128];
};

struct Y2Tile { int n, ox0, oy0; };
__device__ __forceinline__ Y2Tile y2_decode(const Y2Params& p, int t) {
  Y2Tile c;
  const int q1 = p.tiles_x == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_x);
  const int txi = t - q1 * p.tiles_x;
  c.n = p.tiles_y == 1 ? q1 : (int)__umulhi((uint32_t)q1, p.magic_y);
  const int tyi = q1 - c.n * p.tiles_y;
  c.ox0 = txi * Y2_W; c.oy0 = tyi * Y2_SEG;
  return c;
}

// K-major operand descriptor without the start address: 8-row groups `sbo` bytes apart, layout 2 = SWIZZLE_128B,
// 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t y2_desc_hi(uint32_t sbo, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// the order in which the halo rows of an item are consumed: the even input row of an output row (its accumulate-off
// MMA) before the odd rows that touch it
__host__ __device__ constexpr int y2_row_at(int s) { return s == 2 * Y2_SEG ? 2 * Y2_SEG : ((s & 1) ? s - 1 : s + 1); }

template <int DT>
__global__ void __launch_bounds__(Y2_THREADS, 1)
conv_y2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y, const Y2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* rows = smem;                                    // Y2_RING x Y2_SLOT
  uint8_t* wsm = smem + (size_t)Y2_RING * Y2_SLOT;         // 3 x [384 rows][64 B]
  uint8_t* stage = wsm + Y2_WBYTES;                        // Y2_EPI_GROUPS x Y2_STAGE (1024-aligned): finished rows
  Y2Sync* sync = reinterpret_cast<Y2Sync*>(stage + Y2_EPI_GROUPS * Y2_STAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch();
  if (tid == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    for (int b = 0; b < Y2_RING; ++b) { mbar_init(&sync->h_full[b], 1); mbar_init(&sync->h_empty[b], 1); }
    for (int b = 0; b < Y2_SEG; ++b) { mbar_init(&sync->t_full[b], 1); mbar_init(&sync->t_empty[b], 4); }
    mbar_init(&sync->w_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sync->tmem_base, 512);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  if (tid == 0) {
    // resident weights: block (kx, j) <- the packed tile of tap ky*3+kx with j = 0, 1, 2 for ky = 2, 0, 1; one 8 KB
    // bulk copy per live tap (asynchronous: the MMA warp waits on w_full)
    mbar_arrive_expect_tx(&sync->w_full, (uint32_t)p.n_kb * Y2_WTAP);
    for (int k = 0; k < p.n_kb; ++k) {
      const int tap = __ldg(p.kblk + k), ky = tap / 3, kx = tap - ky * 3;
      const int j = ky == 2 ? 0 : (ky == 0 ? 1 : 2);
      bulk_load(p.w_packed + (size_t)k * Y2_WTAP, &sync->w_full, wsm + (size_t)(kx * 3 + j) * Y2_WTAP, Y2_WTAP);
    }
  }
  {  // pruned taps are zero blocks of the stacks
    uint32_t live = 0;
    for (int k = 0; k < p.n_kb; ++k) {
      const int tap = __ldg(p.kblk + k), ky = tap / 3, kx = tap - ky * 3;
      live |= 1u << (kx * 3 + (ky == 2 ? 0 : (ky == 0 ? 1 : 2)));
    }
    for (int blk = 0; blk < 9; ++blk)
      if (!((live >> blk) & 1u))
        for (int i = tid; i < (int)(Y2_WTAP / 16); i += Y2_THREADS)
          reinterpret_cast<uint4*>(wsm + (size_t)blk * Y2_WTAP)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < 128) { sync->scale[tid] = __ldg(p.scale + tid); sync->shift[tid] = __ldg(p.shift + tid); }
  fence_proxy_async_smem();       // written by the generic proxy, read by UMMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sync->tmem_base;

  if (warp == 0) {
    // ===================================================================== row TMA producer (warp-uniform loop)
    int b = 0;                      // ring slot of the next row and the parity of its use
    uint32_t bph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const Y2Tile c = y2_decode(p, t);
      for (int s = 0; s < Y2_ROWS; ++s) {
        const int h = y2_row_at(s);
        mbar_wait(&sync->h_empty[b], bph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&sync->h_full[b], Y2_SLOT);
          // tensor {64, W/2, H, N}; box {64, 136, 1, 1}; zero fill outside the image = the conv padding
          tma_load_4d(&tmap_x, &sync->h_full[b], rows + (size_t)b * Y2_SLOT, 0, c.ox0 - 1, 2 * c.oy0 - 1 + h, c.n);
        }
        __syncwarp();
        if (++b == Y2_RING) { b = 0; bph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform loop)
    const uint64_t a_hi = y2_desc_hi(1024u, 2);            // A: 128-byte rows, SWIZZLE_128B
    const uint64_t b_hi = y2_desc_hi(512u, 4);             // B: 64-byte rows, SWIZZLE_64B
    const uint32_t w16 = smem_u32(wsm) >> 4;
    int b = 0;
    uint32_t bph = 0, item = 0;
    mbar_wait(&sync->w_full, 0);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++item) {
#pragma unroll
      for (int s = 0; s < Y2_ROWS; ++s) {
        const int h = y2_row_at(s);
        mbar_wait(&sync->h_full[b], bph);
        // an even input row (odd h) starts output row (h - 1) / 2: its TMEM slot was read out in the previous item
        if (h & 1) mbar_wait(&sync->t_empty[(h - 1) >> 1], (item & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t h16 = smem_u32(rows + (size_t)b * Y2_SLOT) >> 4;
        if (elect_one()) {
          // tap column kx, K-step ks -> accumulator columns col .. col + n - 1 through the rows of the kx stack that
          // start at block j0 ([ky=2; ky=0; ky=1])
          auto mma = [&](int kx, int ks, int col, int n, int j0, uint32_t accumulate) {
            const uint32_t a_off = (kx == 0 ? 64u : (kx == 1 ? 128u : 192u)) + 32u * (uint32_t)ks;
            umma_f16(tmem_base + (uint32_t)col, a_hi | (uint64_t)(h16 + a_off / 16u),
                     b_hi | (uint64_t)(w16 + ((uint32_t)kx * Y2_WKX + (uint32_t)j0 * Y2_WTAP + 32u * (uint32_t)ks) / 16u),
                     n == 256 ? p.idesc256 : p.idesc128, accumulate);
          };
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              if (h & 1) {                                   // even input row 2j: output row j through ky = 1
                mma(kx, ks, 128 * ((h - 1) >> 1), 128, 2, (kx == 0 && ks == 0) ? 0u : 1u);
              } else if (h == 0) {                           // input row -1: output row 0 through ky = 0
                mma(kx, ks, 0, 128, 1, 1u);
              } else if (h == 2 * Y2_SEG) {                  // last odd input row: output row 3 through ky = 2
                mma(kx, ks, 128 * (Y2_SEG - 1), 128, 0, 1u);
              } else {                                       // odd input row 2j-1: rows j-1 (ky = 2) and j (ky = 0)
                mma(kx, ks, 128 * ((h >> 1) - 1), 256, 0, 1u);
              }
            }
          }
          umma_commit(&sync->h_empty[b]);
          if (!(h & 1) && h >= 2) umma_commit(&sync->t_full[(h >> 1) - 1]);   // output row h/2 - 1 has its three filter rows
        }
        __syncwarp();
        if (++b == Y2_RING) { b = 0; bph ^= 1u; }
      }
    }
  } else {
    // ===================================================================== epilogue: thread = output pixel of the row tile
    const int q = warp & 3;
    const int grp = (warp - Y2_W_EPI) >> 2;
    const int m = q * 32 + lane;
    uint8_t* stg = stage + (size_t)grp * Y2_STAGE;
    const uint32_t srow = (uint32_t)m * 128u, sx = (uint32_t)(m & 7);   // SWIZZLE_128B: chunk j of row m sits at j ^ (m & 7)
    const bool issuer = (q == 0 && lane == 0);
    const bool relu_all = p.relu_n >= 128;
    uint32_t item = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++item) {
      const Y2Tile c = y2_decode(p, t);
#pragma unroll 1
      for (int yo = grp; yo < Y2_SEG; yo += Y2_EPI_GROUPS) {
        mbar_wait(&sync->t_full[yo], item & 1u);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (uint32_t)(yo * 128) + ((uint32_t)(q * 32) << 16);
        // the TMA stores that last read this group's staging buffer have finished reading it
        if (issuer) bulk_wait_group_read<0>();
        named_bar_sync(1 + grp, 128);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                     // couts 0-63 / 64-127
          uint32_t v[2][32];
          tmem_ld32(t_addr + (uint32_t)(64 * hf), v[0]);
          tmem_ld32(t_addr + (uint32_t)(64 * hf + 32), v[1]);
          tmem_ld_wait();
          if (hf == 1) {                                     // accumulator slot read out: the next item may reuse it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sync->t_empty[yo]);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t w[8];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const int ch = 64 * hf + 16 * g + 4 * e4;
              const float4 sc = *reinterpret_cast<const float4*>(&sync->scale[ch]);
              const float4 sh = *reinterpret_cast<const float4*>(&sync->shift[ch]);
              const int vi = (16 * g + 4 * e4) & 31;
              float a0 = fmaf(__uint_as_float(v[g >> 1][vi]), sc.x, sh.x);
              float a1 = fmaf(__uint_as_float(v[g >> 1][vi + 1]), sc.y, sh.y);
              float a2 = fmaf(__uint_as_float(v[g >> 1][vi + 2]), sc.z, sh.z);
              float a3 = fmaf(__uint_as_float(v[g >> 1][vi + 3]), sc.w, sh.w);
              if (relu_all || ch < p.relu_n) a0 = fmaxf(a0, 0.f);
              if (relu_all || ch + 1 < p.relu_n) a1 = fmaxf(a1, 0.f);
              if (relu_all || ch + 2 < p.relu_n) a2 = fmaxf(a2, 0.f);
              if (relu_all || ch + 3 < p.relu_n) a3 = fmaxf(a3, 0.f);
              w[2 * e4] = pack2<DT>(a0, a1);
              w[2 * e4 + 1] = pack2<DT>(a2, a3);
            }
            uint8_t* dst = stg + (size_t)hf * Y2_HALF + srow;
            *reinterpret_cast<uint4*>(dst + ((((uint32_t)(2 * g)) ^ sx) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(dst + ((((uint32_t)(2 * g + 1)) ^ sx) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + grp, 128);
        if (issuer) {
          // tensor {128, OW, OH, N}; two boxes {64, 128, 1, 1}: pixels right of the image are clipped, rows below skipped
          if (c.oy0 + yo < p.OH) {
            tma_store_4d(&tmap_y, stg, 0, c.ox0, c.oy0 + yo, c.n);
            tma_store_4d(&tmap_y, stg + Y2_HALF, 64, c.ox0, c.oy0 + yo, c.n);
          }
          bulk_commit_group();
        }
      }
    }
    if (issuer) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct Y2MapCache {
  const void* ptr = nullptr;
  const void* ptr_y = nullptr;
  CUtensorMap map, map_y;
};

static PFN_cuTensorMapEncodeTiled_v12000 y2_encode_fn() {
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

bool conv_y2_supported(const drnb200_conv_desc& d) {
  static const char* env = getenv("DRNB200_Y2");            // A/B knob: "0" keeps the per-tap pipeline for this layer
  if (env && env[0] == '0') return false;
  // W even and the input tightly packed: a pixel pair must be 128 contiguous bytes inside one image row
  return d.ksize == 3 && d.stride == 2 && d.dilation == 1 && d.Cin == 32 && d.tile_ci == 32 && d.Cout == 128 &&
         d.tile_o == 128 && !d.has_residual && !d.out_f32 && d.proj_cin == 0 && (d.x_cpitch == 0 || d.x_cpitch == 32) &&
         d.W % 2 == 0 && d.W >= 16;
}

int conv_y2_launch(drnb200_conv_plan* plan, cudaStream_t st) {
  const ConvParams& c = plan->p;
  const drnb200_conv_desc& d = plan->d;
  Y2Params p{};
  p.w_packed = c.w_packed; p.kblk = c.kblk; p.scale = c.scale; p.shift = c.shift;
  p.n_kb = plan->h_row_ptr[1];
  if (p.n_kb == 0) return conv_direct_launch(plan, st);     // everything pruned: y = act(shift)
  p.N = c.N; p.OH = c.OH; p.OW = c.OW; p.relu_n = c.relu_n;
  p.tiles_x = (c.OW + Y2_W - 1) / Y2_W;
  p.tiles_y = (c.OH + Y2_SEG - 1) / Y2_SEG;
  p.total_tiles = c.N * p.tiles_x * p.tiles_y;
  if ((uint64_t)p.total_tiles * (uint64_t)std::max(p.tiles_x, p.tiles_y) >= (1ull << 32)) {
    set_error("conv_y2: problem too large for the 32-bit tile decode");
    return DRNB200_E_ARG;
  }
  p.magic_x = p.tiles_x == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_x - 1) / p.tiles_x);
  p.magic_y = p.tiles_y == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_y - 1) / p.tiles_y);
  p.idesc128 = umma_idesc_f16(128, 128, d.act_dtype);
  p.idesc256 = umma_idesc_f16(128, 256, d.act_dtype);
  const size_t smem = 1024 + (size_t)Y2_RING * Y2_SLOT + Y2_WBYTES + Y2_EPI_GROUPS * Y2_STAGE + sizeof(Y2Sync);

  static_assert(sizeof(Y2MapCache) <= sizeof(plan->gather_cache), "tensor-map cache storage too small");
  Y2MapCache* cache = reinterpret_cast<Y2MapCache*>(plan->gather_cache);
  if (!plan->gather_cache_init) { new (cache) Y2MapCache(); plan->gather_cache_init = true; }
  auto fn = y2_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
  const CUtensorMapDataType dt = d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (cache->ptr != c.x) {
    cuuint64_t gdim[4] = {64, (cuuint64_t)(c.W / 2), (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {128, (cuuint64_t)c.W * 64, (cuuint64_t)c.H * c.W * 64};
    cuuint32_t box[4] = {64, (cuuint32_t)Y2_HP, 1, 1};
    CUresult r = fn(&cache->map, dt, 4, const_cast<void*>(c.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(conv_y2) failed with CUresult %d (W=%d H=%d N=%d)", (int)r, c.W, c.H, c.N);
      return DRNB200_E_CUDA;
    }
    cache->ptr = c.x;
  }
  if (cache->ptr_y != c.y) {
    cuuint64_t gdim[4] = {128, (cuuint64_t)c.OW, (cuuint64_t)c.OH, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {256, (cuuint64_t)c.OW * 256, (cuuint64_t)c.OH * c.OW * 256};
    cuuint32_t box[4] = {64, (cuuint32_t)Y2_W, 1, 1};
    CUresult r = fn(&cache->map_y, dt, 4, c.y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(conv_y2 output) failed with CUresult %d (OW=%d OH=%d N=%d)", (int)r, c.OW, c.OH, c.N);
      return DRNB200_E_CUDA;
    }
    cache->ptr_y = c.y;
  }
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_tiles, sms);
  if (grid == 0) return DRNB200_OK;
  static std::atomic<unsigned long long> attr[2];
  const bool bf = d.act_dtype == DRNB200_BF16;
  if (attr_needed_on_this_device(attr[bf ? 1 : 0])) {
    if (bf) DRN_CUDA(cudaFuncSetAttribute(conv_y2_kernel<DRNB200_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else DRN_CUDA(cudaFuncSetAttribute(conv_y2_kernel<DRNB200_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (bf) launch_chained(conv_y2_kernel<DRNB200_BF16>, grid, Y2_THREADS, smem, st, cache->map, cache->map_y, p);
  else launch_chained(conv_y2_kernel<DRNB200_F16>, grid, Y2_THREADS, smem, st, cache->map, cache->map_y, p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

}  // namespace drnb200
