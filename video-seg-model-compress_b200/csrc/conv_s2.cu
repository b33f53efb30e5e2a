// conv_s2.cu — 3x3 STRIDE-2 convolution 16 -> 32 channels (DRN layer2: full resolution in, half resolution out)
// without any im2col copy: the stride is absorbed by the shared-memory operand layout.
//
// Why: conv_gather.cu expands every halo tile into an im2col tile inside shared memory; that smem -> smem pass runs
// the LDS/STS pipe at 75 % of its peak and bounds the layer at 0.196-0.20 ms per 8 frames against an HBM floor of
// 0.125 ms (profiles/r02_ncu_front_kernels.txt).  Here the input is viewed as PIXEL PAIRS — tensor {32, W/2, H, N},
// 64-byte rows, SWIZZLE_64B — so that the A operand "every other pixel" is an ordinary K-major operand whose rows
// are 64 bytes apart and whose 32-byte K-slice picks the even or the odd pixel of the pair:
//   output pixel m of a row tile, tap kx reads input pixel 2*(ox0 + m) + kx - 1
//     kx = 0 -> pair ox0 + m - 1, second half     kx = 1 -> pair ox0 + m, first half     kx = 2 -> pair ox0 + m, second half
// (a descriptor start of +32 bytes inside the 64-byte row; the swizzle XOR is a function of the absolute address).
// The y direction is folded into the WEIGHT operand as in conv_ty.cu: a tile is 128 output pixels of a row (UMMA M) x 4
// output rows, accumulator columns = (output row yo, cout).  Input row e = 2*yo + ky - 1:
//   even rows e = 2j   feed output row j through ky = 1                       (N = 32)
//   odd  rows e = 2j+1 feed output rows j (ky = 2) and j + 1 (ky = 0) at once (N = 64, B = [w(ky=2); w(ky=0)])
// The even row of an output row is issued BEFORE the odd rows that touch it, so the first MMA into every column block is
// the accumulate-off one and nothing has to be split: 27 MMAs (M = 128, K = 16) per 512 output pixels, each input row
// read from shared memory three times instead of being expanded and read nine times.
// Roles (352 threads): warp 0 halo TMA producer, warps 1-2 MMA issue on alternate tiles (warp 1 allocates TMEM),
// warps 3-10 epilogue (two groups on alternate tiles; thread = output pixel, BN affine + ReLU, 2 x 32-byte stores per
// row).  Barriers: tile i uses halo slot i % 2 and accumulator i % 4 but barrier i % 8 of each kind (conv_ty.cu explains
// why a barrier per resource is not enough once several warps poll it).
#include "conv_internal.cuh"
#include <algorithm>
#include <cudaTypedefs.h>
#include <new>

namespace drnb200 {

constexpr int S2_W = 128, S2_R = 4;            // output tile (pixels x rows)
constexpr int S2_HR = 2 * S2_R + 1;            // input rows: 2*oy0 - 1 .. 2*oy0 + 7
constexpr int S2_HP = 136;                     // pixel pairs per halo row: 129 needed (ox0 - 1 .. ox0 + 127), rounded up
                                               // so that a row is a multiple of the 512-byte SWIZZLE_64B period
constexpr uint32_t S2_PAIR = 64;               // bytes per pixel pair (2 x 16 channels x 16 bit)
constexpr uint32_t S2_ROWB = S2_HP * S2_PAIR;  // 8704 = 17 x 512
constexpr uint32_t S2_HALO_TX = S2_HR * S2_ROWB;
constexpr uint32_t S2_SLOT = (S2_HALO_TX + 1023u) & ~1023u;
constexpr uint32_t S2_WTAP = 32 * 32;          // one tap: 32 couts x 16 cin x 16 bit
constexpr uint32_t S2_WKX = 3 * S2_WTAP;       // per kx: [ky=2; ky=0; ky=1]
constexpr uint32_t S2_WBYTES = (3 * S2_WKX + 1023u) & ~1023u;
constexpr int S2_RING = 2;
constexpr int S2_MMA_WARPS = 2;
constexpr int S2_ACC = 4;                      // TMEM accumulators of 128 columns (4 rows x 32 couts)
constexpr int S2_NBAR = 8;
constexpr int S2_EPI_GROUPS = 2;
constexpr int S2_W_EPI = 1 + S2_MMA_WARPS;
constexpr int S2_THREADS = (S2_W_EPI + 4 * S2_EPI_GROUPS) * 32;
static_assert(S2_ACC % S2_MMA_WARPS == 0 && S2_RING % S2_MMA_WARPS == 0 && S2_NBAR % S2_ACC == 0 && S2_NBAR % S2_RING == 0,
              "slot / accumulator / barrier indices are masks of the tile index");

struct S2Params {
  const void* x;
  void* y;
  const uint8_t* w_packed;     // live taps only, 1 KB each (pack_weights, tile 32 x 16, SWIZZLE_32B rows)
  const int32_t* kblk;         // tap index ky*3+kx of every packed tile
  const float* scale;
  const float* shift;
  int n_kb, N, OH, OW, relu_n;
  int tiles_x, tiles_y, total_tiles;
  uint32_t magic_x, magic_y;
  uint32_t idesc32, idesc64;   // M = 128, N = 32 / 64
};

struct __align__(16) S2Sync {
  uint64_t h_full[S2_NBAR], h_empty[S2_NBAR], t_full[S2_NBAR], t_empty[S2_NBAR];
  uint32_t tmem_base, pad[3];
  alignas(16) float scale[32];
  alignas(16) float shift[32];
};

struct S2Tile { int n, ox0, oy0; };
__device__ __forceinline__ S2Tile s2_decode(const S2Params& p, int t) {
  S2Tile c;
  const int q1 = p.tiles_x == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_x);
  const int txi = t - q1 * p.tiles_x;
  c.n = p.tiles_y == 1 ? q1 : (int)__umulhi((uint32_t)q1, p.magic_y);
  const int tyi = q1 - c.n * p.tiles_y;
  c.ox0 = txi * S2_W; c.oy0 = tyi * S2_R;
  return c;
}

// K-major operand descriptor without the start address: 8-row groups `sbo` bytes apart, layout 4 = SWIZZLE_64B,
// 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t s2_desc_hi(uint32_t sbo, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

template <int DT>
__global__ void __launch_bounds__(S2_THREADS, 1)
conv_s2_kernel(const __grid_constant__ CUtensorMap tmap_x, const S2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* halo = smem;                                    // S2_RING x S2_SLOT
  uint8_t* wsm = smem + (size_t)S2_RING * S2_SLOT;         // 3 x [96 rows][32 B]
  S2Sync* sync = reinterpret_cast<S2Sync*>(wsm + S2_WBYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch();
  if (tid == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int b = 0; b < S2_NBAR; ++b) {
      mbar_init(&sync->h_full[b], 1); mbar_init(&sync->h_empty[b], 1);
      mbar_init(&sync->t_full[b], 1); mbar_init(&sync->t_empty[b], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sync->tmem_base, S2_ACC * 128);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  // resident weights: per kx the stack [ky=2; ky=0; ky=1], zeros where the tap is pruned
  for (int it = tid; it < 9 * 64; it += S2_THREADS) {
    const int slot = it >> 6, part = it & 63;
    const int kx = slot / 3, j = slot - kx * 3;
    const int ky = j == 0 ? 2 : (j == 1 ? 0 : 1);
    int idx = -1;
    for (int k = 0; k < p.n_kb; ++k)
      if (__ldg(p.kblk + k) == ky * 3 + kx) idx = k;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (idx >= 0) v = __ldg(reinterpret_cast<const uint4*>(p.w_packed + (size_t)idx * S2_WTAP) + part);
    reinterpret_cast<uint4*>(wsm + (size_t)slot * S2_WTAP)[part] = v;
  }
  if (tid < 32) { sync->scale[tid] = __ldg(p.scale + tid); sync->shift[tid] = __ldg(p.shift + tid); }
  fence_proxy_async_smem();       // written by the generic proxy, read by UMMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sync->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform loop)
    int i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const S2Tile c = s2_decode(p, t);
      // slot i % 2 was read by the MMAs of tile i - 2, which commit to barrier (i - 2) % 8
      if (i >= S2_RING) mbar_wait(&sync->h_empty[(i - S2_RING) & (S2_NBAR - 1)], (uint32_t)((i - S2_RING) / S2_NBAR) & 1u);
      if (elect_one()) {
        uint64_t* full = &sync->h_full[i & (S2_NBAR - 1)];
        mbar_arrive_expect_tx(full, S2_HALO_TX);
        // tensor {32, W/2, H, N}; box {32, 136, 9, 1}; zero fill outside the image = the conv padding
        tma_load_4d(&tmap_x, full, halo + (size_t)(i & (S2_RING - 1)) * S2_SLOT, 0, c.ox0 - 1, 2 * c.oy0 - 1, c.n);
      }
      __syncwarp();
    }
  } else if (warp < S2_W_EPI) {
    // ===================================================================== MMA issuers (warp-uniform loop)
    const uint64_t a_hi = s2_desc_hi(8u * S2_PAIR, 4);     // A: 64-byte rows, SWIZZLE_64B
    const uint64_t b_hi = s2_desc_hi(8u * 32u, 6);         // B: 32-byte rows, SWIZZLE_32B
    const uint32_t w16 = smem_u32(wsm) >> 4;
    const int mw = warp - 1;
    int i = mw;
    for (int t = blockIdx.x + mw * gridDim.x; t < p.total_tiles; t += S2_MMA_WARPS * gridDim.x, i += S2_MMA_WARPS) {
      const int acc = i & (S2_ACC - 1), b = i & (S2_RING - 1), bar = i & (S2_NBAR - 1);
      mbar_wait(&sync->h_full[bar], (uint32_t)(i / S2_NBAR) & 1u);
      // the accumulator was read out by the epilogue of tile i - 4, which arrives on barrier (i - 4) % 8
      if (i >= S2_ACC) mbar_wait(&sync->t_empty[(i - S2_ACC) & (S2_NBAR - 1)], (uint32_t)((i - S2_ACC) / S2_NBAR) & 1u);
      tc_fence_after();
      const uint32_t h16 = smem_u32(halo + (size_t)b * S2_SLOT) >> 4;
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 128u;
      if (elect_one()) {
        // halo row h (input row 2*oy0 - 1 + h), tap kx -> accumulator columns col .. col + n - 1 through the rows of
        // the per-kx weight stack that start at slot j0 ([ky=2; ky=0; ky=1])
        auto mma = [&](int h, int kx, int col, int n, int j0, uint32_t accumulate) {
          const uint32_t a_off = (uint32_t)h * S2_ROWB + (kx == 0 ? 32u : (kx == 1 ? 64u : 96u));
          umma_f16(d_tmem + (uint32_t)col, a_hi | (uint64_t)(h16 + a_off / 16u),
                   b_hi | (uint64_t)(w16 + ((uint32_t)kx * S2_WKX + (uint32_t)j0 * S2_WTAP) / 16u),
                   n == 64 ? p.idesc64 : p.idesc32, accumulate);
        };
#pragma unroll
        for (int j = 0; j < S2_R; ++j) {
          // even input row 2j (halo row 2j + 1): output row j through ky = 1; the first MMA into these columns
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) mma(2 * j + 1, kx, 32 * j, 32, 2, kx == 0 ? 0u : 1u);
          // odd input row 2j - 1 (halo row 2j): output rows j - 1 (ky = 2) and j (ky = 0); row -1 has no j - 1
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (j == 0) mma(0, kx, 0, 32, 1, 1u);
            else mma(2 * j, kx, 32 * (j - 1), 64, 0, 1u);
          }
        }
        // last odd input row (halo row 8): output row 3 through ky = 2
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) mma(2 * S2_R, kx, 32 * (S2_R - 1), 32, 0, 1u);
        umma_commit(&sync->h_empty[bar]);
        umma_commit(&sync->t_full[bar]);
      }
      __syncwarp();
    }
  } else {
    // ===================================================================== epilogue: thread = output pixel of the row tile
    const int q = warp & 3;
    const int grp = (warp - S2_W_EPI) >> 2;
    const int m = q * 32 + lane;
    uint16_t* y16 = reinterpret_cast<uint16_t*>(p.y);
    const bool relu_all = p.relu_n >= 32;
    for (int i = grp, t = blockIdx.x + grp * gridDim.x; t < p.total_tiles;
         t += S2_EPI_GROUPS * gridDim.x, i += S2_EPI_GROUPS) {
      const int acc = i & (S2_ACC - 1), bar = i & (S2_NBAR - 1);
      const S2Tile c = s2_decode(p, t);
      const int ox = c.ox0 + m;
      const bool xok = ox < p.OW;
      uint16_t* yrow = y16 + (((size_t)c.n * p.OH + c.oy0) * p.OW + ox) * 32;
      mbar_wait(&sync->t_full[bar], (uint32_t)(i / S2_NBAR) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * 128u + ((uint32_t)(q * 32) << 16);
      uint32_t v[2][32];
      tmem_ld32(t_addr, v[0]);
#pragma unroll
      for (int yo = 0; yo < S2_R; ++yo) {
        tmem_ld_wait();
        if (yo < S2_R - 1) {
          tmem_ld32(t_addr + (uint32_t)(32 * (yo + 1)), v[(yo + 1) & 1]);
        } else {                               // accumulator read out: hand it back before the last row's math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sync->t_empty[bar]);
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t w[8];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 sc = *reinterpret_cast<const float4*>(&sync->scale[16 * hf + 4 * e4]);
            const float4 sh = *reinterpret_cast<const float4*>(&sync->shift[16 * hf + 4 * e4]);
            float a0 = fmaf(__uint_as_float(v[yo & 1][16 * hf + 4 * e4]), sc.x, sh.x);
            float a1 = fmaf(__uint_as_float(v[yo & 1][16 * hf + 4 * e4 + 1]), sc.y, sh.y);
            float a2 = fmaf(__uint_as_float(v[yo & 1][16 * hf + 4 * e4 + 2]), sc.z, sh.z);
            float a3 = fmaf(__uint_as_float(v[yo & 1][16 * hf + 4 * e4 + 3]), sc.w, sh.w);
            const int ch = 16 * hf + 4 * e4;
            if (relu_all || ch < p.relu_n) a0 = fmaxf(a0, 0.f);
            if (relu_all || ch + 1 < p.relu_n) a1 = fmaxf(a1, 0.f);
            if (relu_all || ch + 2 < p.relu_n) a2 = fmaxf(a2, 0.f);
            if (relu_all || ch + 3 < p.relu_n) a3 = fmaxf(a3, 0.f);
            w[2 * e4] = pack2<DT>(a0, a1);
            w[2 * e4 + 1] = pack2<DT>(a2, a3);
          }
          if (xok && c.oy0 + yo < p.OH) stg256(yrow + (size_t)yo * p.OW * 32 + 16 * hf, w);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S2_ACC * 128);
  }
}

struct S2MapCache {
  const void* ptr = nullptr;
  CUtensorMap map;
};

static PFN_cuTensorMapEncodeTiled_v12000 s2_encode_fn() {
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

bool conv_s2_supported(const drnb200_conv_desc& d) {
  static const char* env = getenv("DRNB200_S2");            // A/B knob: "0" keeps conv_gather for this layer
  if (env && env[0] == '0') return false;
  // W even: the pixel-pair view of a row must not straddle two rows
  return d.ksize == 3 && d.stride == 2 && d.dilation == 1 && d.Cin == 16 && d.tile_ci == 16 && d.Cout == 32 &&
         d.tile_o == 32 && !d.has_residual && !d.out_f32 && (d.x_cpitch == 0 || d.x_cpitch == 16) && d.W % 2 == 0 &&
         d.W >= 16;
}

int conv_s2_launch(drnb200_conv_plan* plan, cudaStream_t st) {
  const ConvParams& c = plan->p;
  const drnb200_conv_desc& d = plan->d;
  S2Params p{};
  p.x = c.x; p.y = c.y; p.w_packed = c.w_packed; p.kblk = c.kblk; p.scale = c.scale; p.shift = c.shift;
  p.n_kb = plan->h_row_ptr[1];
  if (p.n_kb == 0) return conv_direct_launch(plan, st);     // everything pruned: y = act(shift)
  p.N = c.N; p.OH = c.OH; p.OW = c.OW; p.relu_n = c.relu_n;
  p.tiles_x = (c.OW + S2_W - 1) / S2_W;
  p.tiles_y = (c.OH + S2_R - 1) / S2_R;
  p.total_tiles = c.N * p.tiles_x * p.tiles_y;
  if ((uint64_t)p.total_tiles * (uint64_t)std::max(p.tiles_x, p.tiles_y) >= (1ull << 32)) {
    set_error("conv_s2: problem too large for the 32-bit tile decode");
    return DRNB200_E_ARG;
  }
  p.magic_x = p.tiles_x == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_x - 1) / p.tiles_x);
  p.magic_y = p.tiles_y == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_y - 1) / p.tiles_y);
  p.idesc32 = umma_idesc_f16(128, 32, d.act_dtype);
  p.idesc64 = umma_idesc_f16(128, 64, d.act_dtype);
  const size_t smem = 1024 + (size_t)S2_RING * S2_SLOT + S2_WBYTES + sizeof(S2Sync);

  static_assert(sizeof(S2MapCache) <= sizeof(plan->gather_cache), "tensor-map cache storage too small");
  S2MapCache* cache = reinterpret_cast<S2MapCache*>(plan->gather_cache);
  if (!plan->gather_cache_init) { new (cache) S2MapCache(); plan->gather_cache_init = true; }
  if (cache->ptr != p.x) {
    auto fn = s2_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    cuuint64_t gdim[4] = {32, (cuuint64_t)(c.W / 2), (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {64, (cuuint64_t)c.W * 32, (cuuint64_t)c.H * c.W * 32};
    cuuint32_t box[4] = {32, (cuuint32_t)S2_HP, (cuuint32_t)S2_HR, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&cache->map, d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    4, const_cast<void*>(p.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(conv_s2) failed with CUresult %d (W=%d H=%d N=%d)", (int)r, c.W, c.H, c.N);
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
    if (bf) DRN_CUDA(cudaFuncSetAttribute(conv_s2_kernel<DRNB200_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else DRN_CUDA(cudaFuncSetAttribute(conv_s2_kernel<DRNB200_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (bf) launch_chained(conv_s2_kernel<DRNB200_BF16>, grid, S2_THREADS, smem, st, cache->map, p);
  else launch_chained(conv_s2_kernel<DRNB200_F16>, grid, S2_THREADS, smem, st, cache->map, p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

}  // namespace drnb200
