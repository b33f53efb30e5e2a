// conv_ys.cu — 3x3 stride-1 convolution 64 -> 64 channels (+ residual) (DRN layer3 blocks, 1/4 resolution) that
// STREAMS input rows: every input row is loaded once per 8-row segment and multiplied against the filter rows stacked
// in the weight operand, so that one MMA per (tap column, K-step) serves the three output rows the input row feeds.
//
// Why: conv_halo.cu runs these layers as 9 taps x 4 K-steps of N = 64 per 128 pixels = 36 MMAs that read 6 KB of
// operands each (216 KB per 128 pixels; ncu, profiles/r02_ncu_front_kernels.txt: the tensor-core read port of shared
// memory 60-66 % busy, tensor pipe 40-44 %, 0.083-0.094 ms per launch against HBM floors of 0.042-0.062 ms), from 55 KB
// halo tiles of which only two fit beside the 72 KB of weights.  Here a work item is 128 pixels of a row (UMMA M) x 8
// output rows; its ten input rows y0-1 .. y0+8 pass through a ring of single-row slots (136 pixels x 128 B,
// SWIZZLE_128B), and input row e, shifted by kx pixels, K-step ks, is ONE A operand against the stack
// [w(ky=2,kx); w(ky=1,kx); w(ky=0,kx)] (192 couts x 64 cin per kx) or the 64/128-row window of it that stays inside
// the segment: the accumulator columns are (output row yo, cout), the eight output rows of an item own the eight
// 64-column slots of TMEM, output row yo = e - ky.  127 MMAs (N = 64/128/192) per 1024 pixels instead of 288 of N = 64,
// 144 KB instead of 216 KB of operand reads per 128 pixels, and output row e-2 is complete — committed to its own
// barrier and picked up by an epilogue group — as soon as input row e has been issued, six rows before the item ends.
// An output row's columns are first written by its own (ky = 0, kx = 0, ks = 0) MMA (accumulate off), which is why that
// one is issued separately from the ky = 1, 2 window of the same input row.
// Roles (320 threads): warp 0 row TMA producer, warp 1 TMEM allocation + MMA issue, warps 2-9 epilogue (two groups on
// alternate output rows; thread = pixel: BN affine + residual + ReLU).  Both sides of the epilogue go through shared
// memory and TMA: the residual row (128 pixels x 128 B) of the group's NEXT output row is fetched by a TMA load the
// group issues as soon as it has read the current one into registers, the finished row is staged, SWIZZLE_128B, and
// leaves as one TMA store.  With thread = pixel, direct 32-byte global accesses touch 32 cache lines per warp
// instruction and the L1 tag stage, not HBM, sets the pace (measured: 0.074 -> 0.060 ms without residual once the
// stores were staged; with residual 0.105 -> 0.094 ms while its loads were still per-thread).
// Barriers: one per row slot and per accumulator slot; each has ONE waiting warp (group) that meets its phases in
// order (slot yo always belongs to epilogue group yo & 1), which is what makes single barriers safe (see conv_ty.cu).
#include "conv_internal.cuh"
#include <algorithm>
#include <cudaTypedefs.h>
#include <new>

namespace drnb200 {

constexpr int YS_W = 128, YS_SEG = 8;          // work item: pixels of a row (UMMA M) x output rows
constexpr int YS_ROWS = YS_SEG + 2;            // input rows per item
constexpr int YS_HP = 136;                     // pixels per row slot: 130 needed, rounded up so that a slot is a
                                               // multiple of the 1024-byte SWIZZLE_128B period
constexpr uint32_t YS_PIX = 128;               // bytes per pixel (64 channels x 16 bit)
constexpr uint32_t YS_SLOT = YS_HP * YS_PIX;   // 17408 = 17 x 1024
__host__ __device__ constexpr int ys_ring(bool has_res) { return has_res ? 5 : 6; }   // the residual buffers cost a row slot
constexpr uint32_t YS_STAGE = YS_W * YS_PIX;   // one finished output row of an item: 128 pixels x 128 B
constexpr uint32_t YS_WTAP = 64 * 128;         // one tap: 64 couts x 64 cin x 16 bit
constexpr uint32_t YS_WKX = 3 * YS_WTAP;       // per kx: [ky=2; ky=1; ky=0]
constexpr uint32_t YS_WBYTES = 3 * YS_WKX;     // 72 KB
constexpr int YS_EPI_GROUPS = 2;
constexpr int YS_W_EPI = 2;
constexpr int YS_THREADS = (YS_W_EPI + 4 * YS_EPI_GROUPS) * 32;
static_assert(YS_SEG * 64 == 512, "the output rows of an item own the whole TMEM: slot == output row");

struct YsParams {
  const void* x;
  const void* residual;
  void* y;
  const uint8_t* w_packed;     // live taps only, 8 KB each (pack_weights, tile 64 x 64, SWIZZLE_128B rows)
  const int32_t* kblk;         // tap index ky*3+kx of every packed tile
  const float* scale;
  const float* shift;
  int n_kb, N, H, W, relu_n, res_pitch, res_coff;
  int tiles_x, tiles_y, total_tiles;
  uint32_t magic_x, magic_y;
  uint32_t idesc[3];           // N = 64, 128, 192
};

struct __align__(16) YsSync {
  uint64_t h_full[8], h_empty[8], t_full[YS_SEG], t_empty[YS_SEG], w_full, r_full[YS_EPI_GROUPS];
  uint32_t tmem_base, pad[3];
  alignas(16) float scale[64];
  alignas(16) float shift[64];
};

struct YsTile { int n, ox0, oy0; };
__device__ __forceinline__ YsTile ys_decode(const YsParams& p, int t) {
  YsTile c;
  const int q1 = p.tiles_x == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_x);
  const int txi = t - q1 * p.tiles_x;
  c.n = p.tiles_y == 1 ? q1 : (int)__umulhi((uint32_t)q1, p.magic_y);
  const int tyi = q1 - c.n * p.tiles_y;
  c.ox0 = txi * YS_W; c.oy0 = tyi * YS_SEG;
  return c;
}

// K-major SWIZZLE_128B operand descriptor without the start address: 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t ys_desc_hi() {
  uint64_t d = 0;
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;        // SWIZZLE_128B
  return d;
}

template <int DT, bool HAS_RES>
__global__ void __launch_bounds__(YS_THREADS, 1)
conv_ys_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
               const __grid_constant__ CUtensorMap tmap_r, const YsParams p) {
  constexpr int YS_RING = ys_ring(HAS_RES);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* rows = smem;                                    // YS_RING x YS_SLOT
  uint8_t* wsm = smem + (size_t)YS_RING * YS_SLOT;         // 3 x [192 rows][128 B]
  uint8_t* stage = wsm + YS_WBYTES;                        // YS_EPI_GROUPS x YS_STAGE (1024-aligned): finished rows
  uint8_t* resb = stage + YS_EPI_GROUPS * YS_STAGE;        // HAS_RES: YS_EPI_GROUPS x YS_STAGE: residual rows
  YsSync* sync = reinterpret_cast<YsSync*>(resb + (HAS_RES ? YS_EPI_GROUPS * YS_STAGE : 0));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch();
  if (tid == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    if (HAS_RES) tma_prefetch_desc(&tmap_r);
    for (int b = 0; b < YS_EPI_GROUPS; ++b) mbar_init(&sync->r_full[b], 1);
    for (int b = 0; b < YS_RING; ++b) { mbar_init(&sync->h_full[b], 1); mbar_init(&sync->h_empty[b], 1); }
    for (int b = 0; b < YS_SEG; ++b) { mbar_init(&sync->t_full[b], 1); mbar_init(&sync->t_empty[b], 4); }
    mbar_init(&sync->w_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sync->tmem_base, 512);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  if (tid == 0) {
    // resident weights: block (kx, j = 2 - ky) <- the packed tile of tap ky*3+kx, one 8 KB bulk copy per live tap
    // (asynchronous: they land while the first input rows are on their way; the MMA warp waits on w_full)
    mbar_arrive_expect_tx(&sync->w_full, (uint32_t)p.n_kb * YS_WTAP);
    for (int k = 0; k < p.n_kb; ++k) {
      const int tap = __ldg(p.kblk + k), ky = tap / 3, kx = tap - ky * 3;
      bulk_load(p.w_packed + (size_t)k * YS_WTAP, &sync->w_full, wsm + (size_t)(kx * 3 + 2 - ky) * YS_WTAP, YS_WTAP);
    }
  }
  {  // pruned taps are zero blocks of the stacks
    uint32_t live = 0;
    for (int k = 0; k < p.n_kb; ++k) {
      const int tap = __ldg(p.kblk + k), ky = tap / 3, kx = tap - ky * 3;
      live |= 1u << (kx * 3 + 2 - ky);
    }
    for (int blk = 0; blk < 9; ++blk)
      if (!((live >> blk) & 1u))
        for (int i = tid; i < (int)(YS_WTAP / 16); i += YS_THREADS)
          reinterpret_cast<uint4*>(wsm + (size_t)blk * YS_WTAP)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < 64) { sync->scale[tid] = __ldg(p.scale + tid); sync->shift[tid] = __ldg(p.shift + tid); }
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
      const YsTile c = ys_decode(p, t);
      for (int e = 0; e < YS_ROWS; ++e) {
        mbar_wait(&sync->h_empty[b], bph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&sync->h_full[b], YS_SLOT);
          // tensor {64, W, H, N}; box {64, 136, 1, 1}; zero fill outside the image = the conv padding
          tma_load_4d(&tmap_x, &sync->h_full[b], rows + (size_t)b * YS_SLOT, 0, c.ox0 - 1, c.oy0 - 1 + e, c.n);
        }
        __syncwarp();
        if (++b == YS_RING) { b = 0; bph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform loop)
    const uint64_t d_hi = ys_desc_hi();
    const uint32_t w16 = smem_u32(wsm) >> 4;
    int b = 0;
    uint32_t bph = 0, item = 0;
    mbar_wait(&sync->w_full, 0);
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++item) {
#pragma unroll
      for (int e = 0; e < YS_ROWS; ++e) {
        mbar_wait(&sync->h_full[b], bph);
        // output row e starts with this input row: its TMEM slot was read out by the epilogue of the previous item
        if (e < YS_SEG) mbar_wait(&sync->t_empty[e], (item & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t h16 = smem_u32(rows + (size_t)b * YS_SLOT) >> 4;
        if (elect_one()) {
          const int yo_lo = e - 2 > 0 ? e - 2 : 0, yo_hi = e < YS_SEG - 1 ? e : YS_SEG - 1;
          // tap column kx, K-step ks -> output rows yo_a .. yo_a + nblk - 1 (64 columns each) through the rows of the
          // kx stack that start at block j0 (stack of ky = 2, 1, 0: output row yo takes ky = e - yo)
          auto mma = [&](int kx, int ks, int yo_a, int nblk, int j0, uint32_t accumulate) {
            umma_f16(tmem_base + (uint32_t)(yo_a * 64),
                     d_hi | (uint64_t)(h16 + ((uint32_t)kx * YS_PIX + 32u * (uint32_t)ks) / 16u),
                     d_hi | (uint64_t)(w16 + ((uint32_t)kx * YS_WKX + (uint32_t)j0 * YS_WTAP + 32u * (uint32_t)ks) / 16u),
                     p.idesc[nblk - 1], accumulate);
          };
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (kx == 0 && ks == 0 && e < YS_SEG) {
                mma(0, 0, e, 1, 2, 0u);                                   // ky = 0: first write of output row e
                if (e > yo_lo) mma(0, 0, yo_lo, e - yo_lo, 2 - (e - yo_lo), 1u);
              } else {
                mma(kx, ks, yo_lo, yo_hi - yo_lo + 1, 2 - (e - yo_lo), 1u);
              }
            }
          }
          umma_commit(&sync->h_empty[b]);
          if (e >= 2) umma_commit(&sync->t_full[e - 2]);                  // output row e - 2 has its three filter rows
        }
        __syncwarp();
        if (++b == YS_RING) { b = 0; bph ^= 1u; }
      }
    }
  } else {
    // ===================================================================== epilogue: thread = pixel of the row tile
    const int q = warp & 3;
    const int grp = (warp - YS_W_EPI) >> 2;
    const int m = q * 32 + lane;
    uint8_t* stg = stage + (size_t)grp * YS_STAGE;
    const uint32_t srow = (uint32_t)m * YS_PIX, sx = (uint32_t)(m & 7);   // SWIZZLE_128B: chunk j of row m sits at j ^ (m & 7)
    const bool issuer = (q == 0 && lane == 0);
    uint8_t* rbuf = resb + (size_t)grp * YS_STAGE;
    const bool relu_all = p.relu_n >= 64;
    // this group's rows as one sequence k = 0, 1, ...: item k / 4, output row grp + 2 * (k % 4).  The residual of row
    // k + 1 (128 B per pixel) is requested before row k is finished, so that its latency hides behind a whole row.
    constexpr int RPI = YS_SEG / YS_EPI_GROUPS;              // rows per item and group
    int sx0 = 0, sy = 0, sn = 0, sx0_n = 0, sy_n = 0, sn_n = 0;   // TMA coordinates of the current / next row
    auto locate = [&](int k, int& yo, int& cx, int& cy, int& cn) -> bool {   // false: past this CTA's last item
      const int t = blockIdx.x + (k / RPI) * gridDim.x;
      yo = grp + YS_EPI_GROUPS * (k % RPI);
      if (t >= p.total_tiles) return false;
      const YsTile c = ys_decode(p, t);
      cx = c.ox0; cy = c.oy0 + yo; cn = c.n;
      return true;
    };
    uint32_t rv[4][8];
    // residual row of output row (cx, cy, cn): tensor {64, W, H, N} over channels [res_coff, res_coff + 64) of the
    // residual tensor; box {64, 128, 1, 1}; rows below the image are not fetched (their result is not stored either)
    auto request = [&](bool in_range, int cx, int cy, int cn) {
      if (in_range && cy < p.H) {
        mbar_arrive_expect_tx(&sync->r_full[grp], YS_STAGE);
        tma_load_4d(&tmap_r, &sync->r_full[grp], rbuf, 0, cx, cy, cn);
      } else {
        mbar_arrive(&sync->r_full[grp]);
      }
    };
    int yo = 0, yo_n = 0;
    const bool any = locate(0, yo, sx0, sy, sn);              // grid <= items: every CTA has at least one
    if (HAS_RES && issuer) request(any, sx0, sy, sn);
    const int n_rows = ((p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * RPI;
#pragma unroll 1
    for (int k = 0; k < n_rows; ++k) {
      const uint32_t item = (uint32_t)(k / RPI);
      {
        locate(k + 1, yo_n, sx0_n, sy_n, sn_n);
        if (HAS_RES) {
          mbar_wait(&sync->r_full[grp], (uint32_t)k & 1u);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 lo = *reinterpret_cast<const uint4*>(rbuf + srow + ((((uint32_t)(2 * g)) ^ sx) << 4));
            const uint4 hi = *reinterpret_cast<const uint4*>(rbuf + srow + ((((uint32_t)(2 * g + 1)) ^ sx) << 4));
            rv[g][0] = lo.x; rv[g][1] = lo.y; rv[g][2] = lo.z; rv[g][3] = lo.w;
            rv[g][4] = hi.x; rv[g][5] = hi.y; rv[g][6] = hi.z; rv[g][7] = hi.w;
          }
          named_bar_sync(1 + grp, 128);        // everybody has its pixel: the buffer may take the next row's residual
          if (issuer && k + 1 < n_rows) request(true, sx0_n, sy_n, sn_n);
        }
        mbar_wait(&sync->t_full[yo], item & 1u);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (uint32_t)(yo * 64) + ((uint32_t)(q * 32) << 16);
        uint32_t v[2][32];
        tmem_ld32(t_addr, v[0]);
        tmem_ld32(t_addr + 32u, v[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sync->t_empty[yo]);       // accumulator slot read out: the next item may reuse it
        // the TMA store that last read this group's staging buffer has finished reading it
        if (issuer) bulk_wait_group_read<0>();
        named_bar_sync(1 + grp, 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[8];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const int ch = 16 * g + 4 * e4;
            const float4 sc = *reinterpret_cast<const float4*>(&sync->scale[ch]);
            const float4 sh = *reinterpret_cast<const float4*>(&sync->shift[ch]);
            float a0 = fmaf(__uint_as_float(v[g >> 1][(ch & 31)]), sc.x, sh.x);
            float a1 = fmaf(__uint_as_float(v[g >> 1][(ch & 31) + 1]), sc.y, sh.y);
            float a2 = fmaf(__uint_as_float(v[g >> 1][(ch & 31) + 2]), sc.z, sh.z);
            float a3 = fmaf(__uint_as_float(v[g >> 1][(ch & 31) + 3]), sc.w, sh.w);
            if (HAS_RES) {
              a0 += Act<DT>::to_f32((uint16_t)(rv[g][2 * e4] & 0xFFFFu));
              a1 += Act<DT>::to_f32((uint16_t)(rv[g][2 * e4] >> 16));
              a2 += Act<DT>::to_f32((uint16_t)(rv[g][2 * e4 + 1] & 0xFFFFu));
              a3 += Act<DT>::to_f32((uint16_t)(rv[g][2 * e4 + 1] >> 16));
            }
            if (relu_all || ch < p.relu_n) a0 = fmaxf(a0, 0.f);
            if (relu_all || ch + 1 < p.relu_n) a1 = fmaxf(a1, 0.f);
            if (relu_all || ch + 2 < p.relu_n) a2 = fmaxf(a2, 0.f);
            if (relu_all || ch + 3 < p.relu_n) a3 = fmaxf(a3, 0.f);
            w[2 * e4] = pack2<DT>(a0, a1);
            w[2 * e4 + 1] = pack2<DT>(a2, a3);
          }
          *reinterpret_cast<uint4*>(stg + srow + ((((uint32_t)(2 * g)) ^ sx) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(stg + srow + ((((uint32_t)(2 * g + 1)) ^ sx) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + grp, 128);
        if (issuer) {
          // tensor {64, W, H, N}; box {64, 128, 1, 1}: pixels right of the image are clipped, rows below it skipped
          if (sy < p.H) tma_store_4d(&tmap_y, stg, 0, sx0, sy, sn);
          bulk_commit_group();
        }
      }
      yo = yo_n; sx0 = sx0_n; sy = sy_n; sn = sn_n;
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

struct YsMapCache {
  const void* ptr = nullptr;
  const void* ptr_y = nullptr;
  const void* ptr_r = nullptr;
  CUtensorMap map, map_y, map_r;
};

static PFN_cuTensorMapEncodeTiled_v12000 ys_encode_fn() {
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

bool conv_ys_supported(const drnb200_conv_desc& d) {
  static const char* env = getenv("DRNB200_YS");            // A/B knob: "0" keeps conv_halo for these layers
  if (env && env[0] == '0') return false;
  return d.ksize == 3 && d.stride == 1 && d.dilation == 1 && d.Cin == 64 && d.tile_ci == 64 && d.Cout == 64 &&
         d.tile_o == 64 && !d.out_f32 && d.W >= 8 &&
         (!d.has_residual || (d.res_cpitch % 16 == 0 && d.res_coffset % 16 == 0));   // 32-byte residual accesses
}

int conv_ys_launch(drnb200_conv_plan* plan, cudaStream_t st) {
  const ConvParams& c = plan->p;
  const drnb200_conv_desc& d = plan->d;
  YsParams p{};
  p.x = c.x; p.residual = c.residual; p.y = c.y; p.w_packed = c.w_packed; p.kblk = c.kblk;
  p.scale = c.scale; p.shift = c.shift;
  p.n_kb = plan->h_row_ptr[1];
  if (p.n_kb == 0) return conv_direct_launch(plan, st);     // everything pruned: y = act(shift + res)
  p.N = c.N; p.H = c.H; p.W = c.W; p.relu_n = c.relu_n; p.res_pitch = c.res_pitch; p.res_coff = c.res_coff;
  p.tiles_x = (c.W + YS_W - 1) / YS_W;
  p.tiles_y = (c.H + YS_SEG - 1) / YS_SEG;
  p.total_tiles = c.N * p.tiles_x * p.tiles_y;
  if ((uint64_t)p.total_tiles * (uint64_t)std::max(p.tiles_x, p.tiles_y) >= (1ull << 32)) {
    set_error("conv_ys: problem too large for the 32-bit tile decode");
    return DRNB200_E_ARG;
  }
  p.magic_x = p.tiles_x == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_x - 1) / p.tiles_x);
  p.magic_y = p.tiles_y == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_y - 1) / p.tiles_y);
  for (int n = 0; n < 3; ++n) p.idesc[n] = umma_idesc_f16(128, 64 * (n + 1), d.act_dtype);
  const size_t smem = 1024 + (size_t)ys_ring(c.has_res != 0) * YS_SLOT + YS_WBYTES +
                      (c.has_res ? 2 : 1) * YS_EPI_GROUPS * YS_STAGE + sizeof(YsSync);

  static_assert(sizeof(YsMapCache) <= sizeof(plan->gather_cache), "tensor-map cache storage too small");
  YsMapCache* cache = reinterpret_cast<YsMapCache*>(plan->gather_cache);
  if (!plan->gather_cache_init) { new (cache) YsMapCache(); plan->gather_cache_init = true; }
  if (cache->ptr != p.x) {
    auto fn = ys_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    cuuint64_t gdim[4] = {64, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {(cuuint64_t)c.x_cpitch * 2, (cuuint64_t)c.W * c.x_cpitch * 2,
                          (cuuint64_t)c.H * c.W * c.x_cpitch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)YS_HP, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&cache->map, d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    4, const_cast<void*>(p.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(conv_ys) failed with CUresult %d (W=%d H=%d N=%d)", (int)r, c.W, c.H, c.N);
      return DRNB200_E_CUDA;
    }
    cache->ptr = p.x;
  }
  if (cache->ptr_y != p.y) {
    auto fn = ys_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    cuuint64_t gdim[4] = {64, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {128, (cuuint64_t)c.W * 128, (cuuint64_t)c.H * c.W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)YS_W, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&cache->map_y, d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    4, p.y, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(conv_ys output) failed with CUresult %d (W=%d H=%d N=%d)", (int)r, c.W, c.H, c.N);
      return DRNB200_E_CUDA;
    }
    cache->ptr_y = p.y;
  }
  if (c.has_res && cache->ptr_r != p.residual) {
    auto fn = ys_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    cuuint64_t gdim[4] = {64, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {(cuuint64_t)c.res_pitch * 2, (cuuint64_t)c.W * c.res_pitch * 2,
                          (cuuint64_t)c.H * c.W * c.res_pitch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)YS_W, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    void* base = const_cast<uint16_t*>(reinterpret_cast<const uint16_t*>(p.residual) + c.res_coff);
    CUresult r = fn(&cache->map_r, d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(conv_ys residual) failed with CUresult %d (W=%d H=%d N=%d pitch=%d)", (int)r,
                c.W, c.H, c.N, c.res_pitch);
      return DRNB200_E_CUDA;
    }
    cache->ptr_r = p.residual;
  }
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_tiles, sms);
  if (grid == 0) return DRNB200_OK;
#define DRN_YS_LAUNCH(DT, RES)                                                                              \
  do {                                                                                                      \
    static std::atomic<unsigned long long> attr;                                                            \
    if (attr_needed_on_this_device(attr))                                                                   \
      DRN_CUDA(cudaFuncSetAttribute(conv_ys_kernel<DT, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                    (int)smem));                                                            \
    launch_chained(conv_ys_kernel<DT, RES>, grid, YS_THREADS, smem, st, cache->map, cache->map_y,                \
                   RES ? cache->map_r : cache->map_y, p);                     \
  } while (0)
  if (d.act_dtype == DRNB200_BF16) {
    if (c.has_res) DRN_YS_LAUNCH(DRNB200_BF16, true); else DRN_YS_LAUNCH(DRNB200_BF16, false);
  } else {
    if (c.has_res) DRN_YS_LAUNCH(DRNB200_F16, true); else DRN_YS_LAUNCH(DRNB200_F16, false);
  }
#undef DRN_YS_LAUNCH
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

}  // namespace drnb200
