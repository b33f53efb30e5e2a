// conv_halo.cu — stride-1 3x3 convolutions with Cin <= 64 and Cout <= 64 where the tensor core reads the
// filter taps as SHIFTED WINDOWS of one shared-memory halo tile (no im2col copy, no per-tap reload).
//
// Why: for the narrow layers (DRN layer1: 16->16 at full resolution, layer3: 64->64 at 1/4) the per-tap
// TMA path of conv_tc.cu moves every input pixel L2->SM nine times and is bound by that traffic
// (layer3: 0.23 ms per conv at batch 8 vs a 0.04 ms HBM floor).  Here one TMA box per tile brings the halo
//   [16+2d rows][24 pixels][Cin channels]      (pixel rows of `pitch` = Cin*2 bytes, SWIZZLE_{32,64,128}B)
// and tap (ky,kx) of M-block mb (8x16 pixels) of the 16x16-pixel output tile is the UMMA A operand whose descriptor
// simply starts (ky*d*24 + kx*d + 8*mb) pixel rows further: M row m = (ty, tx) -> 8-row group ty (stride = one halo
// row = 24*pitch bytes, a multiple of the swizzle atom) and row tx inside it.  Measured on B200: the UMMA swizzle
// XOR is taken from the ABSOLUTE shared-memory address bits (like TMA's), so a start that is not atom
// aligned needs no compensation — the descriptor base-offset field must stay 0 (setting it to
// (start >> 7) & 7 gives wrong results; DRNB200_HALO=1 reproduces that for the record).
// Weights stay resident in shared memory for the whole CTA.
// Orientation as MODE_P: M = 128 pixels (TMEM lanes), N = Cout (columns), K-step 16.
// Roles (320 threads): warp 0 halo TMA producer, warp 1 TMEM alloc + MMA issue, warps 2-9 epilogue
// (two groups of four warps take alternate tiles; BN affine + residual + ReLU, 16-byte vector accesses,
// thread = pixel).  Four TMEM accumulator stages keep both groups and the MMA issuer busy.
#include "conv_internal.cuh"
#include <algorithm>
#include <cudaTypedefs.h>
#include <new>

namespace drnb200 {

#ifndef DRN_H_EPI_GROUPS
#define DRN_H_EPI_GROUPS 2
#endif
#ifndef DRN_H_MMA_WARPS
#define DRN_H_MMA_WARPS 3
#endif
#ifndef DRN_H_ACC
#define DRN_H_ACC 8
#endif
constexpr int H_EPI_GROUPS = DRN_H_EPI_GROUPS;         // epilogue groups of four warps, one tile each, round-robin
constexpr int H_MMA_WARPS = DRN_H_MMA_WARPS;          // MMA-issuing warps on alternate tiles: the per-tile wait/commit
                                        // latency of one issuer (~800 cycles) bounded the small layers
constexpr int H_W_EPI = 1 + H_MMA_WARPS;
constexpr int H_THREADS = (H_W_EPI + 4 * H_EPI_GROUPS) * 32;
constexpr int H_MB = 2;                 // UMMA M-blocks (8 x 16 pixels each) per tile, side by side in x: the per-tile
                                        // barrier/commit skeleton alone costs ~320 cycles (measured with every load, MMA
                                        // and epilogue body switched off), so a tile carries 256 pixels, not 128
constexpr int H_TW = 8 * H_MB, H_TH = 16;   // output tile (pixels)
constexpr int H_WP = H_TW + 8;          // halo row length in pixels (H_TW + 2*dil, dil <= 4), a multiple of 8
constexpr int H_MAX_RING = 6;
constexpr int H_ACC = DRN_H_ACC;                // TMEM accumulator stages: the MMA -> epilogue -> MMA round trip is
                                        // ~3000 cycles, far longer than a tile of these small layers
constexpr uint32_t H_TMEM_COLS = H_ACC * 64;   // accumulators x 64 columns

struct HaloParams {
  const void* x;
  const void* residual;
  void* y;
  const uint8_t* w_packed;   // n_kb tiles of Cout x Cin, K-major swizzled rows of `pitch` bytes
  const int32_t* kblk;       // live taps
  const float* scale;
  const float* shift;
  int n_kb, N, H, W, Cin, Cout, dil, relu_n, has_res, x_cpitch, res_pitch, res_coff;
  int tiles_x, tiles_y, total_tiles, halo_h, ring;
  int n_mma;                 // MMA-issuing warps in use: min(H_MMA_WARPS, ring) (each owns every n_mma-th ring slot)
  uint32_t magic_x, magic_y;   // ceil(2^32 / tiles_{x,y}): exact quotients by __umulhi for t < 2^32 / divisor
  uint32_t pitch, halo_bytes, halo_tx, w_tile_bytes, idesc, base_off_mode;   // halo_bytes: ring slot stride, halo_tx: box bytes
  int dbg;                   // diagnostics (DRNB200_DBG): 1 = issue one tap only, 2 = skip the global stores,
                             // 4 = load two halo rows only (is the halo TMA the bound?); bits 8 / 16 / 32 = skip the
                             // epilogue body / all MMAs / the halo loads (pipeline skeleton)
};

struct __align__(16) HSync {
  uint64_t h_full[H_MAX_RING], h_empty[H_MAX_RING], t_full[H_ACC], t_empty[H_ACC], w_full;
  uint32_t tmem_base;
  int32_t taps[9];
  alignas(16) float scale[64];
  alignas(16) float shift[64];
};

struct HTile { int n, ox0, oy0; };
__device__ __forceinline__ HTile h_decode(const HaloParams& p, int t) {
  HTile c;
  const int q1 = p.tiles_x == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_x);      // t / tiles_x
  const int txi = t - q1 * p.tiles_x;
  c.n = p.tiles_y == 1 ? q1 : (int)__umulhi((uint32_t)q1, p.magic_y);             // q1 / tiles_y
  const int tyi = q1 - c.n * p.tiles_y;
  c.ox0 = txi * H_TW; c.oy0 = tyi * H_TH;
  return c;
}

// K-major swizzled descriptor with explicit 8-row-group stride and base offset
__device__ __forceinline__ uint64_t umma_desc_ex(uint32_t addr, uint32_t pitch, uint32_t sbo, uint32_t base_off) {
  const uint64_t layout = (pitch == 128u) ? 2ull : (pitch == 64u ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= layout << 61;
  return d;
}

template <int DT, int KSTEPS, bool HAS_RES>
__global__ void __launch_bounds__(H_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* halo = smem;                                            // ring x halo_bytes (1024-multiple)
  uint8_t* wsm = smem + (size_t)p.ring * p.halo_bytes;             // n_kb x w_tile_bytes
  HSync* sync = reinterpret_cast<HSync*>(wsm + ((9u * p.w_tile_bytes + 1023u) & ~1023u));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int b = 0; b < p.ring; ++b) { mbar_init(&sync->h_full[b], 1); mbar_init(&sync->h_empty[b], 1); }
    for (int a = 0; a < H_ACC; ++a) { mbar_init(&sync->t_full[a], 1); mbar_init(&sync->t_empty[a], 4); }
    mbar_init(&sync->w_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sync->tmem_base, H_TMEM_COLS);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  if (threadIdx.x < 9) sync->taps[threadIdx.x] = (int)threadIdx.x < p.n_kb ? __ldg(p.kblk + threadIdx.x) : 0;
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 64) {     // BN affine -> shared memory (read by every pixel thread)
    const int ch = threadIdx.x - 64;
    sync->scale[ch] = ch < p.Cout ? __ldg(p.scale + ch) : 0.f;
    sync->shift[ch] = ch < p.Cout ? __ldg(p.shift + ch) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sync->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    {
      // warp-uniform loop, single elected lane issues (see the MMA warp for why)
      if (elect_one()) {     // resident weights: one bulk copy
        mbar_arrive_expect_tx(&sync->w_full, (uint32_t)p.n_kb * p.w_tile_bytes);
        bulk_load(p.w_packed, &sync->w_full, wsm, (uint32_t)p.n_kb * p.w_tile_bytes);
      }
      int b = 0;
      uint32_t bph = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const HTile c = h_decode(p, t);
        mbar_wait(&sync->h_empty[b], bph ^ 1u);
        if (p.dbg & 32) {
          if (elect_one()) mbar_arrive(&sync->h_full[b]);
        } else if (elect_one()) {
          mbar_arrive_expect_tx(&sync->h_full[b], p.dbg == 4 ? 2u * H_WP * p.pitch : p.halo_tx);
          // tensor {Cin, W, H, N}; box {Cin, 16, 16+2d, 1}; zero fill outside the image = conv padding
          tma_load_4d(&tmap_x, &sync->h_full[b], halo + (size_t)b * p.halo_bytes, 0, c.ox0 - p.dil,
                      c.oy0 - p.dil, c.n);
        }
        __syncwarp();
        if (++b == p.ring) { b = 0; bph ^= 1u; }
      }
    }
  } else if (warp < H_W_EPI) {
    // ===================================================================== MMA issuers (warps 1..H_MMA_WARPS)
    // The whole warp runs this loop with warp-uniform control flow and values; only the tcgen05 issue is
    // predicated on one elected lane.  Inside an `if (lane == 0)` region nothing is provably uniform, so
    // every descriptor went through R2UR (measured ~49 cycles per MMA + ~1000 cycles per tile); here the
    // descriptors live in uniform registers.
    {
      mbar_wait(&sync->w_full, 0);
      const uint32_t row_bytes = (uint32_t)H_WP * p.pitch;          // one halo row = 8-row-group stride
      const uint64_t a_hi = umma_desc_ex(0u, p.pitch, row_bytes, 0u);
      const uint64_t b_hi = umma_desc_ex(0u, p.pitch, 8u * p.pitch, 0u);
      const uint32_t w16 = smem_u32(wsm) >> 4, wt16 = p.w_tile_bytes >> 4;
      uint32_t a_off16[9];
#pragma unroll
      for (int kb = 0; kb < 9; ++kb) {
        const int tap = sync->taps[kb];
        const int ky = tap / 3, kx = tap - ky * 3;
        a_off16[kb] = ((uint32_t)(ky * p.dil) * row_bytes + (uint32_t)(kx * p.dil) * p.pitch) >> 4;
      }
      const int n_kb = p.dbg == 1 ? 1 : ((p.dbg & 16) ? 0 : p.n_kb);
      const uint32_t mb16 = (8u * p.pitch) >> 4;             // M-block mb starts 8 pixels further right
      const int mw = warp - 1;                               // this warp takes tiles i = mw, mw + H_MMA_WARPS, ...
      int i = mw, b = mw;                                    // mw < n_mma <= ring
      uint32_t bph = 0;
      for (int t = blockIdx.x + mw * gridDim.x; mw < p.n_mma && t < p.total_tiles;
           t += p.n_mma * gridDim.x, i += p.n_mma) {
        const int acc = (i * H_MB) % H_ACC;                  // tile i owns accumulators acc .. acc + H_MB - 1
        mbar_wait(&sync->h_full[b], bph);
        mbar_wait(&sync->t_empty[acc], ((uint32_t)(i * H_MB / H_ACC) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t h16 = smem_u32(halo + (size_t)b * p.halo_bytes) >> 4;
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 64u;
        if (elect_one()) {
#pragma unroll
          for (int mb = 0; mb < H_MB; ++mb) {
#pragma unroll
            for (int kb = 0; kb < 9; ++kb) {
              if (kb < n_kb) {
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
                  umma_f16(d_tmem + (uint32_t)mb * 64u, a_hi | (uint64_t)(h16 + a_off16[kb] + mb * mb16 + 2u * ks),
                           b_hi | (uint64_t)(w16 + (uint32_t)kb * wt16 + 2u * ks), p.idesc,
                           (kb > 0 || ks > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(&sync->h_empty[b]);
          umma_commit(&sync->t_full[acc]);
        }
        __syncwarp();
        b += p.n_mma;
        if (b >= p.ring) { b -= p.ring; bph ^= 1u; }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;
    const int grp = (warp - H_W_EPI) >> 2;          // epilogue group g takes tiles i = g, g + H_EPI_GROUPS, ...
    const int m = q * 32 + lane;                    // TMEM lane = pixel (ty = m / 8, tx = m % 8)
    const uint16_t* res16 = reinterpret_cast<const uint16_t*>(p.residual);
    uint16_t* y16 = reinterpret_cast<uint16_t*>(p.y);
    // One tile = H_MB M-blocks = H_MB pixels per thread; the M-blocks' instruction streams run interleaved (one warp
    // finishing one M-block is a ~200-instruction dependent chain).
    static_assert(H_MB == 2 && H_ACC % H_MB == 0, "the epilogue below is written for two M-blocks per tile");
    const bool relu_all = p.relu_n >= p.Cout;       // the usual case: one warp-uniform branch instead of 16 predicates
    constexpr bool has1 = true;
    for (int i = grp, t = blockIdx.x + grp * gridDim.x; t < p.total_tiles;
         t += H_EPI_GROUPS * gridDim.x, i += H_EPI_GROUPS) {
      const int acc = (i * H_MB) % H_ACC;
      const uint32_t tph = (uint32_t)(i * H_MB / H_ACC) & 1u;
      const HTile c = h_decode(p, t);
      bool valid[2];
      size_t pix0[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ox = c.ox0 + 8 * u + (m & 7), oy = c.oy0 + (m >> 3);
        valid[u] = ox < p.W && oy < p.H;
        pix0[u] = ((size_t)c.n * p.H + oy) * p.W + ox;
      }
      // the residual of this pixel (<= 128 B) is requested before the accumulator wait so that its
      // latency overlaps the MMAs instead of serialising behind every 16-channel group
      uint32_t rv[2][4][8];                           // 32-byte (16-channel) pieces: one full sector per access
      if (HAS_RES) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int e = 0; e < 8; ++e) rv[u][g][e] = 0u;
            if (valid[u] && g * 16 < p.Cout)
              ldg256_nc(res16 + pix0[u] * p.res_pitch + p.res_coff + g * 16, rv[u][g]);
          }
      }
      mbar_wait(&sync->t_full[acc], tph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * 64u + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int cbi = 0; cbi < 4; ++cbi) {
        const int cb = cbi * 16;
        if (cb < p.Cout && !(p.dbg & 8)) {
          uint32_t v[2][16];
          tmem_ld16(t_addr + (uint32_t)cb, v[0]);
          if (has1) tmem_ld16(t_addr + 64u + (uint32_t)cb, v[1]);
          tmem_ld_wait();
          float4 sc[4], sh[4];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            sc[e4] = *reinterpret_cast<const float4*>(&sync->scale[cb + 4 * e4]);
            sh[e4] = *reinterpret_cast<const float4*>(&sync->shift[cb + 4 * e4]);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (valid[u]) {
              float f[16];
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                f[4 * e4] = fmaf(__uint_as_float(v[u][4 * e4]), sc[e4].x, sh[e4].x);
                f[4 * e4 + 1] = fmaf(__uint_as_float(v[u][4 * e4 + 1]), sc[e4].y, sh[e4].y);
                f[4 * e4 + 2] = fmaf(__uint_as_float(v[u][4 * e4 + 2]), sc[e4].z, sh[e4].z);
                f[4 * e4 + 3] = fmaf(__uint_as_float(v[u][4 * e4 + 3]), sc[e4].w, sh[e4].w);
              }
              if (HAS_RES) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  f[2 * e] += Act<DT>::to_f32((uint16_t)(rv[u][cbi][e] & 0xFFFFu));
                  f[2 * e + 1] += Act<DT>::to_f32((uint16_t)(rv[u][cbi][e] >> 16));
                }
              }
              if (relu_all) {
#pragma unroll
                for (int e = 0; e < 16; ++e) f[e] = fmaxf(f[e], 0.f);
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                  if (cb + e < p.relu_n) f[e] = fmaxf(f[e], 0.f);
              }
              uint32_t w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) w[e] = pack2<DT>(f[2 * e], f[2 * e + 1]);
              if (p.dbg != 2 || w[0] == 0x12345678u) stg256(y16 + pix0[u] * p.Cout + cb, w);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sync->t_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, H_TMEM_COLS);
  }
}

struct HMapCache {
  const void* ptr = nullptr;
  CUtensorMap map;
};

static PFN_cuTensorMapEncodeTiled_v12000 h_encode_fn() {
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

bool conv_halo_supported(const drnb200_conv_desc& d) {
  // DRNB200_HALO=0 disables the path (A/B measurements); in -DDRNB200_DIAG builds 1 = set the descriptor base offset
  // (reproduces the wrong-swizzle finding of DESIGN.md section 3)
  static const char* env = getenv("DRNB200_HALO");
  if (env && env[0] == '0') return false;
  return d.ksize == 3 && d.stride == 1 && d.dilation >= 1 && d.dilation <= 4 && d.Cin == d.tile_ci &&
         (d.Cin == 16 || d.Cin == 32 || d.Cin == 64) && d.tile_o == d.Cout && d.Cout % 16 == 0 && d.Cout <= 64 &&
         !d.out_f32;
}

int conv_halo_launch(drnb200_conv_plan* plan, cudaStream_t st) {
  const ConvParams& c = plan->p;
  const drnb200_conv_desc& d = plan->d;
  HaloParams p{};
  p.x = c.x; p.residual = c.residual; p.y = c.y; p.w_packed = c.w_packed; p.kblk = c.kblk;
  p.scale = c.scale; p.shift = c.shift;
  p.n_kb = plan->h_row_ptr[1];
  if (p.n_kb == 0) return conv_direct_launch(plan, st);   // everything pruned: y = act(shift + res)
  p.N = c.N; p.H = c.H; p.W = c.W; p.Cin = c.Cin; p.Cout = c.Cout; p.dil = c.dil; p.relu_n = c.relu_n;
  p.has_res = c.has_res; p.x_cpitch = c.x_cpitch; p.res_pitch = c.res_pitch; p.res_coff = c.res_coff;
  p.pitch = (uint32_t)c.Cin * 2u;
  p.halo_h = H_TH + 2 * c.dil;
  p.halo_tx = (uint32_t)p.halo_h * H_WP * p.pitch;
  p.halo_bytes = (p.halo_tx + 1023u) & ~1023u;
  p.w_tile_bytes = (uint32_t)c.Cout * p.pitch;
  p.tiles_x = (c.W + H_TW - 1) / H_TW;
  p.tiles_y = (c.H + H_TH - 1) / H_TH;
  p.total_tiles = c.N * p.tiles_x * p.tiles_y;
  if ((uint64_t)p.total_tiles * (uint64_t)std::max(p.tiles_x, p.tiles_y) >= (1ull << 32)) {
    set_error("conv_halo: problem too large for the 32-bit tile decode");
    return DRNB200_E_ARG;
  }
  p.magic_x = p.tiles_x == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_x - 1) / p.tiles_x);
  p.magic_y = p.tiles_y == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_y - 1) / p.tiles_y);
  p.idesc = umma_idesc_f16(128, c.Cout, d.act_dtype);
  static const int env_halo = diag_env("DRNB200_HALO");
  p.base_off_mode = env_halo == 1 ? 1u : 0u;
  static const int envd = diag_env("DRNB200_DBG");
  p.dbg = envd;
  const size_t kMaxSmem = 232448;
  const size_t fixed = 1024 + ((9u * p.w_tile_bytes + 1023u) & ~1023u) + sizeof(HSync);
  p.ring = (int)std::min<size_t>(H_MAX_RING, (kMaxSmem - fixed) / p.halo_bytes);
  if (p.ring < 2) { set_error("conv_halo: halo tile does not fit shared memory"); return DRNB200_E_ARG; }
  // MMA warps take alternate tiles; tile i uses ring slot i % ring and accumulator set i % (H_ACC / H_MB).  Successive
  // phases of one barrier must be awaited by the SAME warp: a second warp polling "next fill of slot b" while the
  // current fill is still in flight is answered "done" by mbarrier.try_wait.parity (the barrier is two phases behind
  // the question) — found in conv_ty.cu, where it faulted 1-2 % of the launches.  So n_mma divides both counts.
  p.n_mma = 1;
  for (int n = std::min(H_MMA_WARPS, p.ring); n > 1; --n)
    if (p.ring % n == 0 && (H_ACC / H_MB) % n == 0) { p.n_mma = n; break; }

  static_assert(sizeof(HMapCache) <= sizeof(plan->gather_cache), "halo cache storage too small");
  HMapCache* cache = reinterpret_cast<HMapCache*>(plan->gather_cache);
  if (!plan->gather_cache_init) { new (cache) HMapCache(); plan->gather_cache_init = true; }
  if (cache->ptr != p.x) {
    auto fn = h_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    cuuint64_t gdim[4] = {(cuuint64_t)c.Cin, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.N};
    cuuint64_t gstr[3] = {(cuuint64_t)c.x_cpitch * 2, (cuuint64_t)c.W * c.x_cpitch * 2,
                          (cuuint64_t)c.H * c.W * c.x_cpitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)c.Cin, (cuuint32_t)H_WP, (cuuint32_t)(p.dbg == 4 ? 2 : p.halo_h), 1};   // dbg 4: timing probe
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = p.pitch == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : p.pitch == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(&cache->map, d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    4, const_cast<void*>(p.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(halo) failed with CUresult %d (Cin=%d W=%d H=%d N=%d)", (int)r, c.Cin, c.W,
                c.H, c.N);
      return DRNB200_E_CUDA;
    }
    cache->ptr = p.x;
  }
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_tiles, sms);
  if (grid == 0) return DRNB200_OK;
  const int ks = (int)(p.pitch / 32u);
#define DRN_HALO_LAUNCH1(DT, KS, RES)                                                                          \
  do {                                                                                                         \
    static std::atomic<unsigned long long> attr;                                                               \
    if (attr_needed_on_this_device(attr))                                                                      \
      DRN_CUDA(cudaFuncSetAttribute(conv_halo_kernel<DT, KS, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)kMaxSmem));                                                           \
    launch_chained(conv_halo_kernel<DT, KS, RES>, grid, H_THREADS, kMaxSmem, st, cache->map, p);                     \
  } while (0)
#define DRN_HALO_LAUNCH(DT, KS)                        \
  do {                                                 \
    if (p.has_res) DRN_HALO_LAUNCH1(DT, KS, true);     \
    else DRN_HALO_LAUNCH1(DT, KS, false);              \
  } while (0)
  if (d.act_dtype == DRNB200_BF16) {
    if (ks == 4) DRN_HALO_LAUNCH(DRNB200_BF16, 4);
    else if (ks == 2) DRN_HALO_LAUNCH(DRNB200_BF16, 2);
    else DRN_HALO_LAUNCH(DRNB200_BF16, 1);
  } else {
    if (ks == 4) DRN_HALO_LAUNCH(DRNB200_F16, 4);
    else if (ks == 2) DRN_HALO_LAUNCH(DRNB200_F16, 2);
    else DRN_HALO_LAUNCH(DRNB200_F16, 1);
  }
#undef DRN_HALO_LAUNCH
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

}  // namespace drnb200
