// conv_tc.cu — (b) block-sparse implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a).
//
// GEMM view of one conv layer (reference: tools/get_matrix_shapes.py:19-21, M=Cout, K=Cin*k*k, N=OH*OW):
//   K is walked as K-blocks kb = (cib, tap): `tile_ci` input channels of one filter tap.  Only the
//   K-blocks listed as live for the CTA's output-channel tile are loaded and multiplied (compact.cu).
//   activations: NHWC 16-bit; a K-block of a pixel tile is ONE 4-D TMA box {tile_ci, TW*s, TH*s, 1}
//                at (cib*tile_ci, ox0*s + (kx-1)*dil, oy0*s + (ky-1)*dil, n); out-of-bounds elements are
//                zero-filled by TMA, which is exactly the conv's zero padding (padding == dilation);
//                stride-2 layers use elementStrides = 2.
//   weights:     pre-packed per live K-block in the swizzled smem image -> one 1-D bulk copy.
// Two operand orientations share the pipeline:
//   MODE_T (Cout >= 128): A = weights (M = 128 couts), B = activations (N = NT <= 256 pixels).
//                         TMEM lane = cout, column = pixel.  The K list is per 128-cout tile, so the
//                         skipping granularity equals the UMMA M tile and N=256 keeps the MMA at full rate.
//   MODE_P (Cout <= 64):  A = activations (M = 128 pixels), B = weights (N = Cout).
//                         TMEM lane = pixel, column = cout (small-channel, bandwidth-bound layers).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-5 = epilogue (TMEM -> registers -> BN affine (+residual) (+ReLU) -> NHWC global).
// Two accumulator stages in TMEM (2 x 256 columns) let the epilogue of tile i overlap the MMAs of i+1.
// MODE_T epilogue ("staged"): the accumulator is cout-major (lane = cout) but the output is NHWC, so
// each 32-pixel x 128-cout chunk is transposed through a ring of four 8 KB shared-memory buffers:
// the residual chunk arrives there by TMA (prefetched one chunk ahead), every thread fuses
// BN affine + residual + ReLU in place (2-byte accesses, 64 contiguous bytes per warp: conflict-free),
// and one TMA store per chunk writes full 256-byte pixel rows (out-of-image pixels are clipped by TMA).
//
// ROW variant of MODE_T (3x3, stride 1, 64-channel K-blocks, rows of >= 256 output pixels).  Measured on B200: the
// plain K-block pipeline is bound by L2 -> SM bandwidth (48 KB per K-block against ~45 B/clk/SM, i.e. ~1000 cycles
// for 512 cycles of MMA).  The three taps (ky, kx = 0..2) of a filter row read the SAME input row shifted by
// kx*dil pixels, so the row is loaded once as a halo [256 + 2*dil pixels][64 ch] (SWIZZLE_128B) and each tap's B
// operand is the UMMA descriptor started kx*dil pixel rows further (the UMMA swizzle uses absolute shared-memory
// address bits, conv_halo.cu).  Activation traffic drops from 9 x 32 KB to 3 x 33 KB per 64-channel block; rows and
// weight tiles travel through two independent rings (X: 33 KB slots, W: 16 KB slots).
//
// PIX flavour of the ROW variant (tiles with FEW live K-blocks: layers 4-5, 6.0.conv1 at 75 % block sparsity).  The
// staged epilogue above costs ~7000 cycles per 256-pixel x 128-cout tile whatever the K length (TMEM lane = cout
// forces a transpose through shared memory, a residual TMA load and a TMA store per 32-pixel chunk, five mbarrier
// hand-offs per chunk), i.e. as much as 14 K-blocks of MMA.  With the operands swapped — A = the input-row window
// (M = 128 pixels, two M-blocks per 256-pixel row), B = the 128 x 64 weight tile (N = 128 couts) — the SAME shared
// memory images feed the tensor pipe at the same rate, but the accumulator arrives pixel-major (TMEM lane = pixel,
// column = cout): a thread owns one pixel and finishes 16 consecutive couts at a time straight from registers
// (BN affine from shared memory, residual by one 32-byte global load, one 32-byte global store): no staging ring,
// no store warp, no transposition.
#include "conv_internal.cuh"
#include <algorithm>
#include <cstdlib>
#include <cudaTypedefs.h>

namespace drnb200 {

__host__ __device__ constexpr int tc_threads(int ng) { return (3 + 4 * ng) * 32; }   // TMA, MMA, NG x 4 epilogue warps, epilogue-TMA warp (MODE_T); MODE_P uses warps 0-5
constexpr int kTcThreads = tc_threads(2);
constexpr int kMaxStages = 8;
constexpr int MODE_T = 0;      // staged epilogue (16-bit output)
constexpr int MODE_P = 1;
constexpr int MODE_TD = 2;     // MODE_T with the direct (register -> global) epilogue: float32 output
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxEpRing = 8;  // staging buffers of the MODE_T epilogue: 2 per epilogue group
constexpr int kEpChunkPx = 32; // pixels per staged chunk (one tcgen05.ld.32x32b.x32 per warp)
constexpr int kEpBufBytes = kEpChunkPx * 256;
constexpr int kRowPx = 256;                       // ROW variant: output pixels per tile (one row segment)
constexpr int kRowHaloPx = kRowPx + 8;            // + 2*dil (dil <= 4) halo pixels
constexpr int kRowBytes = kRowHaloPx * 128;       // one X-ring slot (33 x 1024 B)
constexpr int kRowWBytes = 128 * 128;             // one W-ring slot: 128 couts x 64 ch x 2 B
constexpr int kMaxXRing = 4;
constexpr double kPixMaxLive = -1.0;              // auto never picks the PIX flavour: measured slower on every layer (see below)

struct __align__(16) TcSync {
  alignas(16) float scale[512];   // MODE_P / PIX: BN affine of all (<= 512) output channels, read by every pixel thread
  alignas(16) float shift[512];   //         (no L1 is left once the CTA takes ~all shared memory)
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t xfull[kMaxXRing];  // ROW variant: input-row ring (full/empty above are then the weight-tile ring)
  uint64_t xempty[kMaxXRing];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t rfull[kMaxEpRing];    // staged epilogue: residual chunk landed in ring slot (TMA)
  uint64_t sfree[kMaxEpRing];    //                  ring slot may be overwritten (no-residual layers)
  uint64_t sdone[kMaxEpRing];    //                  the owning epilogue group finished the chunk
  uint32_t tmem_base;
  uint32_t pad;
};

struct TileCoord {
  int ot, jb, je, n, ox0, oy0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t) {
  TileCoord c;
  const int oi = t / p.n_pix_tiles;
  int pt = t - oi * p.n_pix_tiles;
  c.ot = __ldg(p.ot_order + oi);
  c.jb = __ldg(p.row_ptr + c.ot);
  c.je = __ldg(p.row_ptr + c.ot + 1);
  const int txi = pt % p.tiles_x;
  pt /= p.tiles_x;
  const int tyi = pt % p.tiles_y;
  c.n = pt / p.tiles_y;
  c.ox0 = txi * p.TW;
  c.oy0 = tyi * p.TH;
  return c;
}

template <int DT>
__device__ __forceinline__ float finish(float acc, float sc, float sh, float res, int relu) {
  float v = fmaf(acc, sc, sh) + res;
  return relu ? fmaxf(v, 0.f) : v;
}

// NG = epilogue groups of four warps (MODE_T): 2 when the MMA loop dominates a tile, 4 for tiles with few live K-blocks
// (fused downsample rows, 1x1 projections, heavily pruned layers), where finishing a 128 x 256 accumulator tile
// (~7000 cycles with two groups: dependent 2-byte shared-memory accesses, ~6 cycles per instruction) is the bound.
template <int MODE, int DT, bool ROW = false, int NG = 2, bool PIX = false>
__global__ void __launch_bounds__(tc_threads(NG), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_y,
               const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_x2,
               const ConvParams p) {
  static_assert(!ROW || MODE == MODE_T, "the ROW mainloop feeds the staged MODE_T epilogue");
  static_assert(!PIX || (ROW && NG == 2), "the pixel-major accumulator exists for the ROW mainloop with two epilogue groups");
  static_assert(NG == 2 || (MODE == MODE_T && NG == 4), "four epilogue groups exist for the staged epilogue only");
  constexpr int kEpRing = 2 * NG;             // = 2 * D below: D chunks being finished + D draining / being prepared
  extern __shared__ uint8_t smem_raw[];
  // tiles must sit on 1024-byte boundaries of the shared window (SWIZZLE_128B atom)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stg = smem + (ROW ? (size_t)p.main_bytes : (size_t)p.stages * p.stage_bytes);   // MODE_T staging ring
  TcSync* sync = reinterpret_cast<TcSync*>(stg + (MODE == MODE_T ? kEpRing * kEpBufBytes : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  griddep_launch();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    if (MODE == MODE_T) {
      tma_prefetch_desc(&tmap_y);
      if (p.has_res || p.proj) tma_prefetch_desc(&tmap_r);
    }
    for (int b = 0; b < kEpRing; ++b) {
      mbar_init(&sync->rfull[b], 1);
      mbar_init(&sync->sfree[b], 1);
      mbar_init(&sync->sdone[b], 4);
    }
    for (int s = 0; s < (ROW ? p.w_ring : p.stages); ++s) {
      mbar_init(&sync->full[s], 1);
      mbar_init(&sync->empty[s], 1);
    }
    if (ROW) {
      tma_prefetch_desc(&tmap_x2);
      for (int s = 0; s < p.x_ring; ++s) {
        mbar_init(&sync->xfull[s], 1);
        mbar_init(&sync->xempty[s], 1);
      }
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sync->tfull[a], 1);
      mbar_init(&sync->tempty[a], MODE == MODE_T ? 4 * NG : 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&sync->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  griddep_wait();       // up to here the CTA overlapped the previous kernel's tail; no global memory was read yet
  if (MODE == MODE_P || PIX) {
    for (int ch = threadIdx.x; ch < 512; ch += tc_threads(NG)) {
      sync->scale[ch] = ch < p.Cout ? __ldg(p.scale + ch) : 0.f;
      sync->shift[ch] = ch < p.Cout ? __ldg(p.shift + ch) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sync->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    // (the whole warp runs the loop with warp-uniform values, one elected lane issues: inside an
    //  `if (lane == 0)` region nothing is provably uniform and every operand goes through R2UR)
    if (ROW) {
      // entries of the tile list are sorted by kb = cib*9 + ky*3 + kx: a run with equal kb/3 = one input row
      uint32_t xs = 0, xph = 0, ws = 0, wph = 0;
      uint8_t* wring = smem + (size_t)p.x_ring * kRowBytes;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        int j = c.jb;
        int kb = (j < c.je) ? __ldg(p.kblk + j) : 0;
        while (j < c.je) {
          const int g = kb / 3, cib = g / 3, ky = g - cib * 3;
          uint8_t* sX = smem + (size_t)xs * kRowBytes;
          mbar_wait(&sync->xempty[xs], xph ^ 1u);
          if (p.dbg & 2) {
            if (elect_one()) mbar_arrive(&sync->xfull[xs]);
          } else if (kb >= DRNB200_KB_PROJ) {
            // residual projection (1x1, stride 1) as a K-block: 64 channels of the block's INPUT at the 256 output
            // pixels themselves, no halo; the MMA warp sees a tap with kx = 0 (DRNB200_KB_PROJ is a multiple of 3)
            if (elect_one()) {
              mbar_arrive_expect_tx(&sync->xfull[xs], kRowPx * 128);
              tma_load_4d(&tmap_r, &sync->xfull[xs], sX, p.res_coff + (kb - DRNB200_KB_PROJ) / 3 * 64, c.ox0, c.oy0, c.n);
            }
          } else if (elect_one()) {
            mbar_arrive_expect_tx(&sync->xfull[xs], kRowBytes);
            const int x0 = c.ox0 - p.dil, y = c.oy0 + (ky - 1) * p.dil;
            tma_load_4d(&tmap, &sync->xfull[xs], sX, cib * 64, x0, y, c.n);                        // 256 pixels
            tma_load_4d(&tmap_x2, &sync->xfull[xs], sX + kRowPx * 128, cib * 64, x0 + kRowPx, y, c.n);   // + 8
          }
          __syncwarp();
          if (++xs == (uint32_t)p.x_ring) { xs = 0; xph ^= 1u; }
          do {
            mbar_wait(&sync->empty[ws], wph ^ 1u);
            if (p.dbg & 1) {
              if (elect_one()) mbar_arrive(&sync->full[ws]);
            } else if (elect_one()) {
              mbar_arrive_expect_tx(&sync->full[ws], kRowWBytes);
              bulk_load(p.w_packed + (size_t)j * kRowWBytes, &sync->full[ws], wring + (size_t)ws * kRowWBytes,
                        kRowWBytes);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.w_ring) { ws = 0; wph ^= 1u; }
            ++j;
            if (j < c.je) kb = __ldg(p.kblk + j);
          } while (j < c.je && kb / 3 == g);
        }
      }
    } else {
      uint32_t stage = 0, phase = 0;
      const int half = (p.taps == 9) ? 1 : 0;
      const uint32_t tx_bytes = p.w_tile_bytes + p.x_tile_bytes;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        int kb_next = (c.jb < c.je) ? __ldg(p.kblk + c.jb) : 0;
        for (int j = c.jb; j < c.je; ++j) {
          const int kb = kb_next;
          if (j + 1 < c.je) kb_next = __ldg(p.kblk + j + 1);
          const int cib = kb / p.taps, tap = kb - cib * p.taps;
          const int ky = (p.taps == 9) ? tap / 3 : 0, kx = (p.taps == 9) ? tap - ky * 3 : 0;
          uint8_t* sW = smem + (size_t)stage * p.stage_bytes;
          uint8_t* sX = sW + p.w_stage_bytes;
          mbar_wait(&sync->empty[stage], phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&sync->full[stage], tx_bytes);
            bulk_load(p.w_packed + (size_t)j * p.w_tile_bytes, &sync->full[stage], sW, p.w_tile_bytes);
            tma_load_4d(&tmap, &sync->full[stage], sX, cib * p.tile_ci,
                        c.ox0 * p.stride + (kx - half) * p.dil, c.oy0 * p.stride + (ky - half) * p.dil,
                        c.n);
          }
          __syncwarp();
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform loop)
    if (ROW) {
      uint32_t xs = 0, xph = 0, ws = 0, wph = 0, acc = 0, acc_phase = 0;
      const uint64_t d_hi = umma_smem_desc(0u, 128);
      const uint32_t x16 = smem_u32(smem) >> 4, w16 = x16 + (uint32_t)p.x_ring * (kRowBytes >> 4);
      const uint32_t shift16 = (uint32_t)p.dil * 8u;          // kx*dil pixel rows of 128 B, in 16-byte units
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        if (c.je == c.jb) continue;
        mbar_wait(&sync->tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        int j = c.jb;
        int kb = __ldg(p.kblk + j);
        while (j < c.je) {
          const int g = kb / 3;
          mbar_wait(&sync->xfull[xs], xph);
          const uint32_t sX16 = x16 + xs * (uint32_t)(kRowBytes >> 4);
          do {
            const uint32_t kx = (uint32_t)(kb - g * 3);
            mbar_wait(&sync->full[ws], wph);
            tc_fence_after();
            const uint32_t sW16 = w16 + ws * (uint32_t)(kRowWBytes >> 4);
            if (elect_one()) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint64_t dW = d_hi | (uint64_t)(sW16 + 2u * i), dX = d_hi | (uint64_t)(sX16 + kx * shift16 + 2u * i);
                const uint32_t accum = (j > c.jb || i > 0) ? 1u : 0u;
                if (!PIX) {
                  umma_f16(d_tmem, dW, dX, p.idesc, accum);
                } else {        // pixels as M: rows 0-127 and 128-255 of the window, 128 columns (couts) each
                  umma_f16(d_tmem, dX, dW, p.idesc, accum);
                  umma_f16(d_tmem + 128u, dX + (uint64_t)(128u * 128u >> 4), dW, p.idesc, accum);
                }
              }
              umma_commit(&sync->empty[ws]);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.w_ring) { ws = 0; wph ^= 1u; }
            ++j;
            if (j < c.je) kb = __ldg(p.kblk + j);
          } while (j < c.je && kb / 3 == g);
          if (elect_one()) umma_commit(&sync->xempty[xs]);      // the row's taps have retired
          __syncwarp();
          if (++xs == (uint32_t)p.x_ring) { xs = 0; xph ^= 1u; }
        }
        if (elect_one()) umma_commit(&sync->tfull[acc]);
        __syncwarp();
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    } else {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const int ksteps = (int)(p.pitch / 32u);  // UMMA K = 16 elements = 32 bytes
      const uint64_t d_hi = umma_smem_desc(0u, p.pitch);
      const uint32_t wstage16 = p.w_stage_bytes >> 4;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        if (c.je == c.jb) continue;  // nothing live: the epilogue does not touch TMEM either
        mbar_wait(&sync->tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        for (int j = c.jb; j < c.je; ++j) {
          mbar_wait(&sync->full[stage], phase);
          tc_fence_after();
          const uint32_t sW16 = smem_u32(smem + (size_t)stage * p.stage_bytes) >> 4;
          const uint32_t sX16 = sW16 + wstage16;
          if (elect_one()) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < ksteps) {
                const uint32_t accum = (j > c.jb || i > 0) ? 1u : 0u;
                const uint64_t dW = d_hi | (uint64_t)(sW16 + 2u * i), dX = d_hi | (uint64_t)(sX16 + 2u * i);
                if (MODE != MODE_P) umma_f16(d_tmem, dW, dX, p.idesc, accum);
                else umma_f16(d_tmem, dX, dW, p.idesc, accum);
              }
            }
            umma_commit(&sync->empty[stage]);  // smem slot free once these MMAs retire
          }
          __syncwarp();
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(&sync->tfull[acc]);      // accumulator complete
        __syncwarp();
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t acc = 0, acc_phase = 0;
    const uint16_t* res16 = reinterpret_cast<const uint16_t*>(p.residual);
    uint16_t* y16 = reinterpret_cast<uint16_t*>(p.y);
    float* y32 = reinterpret_cast<float*>(p.y);

    if (PIX) {
      // ------------------------------------------------ pixel-major accumulator: registers -> global, no staging
      // warps 2..9: group (warp-2)>>2 owns the M-block of pixels [128*grp, 128*grp+128) of the row tile, the warp its
      // TMEM lane quarter; thread = one output pixel.  32 couts per step: the TMEM load and the two 32-byte residual
      // loads of step s+1 are in flight while step s is finished (BN affine from shared memory as float4 broadcasts)
      // and written with two 32-byte global stores.
      if (warp < 2 + 4 * NG) {
        const int grp = (warp - 2) >> 2;
        const int px = grp * 128 + q * 32 + lane;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
          const TileCoord c = decode_tile(p, t);
          const bool live = c.je > c.jb;
          const int ox = c.ox0 + px;
          const bool valid = ox < p.OW;                               // TH == 1: the row itself is always inside
          const bool use_res = p.has_res && valid;
          const size_t pix = ((size_t)c.n * p.OH + c.oy0) * p.OW + ox;
          const uint16_t* rrow = res16 + pix * p.res_pitch + p.res_coff + c.ot * 128;
          uint16_t* yrow = y16 + pix * p.Cout + c.ot * 128;
          const float* sc = sync->scale + c.ot * 128;
          const float* sh = sync->shift + c.ot * 128;
          const float floor_v = (c.ot * 128 < p.relu_n) ? 0.f : -INFINITY;   // relu_n is a multiple of 128 (plan check)
          uint32_t rw[2][16], v[2][32];
          if (use_res) { ldg256_nc(rrow, *reinterpret_cast<uint32_t(*)[8]>(&rw[0][0])); ldg256_nc(rrow + 16, *reinterpret_cast<uint32_t(*)[8]>(&rw[0][8])); }
          if (live) {
            mbar_wait(&sync->tfull[acc], acc_phase);
            tc_fence_after();
          }
          const uint32_t t_addr = tmem_base + acc * 256u + (uint32_t)(grp * 128) + ((uint32_t)(q * 32) << 16);
          if (live) tmem_ld32(t_addr, v[0]);
#pragma unroll
          for (int st = 0; st < 4; ++st) {
            const int cb = st * 32;
            if (live) tmem_ld_wait();
            if (st < 3) {
              if (live) tmem_ld32(t_addr + (uint32_t)(cb + 32), v[(st + 1) & 1]);
              if (use_res) {
                ldg256_nc(rrow + cb + 32, *reinterpret_cast<uint32_t(*)[8]>(&rw[(st + 1) & 1][0]));
                ldg256_nc(rrow + cb + 48, *reinterpret_cast<uint32_t(*)[8]>(&rw[(st + 1) & 1][8]));
              }
            } else if (live) {             // the accumulator has been read out: hand it back before finishing the step
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&sync->tempty[acc]);
            }
            uint32_t w[16];
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 s4 = *reinterpret_cast<const float4*>(sc + cb + 4 * i4);
              const float4 h4 = *reinterpret_cast<const float4*>(sh + cb + 4 * i4);
              const uint32_t r0 = rw[st & 1][2 * i4], r1 = rw[st & 1][2 * i4 + 1];
              float a0 = live ? __uint_as_float(v[st & 1][4 * i4 + 0]) : 0.f, a1 = live ? __uint_as_float(v[st & 1][4 * i4 + 1]) : 0.f;
              float a2 = live ? __uint_as_float(v[st & 1][4 * i4 + 2]) : 0.f, a3 = live ? __uint_as_float(v[st & 1][4 * i4 + 3]) : 0.f;
              a0 = fmaf(a0, s4.x, h4.x); a1 = fmaf(a1, s4.y, h4.y); a2 = fmaf(a2, s4.z, h4.z); a3 = fmaf(a3, s4.w, h4.w);
              if (p.has_res) {
                a0 += use_res ? Act<DT>::to_f32((uint16_t)(r0 & 0xFFFFu)) : 0.f;
                a1 += use_res ? Act<DT>::to_f32((uint16_t)(r0 >> 16)) : 0.f;
                a2 += use_res ? Act<DT>::to_f32((uint16_t)(r1 & 0xFFFFu)) : 0.f;
                a3 += use_res ? Act<DT>::to_f32((uint16_t)(r1 >> 16)) : 0.f;
              }
              w[2 * i4] = pack2<DT>(fmaxf(a0, floor_v), fmaxf(a1, floor_v));
              w[2 * i4 + 1] = pack2<DT>(fmaxf(a2, floor_v), fmaxf(a3, floor_v));
            }
            if (valid) {
              stg256(yrow + cb, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
              stg256(yrow + cb + 16, *reinterpret_cast<uint32_t(*)[8]>(&w[8]));
            }
          }
          if (live) {
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
          }
        }
      }
    } else if (MODE == MODE_T) {
      // ------------------------------------------------ staged: smem transpose + TMA load/store
      // warps 2 .. 2+4*NG-1 = NG groups of four warps taking 32-pixel chunks round-robin (flat chunk index k: group
      // k % NG, ring slot k % (2*NG)); the last warp issues every residual load and output store and recycles the slots.
      if (warp == 2 + 4 * NG) {
        // ---- epilogue TMA warp (warp-uniform loop; one elected lane issues and owns the bulk groups)
        const int nch = p.ep_nch;
        auto chunk_xy = [&](const TileCoord& c, int qq, int& cx, int& cy) {
          const int j0 = qq * kEpChunkPx;
          cx = c.ox0 + (j0 & (p.TW - 1));
          cy = c.oy0 + (j0 >> p.tw_shift);
        };
        constexpr int D = NG;                      // chunks in flight (prepared, not yet stored): one per group
        int hist_ot[D], hist_cx[D], hist_cy[D], hist_n[D];
        uint32_t k = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
          const TileCoord c = decode_tile(p, t);
          for (int qq = 0; qq < nch; ++qq, ++k) {
            const uint32_t b = k & (kEpRing - 1);
            int cx, cy;
            chunk_xy(c, qq, cx, cy);
            // (A) prepare chunk k FIRST, so that the group that is finishing chunk k - D finds its next slot ready:
            // ring slot b was last used by chunk k - kEpRing = k - 2D; the stores issued so far are those of chunks
            // <= k - D - 1, D - 1 of them newer than that one, so it has drained once at most D - 1 are still reading
            if (elect_one()) {
              bulk_wait_group_read<D - 1>();
              if (p.has_res) {                      // residual chunk [32 px][128 couts] straight into the slot: two
                mbar_arrive_expect_tx(&sync->rfull[b], kEpBufBytes);        // SWIZZLE_128B boxes of 64 couts
                tma_load_4d(&tmap_r, &sync->rfull[b], stg + b * kEpBufBytes, p.res_coff + c.ot * 128, cx, cy, c.n);
                tma_load_4d(&tmap_r, &sync->rfull[b], stg + b * kEpBufBytes + kEpBufBytes / 2,
                            p.res_coff + c.ot * 128 + 64, cx, cy, c.n);
              } else {
                mbar_arrive(&sync->sfree[b]);
              }
            }
            __syncwarp();
            // (B) finish chunk k - D: its group is done -> store it
            if (k >= (uint32_t)D) {
              const uint32_t kb = k - D, bb = kb & (kEpRing - 1);
              mbar_wait(&sync->sdone[bb], (kb / kEpRing) & 1u);
              if (elect_one()) {
                if (!(p.dbg & 8)) {
                  tma_store_4d(&tmap_y, stg + bb * kEpBufBytes, hist_ot[kb % D] * 128, hist_cx[kb % D],
                               hist_cy[kb % D], hist_n[kb % D]);
                  tma_store_4d(&tmap_y, stg + bb * kEpBufBytes + kEpBufBytes / 2, hist_ot[kb % D] * 128 + 64,
                               hist_cx[kb % D], hist_cy[kb % D], hist_n[kb % D]);
                }
                bulk_commit_group();
              }
              __syncwarp();
            }
            hist_ot[k % D] = c.ot; hist_cx[k % D] = cx; hist_cy[k % D] = cy; hist_n[k % D] = c.n;
          }
        }
        // drain: the last min(k, D) chunks still have to be stored
        for (uint32_t kb = (k >= (uint32_t)D ? k - D : 0u); kb < k; ++kb) {
          const uint32_t bb = kb & (kEpRing - 1);
          mbar_wait(&sync->sdone[bb], (kb / kEpRing) & 1u);
          if (elect_one()) {
            tma_store_4d(&tmap_y, stg + bb * kEpBufBytes, hist_ot[kb % D] * 128, hist_cx[kb % D], hist_cy[kb % D],
                         hist_n[kb % D]);
            tma_store_4d(&tmap_y, stg + bb * kEpBufBytes + kEpBufBytes / 2, hist_ot[kb % D] * 128 + 64, hist_cx[kb % D],
                         hist_cy[kb % D], hist_n[kb % D]);
            bulk_commit_group();
          }
          __syncwarp();
        }
        if (elect_one()) bulk_wait_group<0>();     // all stores complete before the CTA retires
        __syncwarp();
      } else {
        // ---- epilogue groups.  The accumulator is read with tcgen05.ld.16x256b (mma C-fragment layout: thread t holds
        // couts t/4 + {0, 8} (+16 with the second load) x pixel pairs), so after BN / residual / ReLU and packing to 16
        // bits the registers ARE stmatrix fragments: stmatrix.trans writes 16-byte rows of 8 couts per pixel, the
        // residual arrives the same way through ldmatrix.trans.  One chunk costs a warp 4 + 4 shared-memory
        // instructions of 512 bytes each instead of 32 + 32 two-byte accesses (which occupied the 128 B/clk pipe for 64
        // bytes each: the shared-memory pipe is what bounds this kernel, profiles/r02_conv_row_accumulator_layouts.txt).
        // Slot layout: two SWIZZLE_128B boxes [32 px][64 couts]; a pixel's 16-byte chunk c sits at c ^ (pixel & 7).
        const int grp = (warp - 2) >> 2;
        const int g = lane >> 2;                   // cout row of this thread inside an 8-cout group
        const int nch = p.ep_nch;
        const uint32_t half_off = (uint32_t)(q >> 1) * (kEpBufBytes / 2);       // couts 0-63 | 64-127 of the tile
        const uint32_t cbase = (uint32_t)(q & 1) * 4u;                         // first 16-byte chunk of this warp's 32 couts
        const uint32_t row_off = (uint32_t)(8 * (lane >> 3) + (lane & 7)) * 128u;   // matrix m = lane/8 -> pixels 8m..8m+7
        uint32_t k0 = 0;                           // flat chunk index of the tile's first chunk
        // the tile's decode and BN affine are fetched one tile ahead: on tiles with few live K-blocks this global
        // round trip (plus the integer divisions of the decode) would otherwise sit on every warp's critical path
        TileCoord c_next = decode_tile(p, blockIdx.x < p.total_tiles ? blockIdx.x : 0);
        float sc_next[4], sh_next[4];
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) {
          sc_next[cg] = __ldg(p.scale + c_next.ot * 128 + q * 32 + 8 * cg + g);
          sh_next[cg] = __ldg(p.shift + c_next.ot * 128 + q * 32 + 8 * cg + g);
        }
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
          const TileCoord c = c_next;
          float sc[4], sh[4];
          int relu[4];
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            sc[cg] = sc_next[cg]; sh[cg] = sh_next[cg];
            relu[cg] = (c.ot * 128 + q * 32 + 8 * cg + g) < p.relu_n;
          }
          if (t + (int)gridDim.x < p.total_tiles) {
            c_next = decode_tile(p, t + gridDim.x);
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
              sc_next[cg] = __ldg(p.scale + c_next.ot * 128 + q * 32 + 8 * cg + g);
              sh_next[cg] = __ldg(p.shift + c_next.ot * 128 + q * 32 + 8 * cg + g);
            }
          }
          const bool live = c.je > c.jb;
          int qq = (int)(((uint32_t)grp - k0) & (uint32_t)(NG - 1));   // first chunk of this tile owned by this group
          if (live) {
            mbar_wait(&sync->tfull[acc], acc_phase);
            tc_fence_after();
          }
          const uint32_t t_addr = tmem_base + acc * 256u + ((uint32_t)(q * 32) << 16);
          for (; qq < nch; qq += NG) {
            const uint32_t k = k0 + (uint32_t)qq;
            const uint32_t b = k & (kEpRing - 1);
            const uint32_t slot = smem_u32(stg + b * kEpBufBytes) + half_off + row_off;
            uint32_t va[16], vb[16];               // couts g, g+8 | g+16, g+24  x  4 blocks of 8 pixels
            if (live) {
              tmem_ld_16x256b_x4(t_addr + (uint32_t)(qq * kEpChunkPx), va);
              tmem_ld_16x256b_x4(t_addr + (16u << 16) + (uint32_t)(qq * kEpChunkPx), vb);
            }
            // the slot's previous store has drained; with a residual, its chunk [32 px][128 couts] has landed (TMA)
            if (p.has_res) mbar_wait(&sync->rfull[b], (k / kEpRing) & 1u);
            else mbar_wait(&sync->sfree[b], (k / kEpRing) & 1u);
            if (live) {
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) { va[i] = 0u; vb[i] = 0u; }
            }
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
              const uint32_t addr = slot + (((cbase + (uint32_t)cg) ^ (uint32_t)(lane & 7)) << 4);
              uint32_t r[4] = {0u, 0u, 0u, 0u}, o[4];
              if (p.has_res) ldmatrix_x4_trans(addr, r[0], r[1], r[2], r[3]);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {     // pixel block kk: pixels 8kk + 2(lane%4), +1
                const uint32_t* src = cg < 2 ? va : vb;
                const int idx = 4 * kk + 2 * (cg & 1);
                const float r_lo = p.has_res ? Act<DT>::to_f32((uint16_t)(r[kk] & 0xFFFFu)) : 0.f;
                const float r_hi = p.has_res ? Act<DT>::to_f32((uint16_t)(r[kk] >> 16)) : 0.f;
                o[kk] = pack2<DT>(finish<DT>(__uint_as_float(src[idx]), sc[cg], sh[cg], r_lo, relu[cg]),
                                  finish<DT>(__uint_as_float(src[idx + 1]), sc[cg], sh[cg], r_hi, relu[cg]));
              }
              stmatrix_x4_trans(addr, o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async_smem();              // st.shared -> visible to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(&sync->sdone[b]);
          }
          k0 += (uint32_t)nch;
          if (live) {                              // this warp has read all of its chunks of the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sync->tempty[acc]);
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
          }
        }
      }
    } else if (warp < 6) {
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord c = decode_tile(p, t);
      const bool live = c.je > c.jb;
      if (live) {
        mbar_wait(&sync->tfull[acc], acc_phase);
        tc_fence_after();
      }
      const uint32_t t_addr = tmem_base + acc * 256u + ((uint32_t)(q * 32) << 16);

      if (MODE == MODE_TD) {
        // lane = output channel, columns = pixels of the tile; direct 2/4-byte global accesses
        const int co = c.ot * 128 + q * 32 + lane;
        const float sc = __ldg(p.scale + co), sh = __ldg(p.shift + co);
        const int nt = p.TW * p.TH;
        for (int ch = 0; ch < nt; ch += 32) {
          uint32_t v[32];
          if (live) {
            tmem_ld32(t_addr + (uint32_t)ch, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = ch + i;
            const int oy = c.oy0 + (j >> p.tw_shift), ox = c.ox0 + (j & (p.TW - 1));
            if (oy < p.OH && ox < p.OW) {
              const size_t pix = ((size_t)c.n * p.OH + oy) * p.OW + ox;
              const size_t off = pix * p.Cout + co;
              const float r = p.has_res ? Act<DT>::to_f32(__ldg(res16 + pix * p.res_pitch + p.res_coff + co)) : 0.f;
              const float o = finish<DT>(__uint_as_float(v[i]), sc, sh, r, co < p.relu_n);
              if (p.out_f32) y32[off] = o;
              else y16[off] = Act<DT>::from_f32(o);
            }
          }
        }
      } else {
        // lane = pixel of the tile (M = 128), columns = output channels (N = Cout)
        const int j = q * 32 + lane;
        const int oy = c.oy0 + (j >> p.tw_shift), ox = c.ox0 + (j & (p.TW - 1));
        const bool valid = (oy < p.OH) && (ox < p.OW);
        const size_t pix0 = ((size_t)c.n * p.OH + oy) * p.OW + ox;
        const size_t off0 = pix0 * p.Cout;
        for (int cb = 0; cb < p.Cout; cb += 16) {
          uint32_t v[16];
          if (live) {
            tmem_ld16(t_addr + (uint32_t)cb, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0u;
          }
          if (valid) {
            float r[16];
            if (p.has_res) {
              uint32_t rw[8];
              ldg256_nc(res16 + pix0 * p.res_pitch + p.res_coff + cb, rw);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                r[2 * i] = Act<DT>::to_f32((uint16_t)(rw[i] & 0xFFFFu));
                r[2 * i + 1] = Act<DT>::to_f32((uint16_t)(rw[i] >> 16));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = 0.f;
            }
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              o[i] = finish<DT>(__uint_as_float(v[i]), sync->scale[cb + i], sync->shift[cb + i], r[i],
                                cb + i < p.relu_n);
            if (p.out_f32) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(o[8 * h + i]);
                stg256(y32 + off0 + cb + 8 * h, w);
              }
            } else {
              uint32_t w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) w[i] = pack2<DT>(o[2 * i], o[2 * i + 1]);
              stg256(y16 + off0 + cb, w);
            }
          }
        }
      }

      if (live) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sync->tempty[acc]);
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    }
  }

  // ------------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------- host side

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
  return fn;
}

static int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}
static int ilog2(int v) {
  int r = 0;
  while ((1 << r) < v) ++r;
  return r;
}

template <int MODE, int DT, bool ROW = false, int NG = 2, bool PIX = false>
static int set_attr(size_t smem) {
  DRN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<MODE, DT, ROW, NG, PIX>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return DRNB200_OK;
}

int conv_tc_setup(drnb200_conv_plan* plan) {
  ConvParams& p = plan->p;
  const drnb200_conv_desc& d = plan->d;
  // ---- shape support
  const bool ci_ok = (p.tile_ci == 16 || p.tile_ci == 32 || p.tile_ci == 64);
  const bool k_ok = (d.ksize == 1 || d.ksize == 3) && (d.stride == 1 || d.stride == 2);
  int mode = -1;
  if (ci_ok && k_ok) {
    if (p.tile_o == 128 && p.Cout % 128 == 0) mode = d.out_f32 ? MODE_TD : MODE_T;
    else if (p.tile_o == p.Cout && p.Cout % 16 == 0 && p.Cout <= 256) mode = MODE_P;
  }
  if (mode < 0) {
    set_error("tcgen05 conv: unsupported shape Cin=%d Cout=%d k=%d s=%d tile=%dx%d", d.Cin, d.Cout,
              d.ksize, d.stride, p.tile_o, p.tile_ci);
    return DRNB200_E_ARG;
  }
  plan->tc_mode = mode;
  // ---- pixel tile
  const int nt_max = (mode != MODE_P) ? 256 : 128;
  int TW = std::min(std::min(pow2_ceil(p.OW), nt_max), 256 / d.stride);
  int TH = std::min(pow2_ceil(p.OH), nt_max / TW);
  if (mode == MODE_P) TH = 128 / TW;           // UMMA M is exactly 128 pixels
  while (TW * TH < 32) TH <<= 1;               // epilogue walks 32-column chunks
  if (TH * d.stride > 256) {
    set_error("tcgen05 conv: pixel tile %dx%d does not fit a TMA box", TW, TH);
    return DRNB200_E_ARG;
  }
  p.TW = TW; p.TH = TH; p.tw_shift = ilog2(TW);
  p.tiles_x = (p.OW + TW - 1) / TW;
  p.tiles_y = (p.OH + TH - 1) / TH;
  p.n_pix_tiles = p.N * p.tiles_x * p.tiles_y;
  p.total_tiles = p.n_pix_tiles * p.n_ot;
  const int NT = TW * TH;
  p.pitch = (uint32_t)p.tile_ci * 2u;
  p.w_tile_bytes = (uint32_t)p.tile_o * p.pitch;
  p.x_tile_bytes = (uint32_t)NT * p.pitch;
  p.w_stage_bytes = (p.w_tile_bytes + 1023u) & ~1023u;
  p.stage_bytes = p.w_stage_bytes + ((p.x_tile_bytes + 1023u) & ~1023u);
  p.idesc = (mode != MODE_P) ? umma_idesc_f16(128, NT, d.act_dtype)
                             : umma_idesc_f16(128, p.Cout, d.act_dtype);
  // staged epilogue geometry: chunks of 32 consecutive tile pixels = a box of ep_cw x ep_ch pixels
  p.ep_cw = std::min(TW, kEpChunkPx);
  p.ep_ch = kEpChunkPx / p.ep_cw;
  p.ep_nch = NT / kEpChunkPx;
  // ---- shared memory: as many stages as fit in 227 KB (also pins one CTA per SM: TMEM is taken whole)
  const size_t kMaxSmem = 232448;
  // ---- epilogue groups (MODE_T): two by default.  A four-group instantiation exists (DRNB200_NG=4) and was measured:
  //      once the store warp prepares a group's next slot BEFORE waiting for the chunk in flight, two groups finish
  //      tiles as fast as four (layer-5/6 residual convs 0.129/0.249 -> 0.115/0.228 ms with either), and tiles with a
  //      single live K-block stay at ~4700 cycles with no math and no stores at all (pipeline skeleton).
  p.ep_groups = 2;
  static const char* env_row = getenv("DRNB200_ROW");       // A/B knob: "0" keeps the per-tap pipeline
  // ROW variant: 3x3 stride-1 convs over 64-channel K-blocks whose tile is one 256-pixel row segment
  const bool row_ok = mode == MODE_T && p.taps == 9 && d.stride == 1 && p.tile_ci == 64 && d.dilation <= 4 &&
                      TW == kRowPx && TH == 1 && !(env_row && env_row[0] == '0');
  if (mode == MODE_T) {
    // 1x1 convs with a residual and very few live K-blocks per tile (the 512 -> 2048 expand convs of the Bottleneck
    // networks at 75 % sparsity: 2 of 8) are pure streaming work: four groups = eight staging slots keep twice as many
    // residual loads / output stores in flight (DRN-D-54 conv3 launches: 2.675 -> 2.578 ms per step, measured); every
    // other layer loses main-ring stages to the extra staging and is slower with four
    const double avg_live = (double)plan->h_row_ptr[p.n_ot] / std::max(1, p.n_ot);
    if (p.taps == 1 && p.has_res && avg_live <= 4.0) p.ep_groups = 4;
    // row-halo 3x3 convs that read a residual and keep <= 12 K-blocks per tile (layers 4 and 5 at 75 %) wait for their
    // residual chunks, not for MMAs: eight slots in flight win 4-6 % there (4.0.conv2 0.0505 -> 0.0476 ms, 5.1.conv2
    // 0.0878 -> 0.0831, same-box A/B); with 18 live K-blocks (6.1.conv2) or a projection instead of a residual the
    // shallower row ring costs 4-14 %
    if (row_ok && p.has_res && !p.proj && avg_live <= 12.0 && d.acc_layout != 2) p.ep_groups = 4;
    static const char* env_ng = getenv("DRNB200_NG");        // A/B knob: force 2 or 4 epilogue groups
    if (env_ng && (env_ng[0] == '2' || env_ng[0] == '4')) p.ep_groups = env_ng[0] - '0';
  }
  const size_t fixed = 1024 + sizeof(TcSync) + (mode == MODE_T ? 2 * p.ep_groups * kEpBufBytes : 0);
  int stages = (int)((kMaxSmem - fixed) / p.stage_bytes);
  stages = std::max(2, std::min(stages, kMaxStages));
  p.stages = stages;
  plan->smem_bytes = kMaxSmem;
  p.row_mode = 0;
  static const int env_dbg = diag_env("DRNB200_DBG");        // -DDRNB200_DIAG builds only: timing diagnostics of the ROW
  p.dbg = env_dbg;                                           // mainloop (results invalid): 1 no weight loads, 2 no row loads,
                                                             // 4 no residual loads, 8 no output stores, 16 no epilogue math
  if (row_ok) {
    static const char* env_ring = getenv("DRNB200_ROW_RING");   // "x,w" ring sizes for tuning
    int xr = p.ep_groups == 4 ? 2 : 3, wr = 5;            // four groups: 64 KB of staging, shallower row ring
    if (env_ring && sscanf(env_ring, "%d,%d", &xr, &wr) != 2) { xr = 3; wr = 5; }
    xr = std::max(2, std::min(xr, kMaxXRing));
    wr = std::max(2, std::min(wr, kMaxStages));
    while ((size_t)xr * kRowBytes + (size_t)wr * kRowWBytes + fixed > kMaxSmem && wr > 2) --wr;
    p.row_mode = 1; p.x_ring = xr; p.w_ring = wr;
    p.main_bytes = (uint32_t)(xr * kRowBytes + wr * kRowWBytes);
    // pixel-major accumulator for tiles with few live K-blocks (the staged epilogue costs ~14 K-blocks of MMA time
    // per tile; measured cross-over below).  DRNB200_PIX=0/1 forces it off/on for A/B runs.
    static const char* env_pix = getenv("DRNB200_PIX");
    const double avg_live = (double)plan->h_row_ptr[p.n_ot] / std::max(1, p.n_ot);
    const bool pix_ok = p.ep_groups == 2 && p.Cout <= 512 && p.relu_n % 128 == 0;
    const int want = d.acc_layout == 2 || (env_pix && env_pix[0] == '1') ? 1
                     : d.acc_layout == 1 || (env_pix && env_pix[0] == '0') ? 0 : (avg_live <= kPixMaxLive);
    if (pix_ok && want) {
      p.pix_mode = 1;
      p.idesc = umma_idesc_f16(128, 128, d.act_dtype);
    }
  }
  int rc;
  const bool bf16 = d.act_dtype == DRNB200_BF16;
  if (mode == MODE_T && p.row_mode && p.pix_mode)
    rc = bf16 ? set_attr<MODE_T, DRNB200_BF16, true, 2, true>(plan->smem_bytes) : set_attr<MODE_T, DRNB200_F16, true, 2, true>(plan->smem_bytes);
  else if (mode == MODE_T && p.row_mode && p.ep_groups == 4)
    rc = bf16 ? set_attr<MODE_T, DRNB200_BF16, true, 4>(plan->smem_bytes) : set_attr<MODE_T, DRNB200_F16, true, 4>(plan->smem_bytes);
  else if (mode == MODE_T && p.row_mode)
    rc = bf16 ? set_attr<MODE_T, DRNB200_BF16, true>(plan->smem_bytes) : set_attr<MODE_T, DRNB200_F16, true>(plan->smem_bytes);
  else if (mode == MODE_T && p.ep_groups == 4)
    rc = bf16 ? set_attr<MODE_T, DRNB200_BF16, false, 4>(plan->smem_bytes) : set_attr<MODE_T, DRNB200_F16, false, 4>(plan->smem_bytes);
  else if (mode == MODE_T)
    rc = (d.act_dtype == DRNB200_BF16) ? set_attr<MODE_T, DRNB200_BF16>(plan->smem_bytes)
                                       : set_attr<MODE_T, DRNB200_F16>(plan->smem_bytes);
  else if (mode == MODE_TD)
    rc = (d.act_dtype == DRNB200_BF16) ? set_attr<MODE_TD, DRNB200_BF16>(plan->smem_bytes)
                                       : set_attr<MODE_TD, DRNB200_F16>(plan->smem_bytes);
  else
    rc = (d.act_dtype == DRNB200_BF16) ? set_attr<MODE_P, DRNB200_BF16>(plan->smem_bytes)
                                       : set_attr<MODE_P, DRNB200_F16>(plan->smem_bytes);
  if (rc) return rc;
  // ---- work list: output tiles by decreasing live count, so the round-robin persistent schedule
  //      gives every CTA tiles of equal length at the same step (block-sparse load balance)
  std::vector<int32_t> order(p.n_ot);
  for (int i = 0; i < p.n_ot; ++i) order[i] = i;
  const std::vector<int32_t>& rp = plan->h_row_ptr;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return (rp[a + 1] - rp[a]) > (rp[b + 1] - rp[b]);
  });
  DRN_CUDA(cudaMalloc((void**)&plan->d_ot_order, sizeof(int32_t) * p.n_ot));
  DRN_CUDA(cudaMemcpy(plan->d_ot_order, order.data(), sizeof(int32_t) * p.n_ot,
                      cudaMemcpyHostToDevice));
  p.ot_order = plan->d_ot_order;
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  plan->grid = std::min(p.total_tiles, sms);
  plan->tmap_ptr = nullptr;
  plan->tmap_y_ptr = nullptr;
  plan->tmap_r_ptr = nullptr;
  return DRNB200_OK;
}

// output-shaped NHWC tensor, box = HALF a staged chunk (64 couts x ep_cw x ep_ch pixels), SWIZZLE_128B (two per chunk);
// `cpitch` = channels per pixel of the tensor the pointer lives in (the residual may be a channel sub-range)
static int encode_out_tmap(drnb200_conv_plan* plan, CUtensorMap* map, const void* ptr, int cpitch) {
  const ConvParams& p = plan->p;
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return DRNB200_E_CUDA;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)cpitch, (cuuint64_t)p.OW, (cuuint64_t)p.OH, (cuuint64_t)p.N};
  cuuint64_t gstr[3] = {(cuuint64_t)cpitch * 2, (cuuint64_t)p.OW * cpitch * 2,
                        (cuuint64_t)p.OH * p.OW * cpitch * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)p.ep_cw, (cuuint32_t)p.ep_ch, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapDataType dt = plan->d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                             : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(out) failed with CUresult %d (Cout=%d OW=%d OH=%d N=%d box=%u,%u,%u)",
              (int)r, p.Cout, p.OW, p.OH, p.N, box[0], box[1], box[2]);
    return DRNB200_E_CUDA;
  }
  return DRNB200_OK;
}

// residual-projection input [N, H, W, res_pitch]: one row box of 256 pixels x 64 channels, laid out in shared
// memory exactly like the first 256 pixels of a conv input row (SWIZZLE_128B, zero fill beyond the image)
static int encode_proj_tmap(drnb200_conv_plan* plan, const void* ptr) {
  const ConvParams& p = plan->p;
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return DRNB200_E_CUDA;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)p.res_pitch, (cuuint64_t)p.OW, (cuuint64_t)p.OH, (cuuint64_t)p.N};
  cuuint64_t gstr[3] = {(cuuint64_t)p.res_pitch * 2, (cuuint64_t)p.OW * p.res_pitch * 2,
                        (cuuint64_t)p.OH * p.OW * p.res_pitch * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)kRowPx, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapDataType dt = plan->d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                             : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(&plan->tmap_r, dt, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(projection input) failed with CUresult %d (pitch=%d OW=%d OH=%d N=%d)", (int)r,
              p.res_pitch, p.OW, p.OH, p.N);
    return DRNB200_E_CUDA;
  }
  return DRNB200_OK;
}

static int encode_tmap(drnb200_conv_plan* plan, const void* x) {
  const ConvParams& p = plan->p;
  const drnb200_conv_desc& d = plan->d;
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return DRNB200_E_CUDA;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
  cuuint64_t gstr[3] = {(cuuint64_t)p.x_cpitch * 2, (cuuint64_t)p.W * p.x_cpitch * 2,
                        (cuuint64_t)p.H * p.W * p.x_cpitch * 2};
  cuuint32_t box[4] = {(cuuint32_t)p.tile_ci, (cuuint32_t)(p.TW * d.stride),
                       (cuuint32_t)(p.TH * d.stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)d.stride, (cuuint32_t)d.stride, 1};
  CUtensorMapSwizzle sw = p.pitch == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : p.pitch == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                          : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMapDataType dt = d.act_dtype == DRNB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                       : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(&plan->tmap, dt, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (Cin=%d W=%d H=%d N=%d box=%u,%u,%u)",
              (int)r, p.Cin, p.W, p.H, p.N, box[0], box[1], box[2]);
    return DRNB200_E_CUDA;
  }
  if (p.row_mode) {                       // the 8 halo pixels right of the 256-pixel row box
    cuuint32_t box2[4] = {64, (cuuint32_t)(kRowHaloPx - kRowPx), 1, 1};
    r = fn(&plan->tmap_x2, dt, 4, const_cast<void*>(x), gdim, gstr, box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(row halo) failed with CUresult %d", (int)r);
      return DRNB200_E_CUDA;
    }
  } else {
    plan->tmap_x2 = plan->tmap;
  }
  plan->tmap_ptr = x;
  return DRNB200_OK;
}

int conv_tc_launch(drnb200_conv_plan* plan, cudaStream_t st) {
  const ConvParams& p = plan->p;
  if (plan->tmap_ptr != p.x) {
    int rc = encode_tmap(plan, p.x);
    if (rc) return rc;
  }
  if (plan->tc_mode == MODE_T) {
    if (plan->tmap_y_ptr != p.y) {
      int rc = encode_out_tmap(plan, &plan->tmap_y, p.y, p.Cout);
      if (rc) return rc;
      plan->tmap_y_ptr = p.y;
    }
    if (p.proj && plan->tmap_r_ptr != p.residual) {
      int rc = encode_proj_tmap(plan, p.residual);
      if (rc) return rc;
      plan->tmap_r_ptr = p.residual;
    }
    if (p.has_res && plan->tmap_r_ptr != p.residual) {
      int rc = encode_out_tmap(plan, &plan->tmap_r, p.residual, p.res_pitch);
      if (rc) return rc;
      plan->tmap_r_ptr = p.residual;
    }
    if (!p.has_res && !p.proj && plan->tmap_r_ptr == nullptr) plan->tmap_r = plan->tmap_y;   // never dereferenced
  } else if (plan->tmap_y_ptr == nullptr) {
    plan->tmap_y = plan->tmap;                                                    // unused by these modes
    plan->tmap_r = plan->tmap;
    plan->tmap_y_ptr = p.x;
  }
  if (p.total_tiles == 0) return DRNB200_OK;
  const dim3 grid(plan->grid), block(kTcThreads);
  const bool bf = plan->d.act_dtype == DRNB200_BF16;
#define DRN_LAUNCH(M)                                                                                   \
  do {                                                                                                  \
    if (bf) launch_chained(conv_tc_kernel<M, DRNB200_BF16>, grid, block, plan->smem_bytes, st, plan->tmap, plan->tmap_y, plan->tmap_r, plan->tmap_x2, p); \
    else    launch_chained(conv_tc_kernel<M, DRNB200_F16>, grid, block, plan->smem_bytes, st, plan->tmap, plan->tmap_y, plan->tmap_r, plan->tmap_x2, p);  \
  } while (0)
#define DRN_LAUNCH_T(ROWV, NGV)                                                                                       \
  do {                                                                                                                \
    const dim3 blk(tc_threads(NGV));                                                                                  \
    if (bf) launch_chained(conv_tc_kernel<MODE_T, DRNB200_BF16, ROWV, NGV>, grid, blk, plan->smem_bytes, st, plan->tmap, plan->tmap_y, plan->tmap_r, plan->tmap_x2, p); \
    else    launch_chained(conv_tc_kernel<MODE_T, DRNB200_F16, ROWV, NGV>, grid, blk, plan->smem_bytes, st, plan->tmap, plan->tmap_y, plan->tmap_r, plan->tmap_x2, p);  \
  } while (0)
  if (plan->tc_mode == MODE_T && p.row_mode && p.pix_mode) {
    const dim3 blk(tc_threads(2));
    if (bf) launch_chained(conv_tc_kernel<MODE_T, DRNB200_BF16, true, 2, true>, grid, blk, plan->smem_bytes, st, plan->tmap, plan->tmap_y, plan->tmap_r, plan->tmap_x2, p);
    else    launch_chained(conv_tc_kernel<MODE_T, DRNB200_F16, true, 2, true>, grid, blk, plan->smem_bytes, st, plan->tmap, plan->tmap_y, plan->tmap_r, plan->tmap_x2, p);
  } else if (plan->tc_mode == MODE_T && (p.row_mode || p.ep_groups == 4)) {
    if (p.row_mode && p.ep_groups == 4) DRN_LAUNCH_T(true, 4);
    else if (p.row_mode) DRN_LAUNCH_T(true, 2);
    else DRN_LAUNCH_T(false, 4);
  } else if (plan->tc_mode == MODE_T) DRN_LAUNCH(MODE_T);
  else if (plan->tc_mode == MODE_TD) DRN_LAUNCH(MODE_TD);
  else DRN_LAUNCH(MODE_P);
#undef DRN_LAUNCH
#undef DRN_LAUNCH_T
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

}  // namespace drnb200
