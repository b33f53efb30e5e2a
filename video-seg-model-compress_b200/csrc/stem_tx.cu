// stem_tx.cu — DRN stem (7x7, 3 -> 16, stride 1, pad 3 + BN + ReLU; drn.py:132-137) as a tensor-core GEMM whose
// B operand is a banded (Toeplitz) weight matrix, so that NO im2col tile is ever built.
//
// Tile = 128 image rows x 8 columns (1024 output pixels):
//   D[y][(xo, co)] = sum over segments s = (ci, ky), e = 0..15 of  A[y][s*16 + e] * B[(xo, co)][s*16 + e]
//   A[y][s*16 + e] = x[ci][y0 + y + ky - 3][x0 - 4 + e]            (16-bit copy of the input halo)
//   B[(xo, co)][s*16 + e] = w[co][ci][ky][e - 1 - xo]  if 0 <= e - 1 - xo < 7, else 0
// i.e. the x direction of the convolution is folded into the weights (16 of 16 K-elements per segment are
// read, 7 are non-zero per output: the tensor cores have that headroom), and the y direction is a ROW SHIFT
// of the A operand: segment (ci, ky) is the UMMA descriptor that starts ky rows (ky*32 bytes) into plane ci of
// the halo [3][134 rows][16 x 16-bit], SWIZZLE_32B.  (Measured on B200: the UMMA swizzle XOR uses absolute
// shared-memory address bits, so such shifted windows need no base offset; see conv_halo.cu.)
// Per tile: one TMA box {16, 134, 3} of fp32 input -> 8 warps convert it to the 16-bit SWIZZLE_32B halo ->
// 21 tcgen05.mma (M=128 rows, N=128 = 8 pixels x 16 couts, K=16) -> 4+4 epilogue warps (thread = image row:
// 8 pixels x 16 channels = 256 contiguous bytes).  The im2col-gather stem this replaces moved 147 elements
// per output pixel through shared memory (0.87 ms per 8-frame batch); this one moves ~6.
#include "conv_internal.cuh"
#include <algorithm>
#include <cudaTypedefs.h>
#include <new>

namespace drnb200 {

constexpr int TX_ROWS = 128;                 // tile height (UMMA M)
constexpr int TX_COLS = 8;                   // tile width in pixels
constexpr int TX_HROWS = TX_ROWS + 6;        // halo rows
constexpr int TX_HW = 16;                    // halo columns: x0-4 .. x0+11
constexpr int TX_SEG = 21;                   // (ci, ky) segments = K-steps of 16
constexpr int TX_PLANE = 17 * 256;           // 134 rows x 32 B, rounded up to the 256-byte swizzle period
constexpr int TX_A_BYTES = 3 * TX_PLANE;     // 16-bit halo (UMMA A operand)
constexpr int TX_F_BYTES = 3 * TX_HROWS * TX_HW * 4;   // fp32 halo as TMA delivers it
constexpr int TX_B_BYTES = TX_SEG * 128 * 32;          // resident Toeplitz weights
constexpr int TX_RING = 3;                   // fp32 halo ring
constexpr int TX_ABUF = 2;                   // 16-bit halo buffers
constexpr int TX_ACC = 2;                    // TMEM accumulators (128 columns each)
constexpr int TX_CVT_WARPS = 8;
constexpr int TX_W_MMA = TX_CVT_WARPS, TX_W_TMA = TX_CVT_WARPS + 1, TX_W_EPI = TX_CVT_WARPS + 2;
constexpr int TX_THREADS = (TX_W_EPI + 8) * 32;

struct StemTxParams {
  const float* x;
  uint16_t* y;
  const uint8_t* w_packed;      // 21 tiles of 128 x 16 (32-byte rows, SWIZZLE_32B)
  const float* scale;
  const float* shift;
  int N, H, W, tiles_x, tiles_y, total_tiles;
  uint32_t magic_x, magic_y, idesc;
};

struct __align__(16) TxSync {
  alignas(16) float scale[16];
  alignas(16) float shift[16];
  uint64_t f_full[TX_RING], f_empty[TX_RING], a_full[TX_ABUF], a_empty[TX_ABUF], t_full[TX_ACC],
      t_empty[TX_ACC], w_full;
  uint32_t tmem_base, pad;
};

struct TxTile { int n, x0, y0; };
__device__ __forceinline__ TxTile tx_decode(const StemTxParams& p, int t) {
  TxTile c;
  const int q1 = p.tiles_x == 1 ? t : (int)__umulhi((uint32_t)t, p.magic_x);
  const int txi = t - q1 * p.tiles_x;
  c.n = p.tiles_y == 1 ? q1 : (int)__umulhi((uint32_t)q1, p.magic_y);
  const int tyi = q1 - c.n * p.tiles_y;
  c.x0 = txi * TX_COLS; c.y0 = tyi * TX_ROWS;
  return c;
}

template <int DT> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<DRNB200_F16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <> __device__ __forceinline__ uint32_t pack2<DRNB200_BF16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int DT>
__global__ void __launch_bounds__(TX_THREADS, 1)
stem_tx_kernel(const __grid_constant__ CUtensorMap tmap_x, const StemTxParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* abuf = smem;                                          // TX_ABUF x TX_A_BYTES (1024-aligned)
  uint8_t* wsm = abuf + TX_ABUF * ((TX_A_BYTES + 1023) & ~1023); // TX_B_BYTES
  uint8_t* fbuf = wsm + TX_B_BYTES;                              // TX_RING x fp32 halo
  constexpr int F_STRIDE = (TX_F_BYTES + 127) & ~127;
  TxSync* sync = reinterpret_cast<TxSync*>(fbuf + TX_RING * F_STRIDE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int b = 0; b < TX_RING; ++b) { mbar_init(&sync->f_full[b], 1); mbar_init(&sync->f_empty[b], TX_CVT_WARPS); }
    for (int b = 0; b < TX_ABUF; ++b) { mbar_init(&sync->a_full[b], TX_CVT_WARPS); mbar_init(&sync->a_empty[b], 1); }
    for (int b = 0; b < TX_ACC; ++b) { mbar_init(&sync->t_full[b], 1); mbar_init(&sync->t_empty[b], 4); }
    mbar_init(&sync->w_full, 1);
    mbar_fence_init();
  }
  if (tid < 16) { sync->scale[tid] = __ldg(p.scale + tid); sync->shift[tid] = __ldg(p.shift + tid); }
  if (warp == TX_W_MMA) {
    tmem_alloc(&sync->tmem_base, TX_ACC * 128);
    tmem_relinquish();
  }
  // the pad rows/bytes of the 16-bit halo planes are never written by the converters: clear them once
  for (int i = tid; i < TX_ABUF * ((TX_A_BYTES + 1023) & ~1023) / 16; i += TX_THREADS)
    reinterpret_cast<uint4*>(abuf)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sync->tmem_base;

  if (warp == TX_W_TMA) {
    // ================================================================= TMA producer (warp-uniform loop)
    if (elect_one()) {
      mbar_arrive_expect_tx(&sync->w_full, TX_B_BYTES);
      bulk_load(p.w_packed, &sync->w_full, wsm, TX_B_BYTES);
    }
    int b = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TxTile c = tx_decode(p, t);
      mbar_wait(&sync->f_empty[b], ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sync->f_full[b], TX_F_BYTES);
        // tensor {W, H, 3, N} fp32; box {16, 134, 3, 1}; x origin x0-4 keeps the box 16-byte aligned;
        // elements outside the image are zero-filled = the conv padding
        tma_load_4d(&tmap_x, &sync->f_full[b], fbuf + b * F_STRIDE, c.x0 - 4, c.y0 - 3, 0, c.n);
      }
      __syncwarp();
      if (++b == TX_RING) { b = 0; ph ^= 1u; }
    }
  } else if (warp == TX_W_MMA) {
    // ================================================================= MMA issuer (warp-uniform loop)
    mbar_wait(&sync->w_full, 0);
    const uint64_t d_hi = umma_smem_desc(0u, 32);           // K-major, 32-byte rows, SBO = 256
    const uint32_t w16 = smem_u32(wsm) >> 4;
    int i = 0, b = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const int acc = i % TX_ACC;
      mbar_wait(&sync->a_full[b], ph);
      mbar_wait(&sync->t_empty[acc], ((uint32_t)(i / TX_ACC) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t a16 = smem_u32(abuf + b * ((TX_A_BYTES + 1023) & ~1023)) >> 4;
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < TX_SEG; ++s) {
          const int ci = s / 7, ky = s % 7;                 // A window: plane ci, shifted down by ky rows
          umma_f16(tmem_base + (uint32_t)acc * 128u,
                   d_hi | (uint64_t)(a16 + (uint32_t)(ci * TX_PLANE + ky * 32) / 16u),
                   d_hi | (uint64_t)(w16 + (uint32_t)s * (128u * 32u / 16u)), p.idesc, s > 0 ? 1u : 0u);
        }
        umma_commit(&sync->a_empty[b]);
        umma_commit(&sync->t_full[acc]);
      }
      __syncwarp();
      if (++b == TX_ABUF) { b = 0; ph ^= 1u; }
    }
  } else if (warp >= TX_W_EPI) {
    // ================================================================= epilogue: two groups, alternate tiles
    const int q = warp & 3, grp = (warp - TX_W_EPI) >> 2;
    const int r = q * 32 + lane;                            // TMEM lane = image row of the tile
    for (int i = grp, t = blockIdx.x + grp * gridDim.x; t < p.total_tiles; t += 2 * gridDim.x, i += 2) {
      const int acc = i % TX_ACC;
      const TxTile c = tx_decode(p, t);
      const int y = c.y0 + r;
      const bool valid = y < p.H;
      uint16_t* yp = p.y + (((size_t)c.n * p.H + y) * p.W + c.x0) * 16;
      mbar_wait(&sync->t_full[acc], (uint32_t)(i / TX_ACC) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * 128u + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int xo = 0; xo < TX_COLS; ++xo) {
        uint32_t v[16];
        tmem_ld16(t_addr + (uint32_t)(xo * 16), v);
        tmem_ld_wait();
        if (valid && c.x0 + xo < p.W) {
          uint32_t w[8];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 sc = *reinterpret_cast<const float4*>(&sync->scale[4 * e4]);
            const float4 sh = *reinterpret_cast<const float4*>(&sync->shift[4 * e4]);
            const float a0 = fmaxf(fmaf(__uint_as_float(v[4 * e4]), sc.x, sh.x), 0.f);
            const float a1 = fmaxf(fmaf(__uint_as_float(v[4 * e4 + 1]), sc.y, sh.y), 0.f);
            const float a2 = fmaxf(fmaf(__uint_as_float(v[4 * e4 + 2]), sc.z, sh.z), 0.f);
            const float a3 = fmaxf(fmaf(__uint_as_float(v[4 * e4 + 3]), sc.w, sh.w), 0.f);
            w[2 * e4] = pack2<DT>(a0, a1);
            w[2 * e4 + 1] = pack2<DT>(a2, a3);
          }
          uint4* o = reinterpret_cast<uint4*>(yp + xo * 16);
          o[0] = make_uint4(w[0], w[1], w[2], w[3]);
          o[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sync->t_empty[acc]);
    }
  } else {
    // ================================================================= convert: fp32 halo -> 16-bit SWIZZLE_32B halo
    // work item = (plane ci, halo row hr, 16-byte chunk c): 8 floats -> 8 x 16-bit
    int b = 0, fb = 0;
    uint32_t ph = 0, fph = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(&sync->f_full[fb], fph);
      mbar_wait(&sync->a_empty[b], ph ^ 1u);
      const float* f = reinterpret_cast<const float*>(fbuf + fb * F_STRIDE);
      uint8_t* a = abuf + b * ((TX_A_BYTES + 1023) & ~1023);
      for (int it = tid; it < 3 * TX_HROWS * 2; it += TX_CVT_WARPS * 32) {
        const int c = it & 1, hr = (it >> 1) % TX_HROWS, ci = (it >> 1) / TX_HROWS;
        const float4* src = reinterpret_cast<const float4*>(f + (ci * TX_HROWS + hr) * TX_HW + c * 8);
        const float4 u = src[0], v = src[1];
        *reinterpret_cast<uint4*>(a + ci * TX_PLANE + swz_offset((uint32_t)hr, (uint32_t)c, 32)) =
            make_uint4(pack2<DT>(u.x, u.y), pack2<DT>(u.z, u.w), pack2<DT>(v.x, v.y), pack2<DT>(v.z, v.w));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sync->a_full[b]);
        mbar_arrive(&sync->f_empty[fb]);
      }
      if (++b == TX_ABUF) { b = 0; ph ^= 1u; }
      if (++fb == TX_RING) { fb = 0; fph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TX_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TX_ACC * 128);
  }
}

// Toeplitz weight matrix [128 = (xo, co)][336 = (ci, ky, e)] from the OIHW stem weights
__global__ void stem_tx_weights_kernel(const float* __restrict__ w, float* __restrict__ wt,
                                       int32_t* __restrict__ row_ptr, int32_t* __restrict__ kblk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128 * 336) {
    const int nrow = i / 336, k = i - nrow * 336;
    const int xo = nrow >> 4, co = nrow & 15;
    const int s = k >> 4, e = k & 15, ci = s / 7, ky = s - ci * 7, kx = e - 1 - xo;
    wt[i] = (kx >= 0 && kx < 7) ? __ldg(w + ((co * 3 + ci) * 7 + ky) * 7 + kx) : 0.f;
  }
  if (i < TX_SEG) kblk[i] = i;
  if (i == 0) { row_ptr[0] = 0; row_ptr[1] = TX_SEG; }
}

static PFN_cuTensorMapEncodeTiled_v12000 tx_encode_fn() {
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

// ---- host API used by drnb200_stem_plan_* (conv_gather.cu) ------------------------------------------------
struct StemTxState {
  float* d_wt = nullptr;
  int32_t *d_row_ptr = nullptr, *d_kblk = nullptr;
  uint16_t* d_wpacked = nullptr;
  const void* map_ptr = nullptr;
  CUtensorMap map;
};

int stem_tx_create(StemTxState** out, const float* w_oihw, int act_dtype, cudaStream_t st) {
  StemTxState* s = new (std::nothrow) StemTxState();
  if (!s) { set_error("stem_tx: out of host memory"); return DRNB200_E_NOMEM; }
  cudaError_t e = cudaMalloc((void**)&s->d_wt, 128 * 336 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_row_ptr, 2 * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_kblk, TX_SEG * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_wpacked, TX_B_BYTES);
  if (e != cudaSuccess) { delete s; return cuda_fail(e, "cudaMalloc(stem_tx)"); }
  stem_tx_weights_kernel<<<(128 * 336 + 255) / 256, 256, 0, st>>>(w_oihw, s->d_wt, s->d_row_ptr, s->d_kblk);
  int rc = drnb200_pack_weights(s->d_wt, nullptr, 128, 336, 1, 1, 128, 16, s->d_row_ptr, s->d_kblk, act_dtype,
                                s->d_wpacked, (void*)st);
  if (rc) { delete s; return rc; }
  *out = s;
  return DRNB200_OK;
}

void stem_tx_destroy(StemTxState* s) {
  if (!s) return;
  cudaFree(s->d_wt); cudaFree(s->d_row_ptr); cudaFree(s->d_kblk); cudaFree(s->d_wpacked);
  delete s;
}

int stem_tx_forward(StemTxState* s, const float* x, void* y, const float* scale, const float* shift, int N,
                    int H, int W, int act_dtype, cudaStream_t st) {
  StemTxParams p{};
  p.x = x; p.y = reinterpret_cast<uint16_t*>(y); p.w_packed = reinterpret_cast<const uint8_t*>(s->d_wpacked);
  p.scale = scale; p.shift = shift; p.N = N; p.H = H; p.W = W;
  p.tiles_x = (W + TX_COLS - 1) / TX_COLS;
  p.tiles_y = (H + TX_ROWS - 1) / TX_ROWS;
  p.total_tiles = N * p.tiles_x * p.tiles_y;
  if ((uint64_t)p.total_tiles * (uint64_t)std::max(p.tiles_x, p.tiles_y) >= (1ull << 32)) {
    set_error("stem_tx: problem too large for the 32-bit tile decode");
    return DRNB200_E_ARG;
  }
  p.magic_x = p.tiles_x == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_x - 1) / p.tiles_x);
  p.magic_y = p.tiles_y == 1 ? 0u : (uint32_t)(((1ull << 32) + p.tiles_y - 1) / p.tiles_y);
  p.idesc = umma_idesc_f16(128, 128, act_dtype);
  if (s->map_ptr != x) {
    auto fn = tx_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return DRNB200_E_CUDA; }
    cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * 12};
    cuuint32_t box[4] = {(cuuint32_t)TX_HW, (cuuint32_t)TX_HROWS, 3, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&s->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(stem_tx) failed with CUresult %d (W=%d H=%d N=%d)", (int)r, W, H, N);
      return DRNB200_E_CUDA;
    }
    s->map_ptr = x;
  }
  constexpr int F_STRIDE = (TX_F_BYTES + 127) & ~127;
  const size_t smem = 1024 + TX_ABUF * ((TX_A_BYTES + 1023) & ~1023) + TX_B_BYTES + TX_RING * F_STRIDE + sizeof(TxSync);
  static bool attr[2] = {false, false};
  if (!attr[act_dtype]) {
    if (act_dtype == DRNB200_BF16)
      DRN_CUDA(cudaFuncSetAttribute(stem_tx_kernel<DRNB200_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      DRN_CUDA(cudaFuncSetAttribute(stem_tx_kernel<DRNB200_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[act_dtype] = true;
  }
  int dev = 0, sms = 148;
  DRN_CUDA(cudaGetDevice(&dev));
  DRN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.total_tiles, sms);
  if (grid == 0) return DRNB200_OK;
  if (act_dtype == DRNB200_BF16) stem_tx_kernel<DRNB200_BF16><<<grid, TX_THREADS, smem, st>>>(s->map, p);
  else stem_tx_kernel<DRNB200_F16><<<grid, TX_THREADS, smem, st>>>(s->map, p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

}  // namespace drnb200
