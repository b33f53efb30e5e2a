// compact.cu — (a) mask -> block-sparse tile list, and packing of the live weight blocks.
//
// Reference semantics restated in oracle/compact_oracle.py:
//   matricised weight (O, I*kh*kw), column = ci*kh*kw + tap           (pruners/BlockPruner.py:144)
//   block live <=> any element != 0                                   (tools/visualize_layers.py:8-13)
//   indices / rowBlockPtr form of the list                            (pruners/BlockPruner.py:344-413)
// Both kernels are one-time, HBM-bound passes: the mask is read exactly once (4*O*I*kh*kw bytes).
#include "common.cuh"

namespace drnb200 {

// grid = (n_cib, O / rows_per_cta).  One CTA scans `rows_per_cta` rows of the tile_o x (tile_ci*taps) slab of the
// mask that belongs to (ot, cib); along a mask row the slab is contiguous, so consecutive threads read consecutive
// floats (16-byte loads when the slab is 16-byte aligned).  The host picks rows_per_cta so that a 512x512x3x3 mask
// (9.4 MB) is spread over >= 4 CTAs per SM: the pass is HBM-bound instead of latency-bound on n_cib*n_ot CTAs.
// `live` is zeroed by the caller; CTAs of the same (ot, cib) only ever store 1 (no atomics needed).
template <bool VEC4>
__global__ void __launch_bounds__(256)
compact_flags_kernel(const float* __restrict__ mask, int I, int taps, int tile_o, int tile_ci, int n_kb,
                     int rows_per_cta, uint8_t* __restrict__ live) {
  __shared__ unsigned int sflags[2];   // bit t of the 64-bit pair = tap t has a non-zero element
  const int cib = blockIdx.x, r0 = blockIdx.y * rows_per_cta, ot = r0 / tile_o;
  if (threadIdx.x < 2) sflags[threadIdx.x] = 0u;
  __syncthreads();
  const int slab_w = tile_ci * taps;  // contiguous floats per row
  const size_t row_len = (size_t)I * taps;
  const float* base = mask + (size_t)r0 * row_len + (size_t)cib * slab_w;
  unsigned long long f = 0ull;
  if (VEC4) {
    const int qw = slab_w >> 2, total = rows_per_cta * qw;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int r = idx / qw, q = idx - r * qw;
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)r * row_len) + q);
      const int t0 = (q << 2) % taps;
      int t1 = t0 + 1; t1 = t1 == taps ? 0 : t1;
      int t2 = t1 + 1; t2 = t2 == taps ? 0 : t2;
      int t3 = t2 + 1; t3 = t3 == taps ? 0 : t3;
      if (v.x != 0.0f) f |= 1ull << t0;
      if (v.y != 0.0f) f |= 1ull << t1;
      if (v.z != 0.0f) f |= 1ull << t2;
      if (v.w != 0.0f) f |= 1ull << t3;
    }
  } else {
    const int total = rows_per_cta * slab_w;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int r = idx / slab_w, c = idx - r * slab_w;
      if (__ldg(base + (size_t)r * row_len + c) != 0.0f) f |= 1ull << (c % taps);
    }
  }
  const unsigned int lo = __reduce_or_sync(0xffffffffu, (unsigned int)f);
  const unsigned int hi = __reduce_or_sync(0xffffffffu, (unsigned int)(f >> 32));
  if ((threadIdx.x & 31) == 0) {
    if (lo) atomicOr(&sflags[0], lo);
    if (hi) atomicOr(&sflags[1], hi);
  }
  __syncthreads();
  if (threadIdx.x < taps && ((sflags[threadIdx.x >> 5] >> (threadIdx.x & 31)) & 1u))
    live[(size_t)ot * n_kb + cib * taps + threadIdx.x] = (uint8_t)1;
}

// single CTA: counts per output tile, exclusive scan, ordered fill of kblk.
__global__ void compact_scan_kernel(const uint8_t* __restrict__ live, int n_ot, int n_kb,
                                    int32_t* __restrict__ row_ptr, int32_t* __restrict__ kblk,
                                    int32_t* __restrict__ n_live) {
  extern __shared__ int32_t counts[];  // n_ot + 1
  for (int ot = threadIdx.x; ot < n_ot; ot += blockDim.x) {
    int c = 0;
    for (int kb = 0; kb < n_kb; ++kb) c += live[(size_t)ot * n_kb + kb];
    counts[ot + 1] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    counts[0] = 0;
    for (int ot = 0; ot < n_ot; ++ot) counts[ot + 1] += counts[ot];
    *n_live = counts[n_ot];
  }
  __syncthreads();
  for (int ot = threadIdx.x; ot <= n_ot; ot += blockDim.x) row_ptr[ot] = counts[ot];
  for (int ot = threadIdx.x; ot < n_ot; ot += blockDim.x) {
    int w = counts[ot];
    for (int kb = 0; kb < n_kb; ++kb)
      if (live[(size_t)ot * n_kb + kb]) kblk[w++] = kb;
  }
}

// grid = n_ot * n_kb (upper bound on the number of live tiles); CTA j packs live tile j.
template <int DT>
__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ mask,
                                    int I, int taps, int tile_o, int tile_ci, int n_ot,
                                    const int32_t* __restrict__ row_ptr,
                                    const int32_t* __restrict__ kblk,
                                    uint8_t* __restrict__ packed) {
  const int j = blockIdx.x;
  if (j >= row_ptr[n_ot]) return;
  // locate the output tile that owns entry j (n_ot is small)
  int ot = 0;
  while (row_ptr[ot + 1] <= j) ++ot;
  const int kb = kblk[j];
  const int cib = kb / taps, tap = kb - cib * taps;
  const uint32_t pitch = (uint32_t)tile_ci * 2u;
  uint8_t* tile = packed + (size_t)j * tile_o * pitch;
  const size_t row_len = (size_t)I * taps;
  const int total = tile_o * tile_ci;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    int r = idx / tile_ci, k = idx - r * tile_ci;
    size_t src = (size_t)(ot * tile_o + r) * row_len + (size_t)(cib * tile_ci + k) * taps + tap;
    float v = __ldg(w + src);
    if (mask != nullptr && __ldg(mask + src) == 0.0f) v = 0.0f;
    uint32_t off = swz_offset((uint32_t)r, (uint32_t)k >> 3, pitch) + ((uint32_t)k & 7u) * 2u;
    *reinterpret_cast<uint16_t*>(tile + off) = Act<DT>::from_f32(v);
  }
}

}  // namespace drnb200

using namespace drnb200;

extern "C" int drnb200_compact_mask(const float* mask_oihw, int O, int I, int kh, int kw,
                                    int tile_o, int tile_ci, int32_t* row_ptr, int32_t* kblk,
                                    int32_t* n_live, void* stream) {
  DRN_REQUIRE(mask_oihw && row_ptr && kblk && n_live, "compact_mask: null pointer");
  DRN_REQUIRE(O > 0 && I > 0 && kh > 0 && kw > 0 && kh * kw <= 64,
              "compact_mask: bad shape O=%d I=%d k=%dx%d", O, I, kh, kw);
  DRN_REQUIRE(tile_o > 0 && tile_ci > 0 && O % tile_o == 0 && I % tile_ci == 0,
              "compact_mask: O=%d / I=%d not divisible by tile %dx%d", O, I, tile_o, tile_ci);
  const int taps = kh * kw, n_ot = O / tile_o, n_cib = I / tile_ci, n_kb = n_cib * taps;
  DRN_REQUIRE(n_ot <= 8192, "compact_mask: too many output tiles (%d)", n_ot);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* live = nullptr;
  DRN_CUDA(cudaMallocAsync((void**)&live, (size_t)n_ot * n_kb, st));
  DRN_CUDA(cudaMemsetAsync(live, 0, (size_t)n_ot * n_kb, st));
  int rows_per_cta = tile_o;       // halve while the grid has fewer than 4 CTAs per SM (148 SMs)
  while (rows_per_cta % 2 == 0 && (long long)n_cib * (O / rows_per_cta) < 4 * 148) rows_per_cta /= 2;
  DRN_REQUIRE(O / rows_per_cta <= 65535, "compact_mask: grid too large (O=%d)", O);
  const dim3 grid(n_cib, O / rows_per_cta);
  const bool vec4 = ((tile_ci * taps) % 4 == 0) && (((size_t)I * taps) % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(mask_oihw) & 15u) == 0);
  if (vec4)
    compact_flags_kernel<true><<<grid, 256, 0, st>>>(mask_oihw, I, taps, tile_o, tile_ci, n_kb, rows_per_cta, live);
  else
    compact_flags_kernel<false><<<grid, 256, 0, st>>>(mask_oihw, I, taps, tile_o, tile_ci, n_kb, rows_per_cta, live);
  compact_scan_kernel<<<1, 256, (n_ot + 1) * sizeof(int32_t), st>>>(live, n_ot, n_kb, row_ptr, kblk,
                                                                   n_live);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(live, st);
  DRN_CUDA(e);
  return DRNB200_OK;
}

extern "C" int drnb200_pack_weights(const float* w_oihw, const float* mask_oihw_or_null, int O,
                                    int I, int kh, int kw, int tile_o, int tile_ci,
                                    const int32_t* row_ptr, const int32_t* kblk, int act_dtype,
                                    uint16_t* w_packed, void* stream) {
  DRN_REQUIRE(w_oihw && row_ptr && kblk && w_packed, "pack_weights: null pointer");
  DRN_REQUIRE(tile_ci == 16 || tile_ci == 32 || tile_ci == 64,
              "pack_weights: tile_ci must be 16, 32 or 64 (got %d)", tile_ci);
  DRN_REQUIRE(tile_o > 0 && tile_o % 8 == 0 && O % tile_o == 0 && I % tile_ci == 0,
              "pack_weights: O=%d / I=%d not divisible by tile %dx%d (tile_o %% 8 == 0)", O, I,
              tile_o, tile_ci);
  DRN_REQUIRE(act_dtype == DRNB200_BF16 || act_dtype == DRNB200_F16, "pack_weights: bad act_dtype");
  const int taps = kh * kw, n_ot = O / tile_o, n_kb = (I / tile_ci) * taps;
  cudaStream_t st = (cudaStream_t)stream;
  if (act_dtype == DRNB200_BF16)
    pack_weights_kernel<DRNB200_BF16><<<n_ot * n_kb, 256, 0, st>>>(
        w_oihw, mask_oihw_or_null, I, taps, tile_o, tile_ci, n_ot, row_ptr, kblk,
        reinterpret_cast<uint8_t*>(w_packed));
  else
    pack_weights_kernel<DRNB200_F16><<<n_ot * n_kb, 256, 0, st>>>(
        w_oihw, mask_oihw_or_null, I, taps, tile_o, tile_ci, n_ot, row_ptr, kblk,
        reinterpret_cast<uint8_t*>(w_packed));
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}
