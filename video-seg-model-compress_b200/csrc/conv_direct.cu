// conv_direct.cu — CUDA-core direct convolution over the same tile list / packed weights as the
// tcgen05 path.  It accepts any (Cin, Cout) the packer accepts, so it serves (i) layers whose shape the
// tensor-core kernel does not take and (ii) as the on-device cross-check of the tcgen05 kernel at sizes
// where the CPU oracle would take minutes.  Semantics: drn.py:49-65 / :201-211 (conv -> BN -> [+res] -> ReLU).
#include "conv_internal.cuh"

namespace drnb200 {

// block = (32 pixels, CG cout-groups); each thread: 1 output pixel x 8 output channels, fp32 accumulate.
template <int DT>
__global__ void __launch_bounds__(128) conv_direct_kernel(const ConvParams p) {
  const int64_t P = (int64_t)p.N * p.OH * p.OW;
  const int64_t pix = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int c0 = (blockIdx.y * blockDim.y + threadIdx.y) * 8;
  if (pix >= P || c0 >= p.Cout) return;
  const int ox = (int)(pix % p.OW);
  const int oy = (int)((pix / p.OW) % p.OH);
  const int n = (int)(pix / ((int64_t)p.OW * p.OH));
  const int ot = c0 / p.tile_o, r0 = c0 - ot * p.tile_o;
  const uint32_t pitch = (uint32_t)p.tile_ci * 2u;
  const size_t tile_bytes = (size_t)p.tile_o * pitch;
  const uint16_t* x = reinterpret_cast<const uint16_t*>(p.x);

  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;

  const int jb = p.row_ptr[ot], je = p.row_ptr[ot + 1];
  const int half = (p.taps == 9) ? 1 : 0;  // 3x3: taps offset by (k-1)*dil, 1x1: none
  for (int j = jb; j < je; ++j) {
    const int kb = __ldg(p.kblk + j);
    const int cib = kb / p.taps, tap = kb - cib * p.taps;
    const int ky = (p.taps == 9) ? tap / 3 : 0, kx = (p.taps == 9) ? tap - ky * 3 : 0;
    const int iy = oy * p.stride + (ky - half) * p.dil;
    const int ix = ox * p.stride + (kx - half) * p.dil;
    if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;  // zero padding
    const uint16_t* xp = x + (((size_t)n * p.H + iy) * p.W + ix) * p.x_cpitch + (size_t)cib * p.tile_ci;
    const uint8_t* tile = p.w_packed + (size_t)j * tile_bytes;
    for (int k8 = 0; k8 < p.tile_ci / 8; ++k8) {
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xp) + k8);
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
      float xf[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        xf[2 * q] = Act<DT>::to_f32((uint16_t)(xw[q] & 0xFFFFu));
        xf[2 * q + 1] = Act<DT>::to_f32((uint16_t)(xw[q] >> 16));
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(
            tile + swz_offset((uint32_t)(r0 + r), (uint32_t)k8, pitch)));
        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[r] = fmaf(xf[2 * q], Act<DT>::to_f32((uint16_t)(ww[q] & 0xFFFFu)), acc[r]);
          acc[r] = fmaf(xf[2 * q + 1], Act<DT>::to_f32((uint16_t)(ww[q] >> 16)), acc[r]);
        }
      }
    }
  }

  const size_t off = (size_t)pix * p.Cout + c0;
  float out[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) out[r] = fmaf(acc[r], __ldg(p.scale + c0 + r), __ldg(p.shift + c0 + r));
  if (p.has_res) {
    const uint4 rv = __ldg(reinterpret_cast<const uint4*>(
        reinterpret_cast<const uint16_t*>(p.residual) + (size_t)pix * p.res_pitch + p.res_coff + c0));
    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      out[2 * q] += Act<DT>::to_f32((uint16_t)(rw[q] & 0xFFFFu));
      out[2 * q + 1] += Act<DT>::to_f32((uint16_t)(rw[q] >> 16));
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (c0 + r < p.relu_n) out[r] = fmaxf(out[r], 0.f);
  if (p.out_f32) {
    float4* yp = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + off);
    yp[0] = make_float4(out[0], out[1], out[2], out[3]);
    yp[1] = make_float4(out[4], out[5], out[6], out[7]);
  } else {
    uint4 o;
    o.x = (uint32_t)Act<DT>::from_f32(out[0]) | ((uint32_t)Act<DT>::from_f32(out[1]) << 16);
    o.y = (uint32_t)Act<DT>::from_f32(out[2]) | ((uint32_t)Act<DT>::from_f32(out[3]) << 16);
    o.z = (uint32_t)Act<DT>::from_f32(out[4]) | ((uint32_t)Act<DT>::from_f32(out[5]) << 16);
    o.w = (uint32_t)Act<DT>::from_f32(out[6]) | ((uint32_t)Act<DT>::from_f32(out[7]) << 16);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y) + off) = o;
  }
}

int conv_direct_launch(const drnb200_conv_plan* plan, cudaStream_t st) {
  const ConvParams& p = plan->p;
  const int64_t P = (int64_t)p.N * p.OH * p.OW;
  const int groups = p.Cout / 8;
  const int cg = groups >= 4 ? 4 : groups;
  dim3 block(32, cg);
  dim3 grid((unsigned)((P + 31) / 32), (unsigned)((groups + cg - 1) / cg));
  if (plan->d.act_dtype == DRNB200_BF16)
    conv_direct_kernel<DRNB200_BF16><<<grid, block, 0, st>>>(p);
  else
    conv_direct_kernel<DRNB200_F16><<<grid, block, 0, st>>>(p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

}  // namespace drnb200
