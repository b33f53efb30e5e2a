// multiscale.cu — the reference's multi-scale test on the device, semantic_seg.py:507-557:
//   final = sum([resize_4d_tensor(out, w, h) for out in outputs]);  pred = final.argmax(axis=1)
// resize_4d_tensor (semantic_seg.py:471-504) resamples every float32 plane with Pillow's BILINEAR filter
// (src/libImaging/Resample.c: a horizontal pass into a float32 temporary, then a vertical pass; per output sample
// `ss = 0.0; ss += pixel * k[x]` in double over the taps in order, stored as float32; when downscaling the triangle
// filter is widened by the scale factor).  The coefficient tables (xmin, count, k[] in double) are computed on the
// host exactly as precompute_coeffs() does and passed in; the kernels repeat the same separately rounded double
// multiplies and adds, so the result is bit-identical to the reference's (tests: against oracle/ms_oracle.py, which is
// pinned to fixtures produced by the real resize_4d_tensor).
//
// ms_accumulate_kernel: thread = one output column x of one plane and a strip of TY output rows.  Phase 1 resamples
// every source row the strip needs horizontally at x (float32 — the value Pillow's temporary image would hold) into
// the thread's own column of a shared-memory scratch [rows][128]; phase 2 runs the vertical taps of the TY output
// rows over that column (tap order preserved) and adds into the accumulator.  A thread only ever reads what it wrote,
// so no barrier is needed.  The first version re-tested all TY rows for every source row and was instruction-issue
// bound (ncu: 14 instructions per FP64 operation, 0.76 ms for the 1.75x scale of a 1024x2048x19 frame); this one
// unrolls the taps (KMAX) and touches each (row, tap) pair once.  HBM-bound by design: the source plane is read
// ~once (neighbouring threads share taps through L1), the accumulator is read (not on the first scale) and written.
#include "common.cuh"

namespace drnb200 {

struct MsParams {
  const float* src; float* acc;
  int Hs, Ws, H, W;
  const int32_t* xmin; const int32_t* xcnt; const double* xk; int kx;   // kx == 0: no horizontal pass (Ws == W)
  const int32_t* ymin; const int32_t* ycnt; const double* yk; int ky;   // ky == 0: no vertical pass (Hs == H)
  int first, rows_max;
};

// KMAX > 0: both axes have at most KMAX taps (loops fully unrolled, predicated); KMAX == 0: run-time tap loops
template <int KMAX, int TY>
__global__ void __launch_bounds__(128) ms_accumulate_kernel(const MsParams p) {
  extern __shared__ float tmp[];                    // [rows_max][128]: column threadIdx.x belongs to this thread
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= p.W) return;
  const int y0 = blockIdx.y * TY;
  const int ny = min(TY, p.H - y0);
  const float* s = p.src + (size_t)blockIdx.z * p.Hs * p.Ws;
  float* a = p.acc + (size_t)blockIdx.z * p.H * p.W + (size_t)y0 * p.W + x;
  float* col = tmp + threadIdx.x;
  constexpr int KR = KMAX > 0 ? KMAX : 1;

  // the accumulator's old values are fetched first: their HBM latency hides behind phase 1 (ncu on the previous
  // version: 70 % of the stall samples sat on the FADD waiting for `*dst`)
  float old[TY];
#pragma unroll
  for (int j = 0; j < TY; ++j) old[j] = (!p.first && j < ny) ? __ldcs(a + (size_t)j * p.W) : 0.0f;

  int r_lo = y0, r_hi = y0 + ny;
  if (p.ky) {
    r_lo = __ldg(p.ymin + y0);
    r_hi = __ldg(p.ymin + y0 + ny - 1) + __ldg(p.ycnt + y0 + ny - 1);
    r_hi = min(r_hi, r_lo + p.rows_max);            // tables that are not Pillow's cannot overrun the scratch
  }
  // ---- phase 1: horizontal pass of rows [r_lo, r_hi) at column x
  if (p.kx) {
    const int xm = __ldg(p.xmin + x), xc = __ldg(p.xcnt + x);
    const double* xk = p.xk + (size_t)x * p.kx;
    double kr[KR];
    if (KMAX > 0) {
#pragma unroll
      for (int t = 0; t < KR; ++t) kr[t] = t < xc ? __ldg(xk + t) : 0.0;
    }
    const float* row = s + (size_t)r_lo * p.Ws + xm;
#pragma unroll 2
    for (int r = r_lo; r < r_hi; ++r, row += p.Ws) {
      double ss = 0.0;                              // ss += pixel * k[t]: separately rounded, never an FMA
      if (KMAX > 0) {
        float px[KR];
#pragma unroll
        for (int t = 0; t < KR; ++t) px[t] = t < xc ? __ldg(row + t) : 0.f;
#pragma unroll
        for (int t = 0; t < KR; ++t)
          if (t < xc) ss = __dadd_rn(ss, __dmul_rn((double)px[t], kr[t]));
      } else {
        for (int t = 0; t < xc; ++t) ss = __dadd_rn(ss, __dmul_rn((double)__ldg(row + t), __ldg(xk + t)));
      }
      col[(r - r_lo) * 128] = __double2float_rn(ss);
    }
  } else {
    const float* row = s + (size_t)r_lo * p.Ws + x;
#pragma unroll 4
    for (int r = r_lo; r < r_hi; ++r, row += p.Ws) col[(r - r_lo) * 128] = __ldg(row);
  }
  // ---- phase 2: vertical pass + accumulate (Python's sum() starts from 0: (0 + r0) + r1 + ..., semantic_seg.py:540)
#pragma unroll
  for (int j = 0; j < TY; ++j) {
    if (j >= ny) break;
    float out;
    if (p.ky) {
      const int ym = __ldg(p.ymin + y0 + j) - r_lo, yc = __ldg(p.ycnt + y0 + j);
      const double* yk = p.yk + (size_t)(y0 + j) * p.ky;
      double ss = 0.0;
      if (KMAX > 0) {
#pragma unroll
        for (int t = 0; t < KR; ++t)
          if (t < yc && ym + t < p.rows_max)
            ss = __dadd_rn(ss, __dmul_rn((double)col[(ym + t) * 128], __ldg(yk + t)));
      } else {
        for (int t = 0; t < yc && ym + t < p.rows_max; ++t)
          ss = __dadd_rn(ss, __dmul_rn((double)col[(ym + t) * 128], __ldg(yk + t)));
      }
      out = __double2float_rn(ss);
    } else {
      out = col[j * 128];
    }
    a[(size_t)j * p.W] = __fadd_rn(old[j], out);
  }
}

// pred = final.argmax(axis=1): first maximum wins (strict >), uint8 labels.  thread = V adjacent pixels.
template <int V>
__global__ void __launch_bounds__(256) ms_argmax_kernel(const float* __restrict__ acc, int C, int64_t plane,
                                                        int64_t n_items, uint8_t* __restrict__ labels) {
  const int64_t per_frame = plane / V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / per_frame, q = i - n * per_frame;
    const float* base = acc + (size_t)n * C * plane + (size_t)q * V;
    float best[V];
    uint32_t arg[V];
    if (V == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(base));
      best[0] = v.x; best[1 % V] = v.y; best[2 % V] = v.z; best[3 % V] = v.w;
    } else {
      best[0] = __ldg(base);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) arg[j] = 0u;
    for (int c = 1; c < C; ++c) {
      float cur[V];
      if (V == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)c * plane));
        cur[0] = v.x; cur[1 % V] = v.y; cur[2 % V] = v.z; cur[3 % V] = v.w;
      } else {
        cur[0] = __ldg(base + (size_t)c * plane);
      }
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (cur[j] > best[j]) { best[j] = cur[j]; arg[j] = (uint32_t)c; }
    }
    if (V == 4)
      reinterpret_cast<uint32_t*>(labels)[(size_t)n * per_frame + q] = arg[0] | (arg[1 % V] << 8) | (arg[2 % V] << 16) | (arg[3 % V] << 24);
    else
      labels[(size_t)n * plane + q] = (uint8_t)arg[0];
  }
}

}  // namespace drnb200

using namespace drnb200;

extern "C" int drnb200_ms_accumulate(const float* src, int N, int C, int Hs, int Ws, float* acc, int H, int W,
                                     const int32_t* xmin, const int32_t* xcnt, const double* xk, int kx,
                                     const int32_t* ymin, const int32_t* ycnt, const double* yk, int ky,
                                     int first, void* stream) {
  DRN_REQUIRE(src && acc, "ms_accumulate: null pointer");
  DRN_REQUIRE(N >= 0 && C > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "ms_accumulate: bad shape");
  DRN_REQUIRE(kx >= 0 && ky >= 0, "ms_accumulate: negative tap count");
  DRN_REQUIRE(kx ? (xmin && xcnt && xk) : (Ws == W),
              "ms_accumulate: the horizontal pass needs coefficient tables (Ws=%d, W=%d, kx=%d)", Ws, W, kx);
  DRN_REQUIRE(ky ? (ymin && ycnt && yk) : (Hs == H),
              "ms_accumulate: the vertical pass needs coefficient tables (Hs=%d, H=%d, ky=%d)", Hs, H, ky);
  DRN_REQUIRE((long long)N * C <= 65535, "ms_accumulate: N*C = %lld planes exceed the grid limit", (long long)N * C);
  if (N == 0) return DRNB200_OK;
  // rows of scratch a strip of TY output rows can need: (TY-1)*scale + 2*support + 1 source rows, rounded up
  const double scale = (double)Hs / (double)H, support = scale > 1.0 ? scale : 1.0;
  auto rows_for = [&](int ty) { return ky ? (int)((ty - 1) * scale + 2.0 * support) + 2 : ty; };
  int ty = 16;
  while (ty > 1 && (size_t)rows_for(ty) * 512 > 48 * 1024) ty /= 4;
  DRN_REQUIRE((size_t)rows_for(ty) * 512 <= 48 * 1024,
              "ms_accumulate: vertical scale factor %.1f too large for the shared-memory scratch", scale);
  const int kmax = (kx <= 3 && ky <= 3) ? 3 : (kx <= 5 && ky <= 5) ? 5 : 0;
  MsParams p{src, acc, Hs, Ws, H, W, xmin, xcnt, xk, kx, ymin, ycnt, yk, ky, first, rows_for(ty)};
  const dim3 grid((W + 127) / 128, (H + ty - 1) / ty, N * C);
  DRN_REQUIRE(grid.y <= 65535, "ms_accumulate: H=%d too large", H);
  const size_t smem = (size_t)p.rows_max * 512;
  cudaStream_t st = (cudaStream_t)stream;
  if (ty == 16 && kmax == 3) ms_accumulate_kernel<3, 16><<<grid, 128, smem, st>>>(p);
  else if (ty == 16 && kmax == 5) ms_accumulate_kernel<5, 16><<<grid, 128, smem, st>>>(p);
  else if (ty == 16) ms_accumulate_kernel<0, 16><<<grid, 128, smem, st>>>(p);
  else if (ty == 4) ms_accumulate_kernel<0, 4><<<grid, 128, smem, st>>>(p);
  else ms_accumulate_kernel<0, 1><<<grid, 128, smem, st>>>(p);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}

extern "C" int drnb200_ms_argmax(const float* acc, int N, int C, int H, int W, uint8_t* labels, void* stream) {
  DRN_REQUIRE(acc && labels, "ms_argmax: null pointer");
  DRN_REQUIRE(N >= 0 && C > 0 && C <= 256 && H > 0 && W > 0, "ms_argmax: bad shape (classes must fit uint8)");
  if (N == 0) return DRNB200_OK;
  const int64_t plane = (int64_t)H * W;
  const bool v4 = plane % 4 == 0 && (reinterpret_cast<uintptr_t>(acc) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(labels) & 3u) == 0;
  const int64_t items = v4 ? (int64_t)N * (plane / 4) : (int64_t)N * plane;
  int64_t blocks = (items + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (v4) ms_argmax_kernel<4><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(acc, C, plane, items, labels);
  else ms_argmax_kernel<1><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(acc, C, plane, items, labels);
  DRN_CUDA(cudaGetLastError());
  return DRNB200_OK;
}
