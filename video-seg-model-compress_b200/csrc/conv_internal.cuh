// conv_internal.cuh — plan object and kernel parameter blocks shared by the conv translation units.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <vector>

namespace drnb200 {

// parameters common to the direct and tcgen05 kernels (passed by value)
struct ConvParams {
  const int32_t* row_ptr;   // [n_ot + 1]
  const int32_t* kblk;      // [n_live]  kb = cib * taps + tap
  const uint8_t* w_packed;  // n_live tiles of tile_o x tile_ci 16-bit, swizzled (compact.cu)
  const float* scale;       // [Cout] folded BatchNorm scale
  const float* shift;       // [Cout] folded BatchNorm shift
  const void* x;            // NHWC act_dtype
  const void* residual;     // NHWC act_dtype or null
  void* y;                  // NHWC act_dtype (or float32 when out_f32)
  int N, H, W, OH, OW, Cin, Cout;
  int taps, stride, dil;
  int tile_o, tile_ci, n_ot, n_cib;
  int relu, has_res, out_f32;
  int x_cpitch;             // channels per pixel of the tensor x lives in
  int res_pitch, res_coff;  // residual: channels per pixel of its tensor, first channel
  int relu_n;               // ReLU applies to output channels < relu_n (Cout: all, 0: none)
  int proj;                 // > 0: `residual` is a second input projected by DRNB200_KB_PROJ K-blocks (ROW kernel)
  // ---- tcgen05 path only
  const int32_t* ot_order;  // output tiles sorted by decreasing live count
  int TW, TH, tw_shift;     // pixel tile (powers of two), NT = TW*TH
  int tiles_x, tiles_y, n_pix_tiles, total_tiles;
  int stages;
  uint32_t idesc;
  uint32_t w_tile_bytes, x_tile_bytes, w_stage_bytes, stage_bytes, pitch;
  // ---- staged epilogue (MODE_T, 16-bit output): 32-pixel x 128-cout chunks through shared memory + TMA
  int ep_cw, ep_ch, ep_nch;   // chunk box (pixels wide x high), chunks per tile
  // ---- ROW variant of MODE_T (conv_tc.cu): input-row halo ring + weight-tile ring
  int row_mode, x_ring, w_ring, dbg;
  int pix_mode;               // ROW variant with the pixel-major accumulator (operands swapped, register epilogue)
  int ep_groups;              // MODE_T: epilogue groups of four warps (2 or 4)
  uint32_t main_bytes;
};

}  // namespace drnb200

struct drnb200_conv_plan {
  drnb200_conv_desc d;
  drnb200::ConvParams p;
  int impl;         // 1 direct, 2 tcgen05
  int tc_mode;      // 0 = T (M = couts, N = pixels), 1 = P (M = pixels, N = couts)
  int grid;         // persistent grid size (tcgen05)
  size_t smem_bytes;
  int64_t tile_macs;
  int32_t* d_ot_order;
  CUtensorMap tmap;        // activations, bound to `tmap_ptr`
  const void* tmap_ptr;
  CUtensorMap tmap_x2;     // ROW variant: the 8 halo pixels right of the row box (same tensor as `tmap`)
  CUtensorMap tmap_y, tmap_r;   // output / residual chunk boxes (staged epilogue), bound to the pointers below
  const void* tmap_y_ptr;
  const void* tmap_r_ptr;
  std::vector<int32_t> h_row_ptr;
  alignas(64) unsigned char gather_cache[512];   // GMapCache of conv_gather.cu (halo tensor map)
  bool gather_cache_init = false;
};

namespace drnb200 {
int conv_direct_launch(const drnb200_conv_plan* plan, cudaStream_t st);
int conv_tc_setup(drnb200_conv_plan* plan);                 // decide tile shapes, smem, attributes
int conv_tc_launch(drnb200_conv_plan* plan, cudaStream_t st);
bool conv_gather_supported(const drnb200_conv_desc& d);     // 16-channel 3x3 layers (conv_gather.cu)
int conv_gather_launch(drnb200_conv_plan* plan, cudaStream_t st);
constexpr int TC_MODE_GATHER = 3;
bool conv_halo_supported(const drnb200_conv_desc& d);       // stride-1 3x3, Cin,Cout <= 64 (conv_halo.cu)
int conv_halo_launch(drnb200_conv_plan* plan, cudaStream_t st);
constexpr int TC_MODE_HALO = 4;
bool conv_ty_supported(const drnb200_conv_desc& d);         // 3x3 stride-1 16 -> 16 (conv_ty.cu: Toeplitz along y)
int conv_ty_launch(drnb200_conv_plan* plan, cudaStream_t st);
constexpr int TC_MODE_TY = 7;
bool conv_s2_supported(const drnb200_conv_desc& d);         // 3x3 stride-2 16 -> 32 (conv_s2.cu: pixel-pair operand rows)
int conv_s2_launch(drnb200_conv_plan* plan, cudaStream_t st);
constexpr int TC_MODE_S2 = 8;
bool conv_ys_supported(const drnb200_conv_desc& d);         // 3x3 stride-1 64 -> 64 (+ residual) (conv_ys.cu: streamed input rows)
int conv_ys_launch(drnb200_conv_plan* plan, cudaStream_t st);
constexpr int TC_MODE_YS = 9;
bool conv_y2_supported(const drnb200_conv_desc& d);         // 3x3 stride-2 32 -> 128 (conv_y2.cu: streamed pixel-pair rows)
int conv_y2_launch(drnb200_conv_plan* plan, cudaStream_t st);
constexpr int TC_MODE_Y2 = 10;
// Toeplitz-weight stem (stem_tx.cu)
struct StemTxState;
int stem_tx_create(StemTxState** out, const float* w_oihw, int act_dtype, cudaStream_t st);
void stem_tx_destroy(StemTxState* s);
int stem_tx_forward(StemTxState* s, const void* x, int src_kind, const uint16_t* lut, int bgr, void* y,
                    const float* scale, const float* shift, int N, int H, int W, int act_dtype, cudaStream_t st);
}  // namespace drnb200
