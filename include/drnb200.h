/*
 * drnb200.h — C ABI of libdrnb200.so: the B200 (sm_100a) inference path for block-pruned DRN
 * semantic segmentation.
 *
 * The reference (thejasvi-konduru/video-seg-model-compress) is pure Python/PyTorch and defines no
 * FFI of its own; the boundary it exposes for this path is two Python interfaces, `DRNSeg.forward`
 * (semantic_seg.py:126-164) and `Pruner.mask_dict` (pruners/Pruner.py:6-27).  The host-side mirror of
 * those interfaces (video-seg-model-compress_b200/drnb200/) binds exactly the entry points below with
 * ctypes; each one names the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success and a negative DRNB200_E_* code otherwise; it never throws.
 *     drnb200_last_error() returns a thread-local, human readable description of the last failure.
 *   - all data pointers are DEVICE pointers owned by the caller (the Python host passes
 *     torch tensors' data_ptr()); the library owns only the opaque plan handles.
 *   - every launch is enqueued on the cudaStream_t passed as `stream` (a void* so that no CUDA header is
 *     needed to bind); there are no hidden synchronisations on the forward calls.  Plan creation may
 *     synchronise (it reads the tile list back to sort the work list).
 *   - there is NO CPU fallback: without an sm_100 device the calls fail with DRNB200_E_CUDA.
 *   - 16-bit activation tensors are NHWC; `act_dtype` says whether the 16 bits are bf16 or fp16.
 */
#ifndef DRNB200_H_
#define DRNB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRNB200_VERSION 110

/* error codes */
#define DRNB200_OK          0
#define DRNB200_E_ARG      -1   /* bad argument / unsupported shape */
#define DRNB200_E_CUDA     -2   /* CUDA runtime / driver error */
#define DRNB200_E_NOMEM    -3
#define DRNB200_E_STATE    -4   /* plan used in a way it was not built for */

/* 16-bit activation / packed-weight storage type */
#define DRNB200_BF16 0
#define DRNB200_F16  1

/* conv implementation selector (drnb200_conv_desc.impl) */
#define DRNB200_IMPL_AUTO    0   /* tcgen05 when the shape allows it, else CUDA-core direct kernel */
#define DRNB200_IMPL_DIRECT  1   /* CUDA-core direct convolution (any shape; cross-check path)       */
#define DRNB200_IMPL_TCGEN05 2   /* tcgen05/TMEM/TMA implicit GEMM; fails if the shape is unsupported */

int         drnb200_version(void);
const char* drnb200_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * (a) mask -> block-sparse tile list.
 * Replaces: nothing executable in the reference at inference time (masks are only multiplied into
 * dense weights, pruners/Pruner.py:17-20).  The format follows the reference's BSR exporter
 * BlockPruner.generate_block_matrix (pruners/BlockPruner.py:344-413: indices + rowBlockPtr) and its
 * liveness rule tools/visualize_layers.py:8-13, re-ordered for an NHWC implicit GEMM:
 *   the matricised weight (O, I*kh*kw) has column  ci*kh*kw + tap  (BlockPruner.py:144);
 *   a K-block is  kb = cib * (kh*kw) + tap  with cib = ci / tile_ci;   an output tile is ot = o / tile_o;
 *   block (ot, kb) is live  <=>  any mask element in it is != 0  (HbPruner masks may exceed 1).
 * Outputs (device): row_ptr[n_ot+1] (exclusive prefix of live counts), kblk[row_ptr[n_ot]] (live kb ids,
 * ascending inside each ot), *n_live = row_ptr[n_ot].   kblk must have room for n_ot*n_cib*kh*kw ints.
 * O % tile_o == 0 and I % tile_ci == 0 are required.
 * ------------------------------------------------------------------------------------------- */
int drnb200_compact_mask(const float* mask_oihw, int O, int I, int kh, int kw,
                         int tile_o, int tile_ci,
                         int32_t* row_ptr, int32_t* kblk, int32_t* n_live, void* stream);

/* Pack the live blocks of (w * (mask != 0)) as 16-bit K-major tiles in the shared-memory image the
 * tcgen05 kernels consume: tile j (in kblk order) is tile_o rows x tile_ci elements, row pitch
 * tile_ci*2 bytes (32/64/128), 16-byte chunks XOR-swizzled exactly as TMA/UMMA SWIZZLE_{32,64,128}B
 * would lay them out (tile_ci must be 16, 32 or 64; larger Cin is split into 64-wide K-blocks by
 * choosing tile_ci = 64).  mask may be NULL (use w as is).  Replaces Pruner.apply_masks
 * (pruners/Pruner.py:17-20) + the `values` array of BlockPruner.generate_block_matrix. */
int drnb200_pack_weights(const float* w_oihw, const float* mask_oihw_or_null,
                         int O, int I, int kh, int kw, int tile_o, int tile_ci,
                         const int32_t* row_ptr, const int32_t* kblk,
                         int act_dtype, uint16_t* w_packed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (b) convolution + folded BatchNorm (+ residual) (+ ReLU).
 * Replaces nn.Conv2d -> BatchNorm2d(eval) -> [+= residual] -> ReLU as composed in
 * drn.py:49-65 (BasicBlock.forward), drn.py:86-106 (Bottleneck.forward), drn.py:201-211
 * (_make_conv_layers) and drn.py:181-186 (downsample).  padding == dilation*(ksize/2) as everywhere
 * in drn.py.   y = act( conv(x) * bn_scale[c] + bn_shift[c] (+ residual) ),  fp32 accumulation.
 * ------------------------------------------------------------------------------------------- */
typedef struct drnb200_conv_desc {
  int32_t N, H, W;          /* input batch, height, width                                        */
  int32_t Cin, Cout;
  int32_t ksize;            /* 1 or 3                                                            */
  int32_t stride;           /* 1 or 2                                                            */
  int32_t dilation;         /* >= 1                                                              */
  int32_t relu;             /* apply ReLU last                                                   */
  int32_t has_residual;     /* add `residual` (output-shaped, act_dtype NHWC) before the ReLU    */
  int32_t act_dtype;        /* DRNB200_BF16 / DRNB200_F16: x, residual, y and packed weights     */
  int32_t out_f32;          /* write y as float32 NHWC instead of act_dtype (hand-off to the head) */
  int32_t tile_o, tile_ci;  /* granularity the tile list / packed weights were built with        */
  int32_t impl;             /* DRNB200_IMPL_*                                                    */
  /* channel sub-ranges of wider NHWC tensors (0 = tightly packed).  They let one launch compute a residual
   * block's conv1 AND its 1x1 downsample as [conv1 | downsample] output channels (the downsample is the centre tap
   * of a 3x3 with the same stride; its other taps are dead K-blocks the tile list skips), and let conv2 then read
   * its input and its residual from the two halves of that tensor (drn.py:49-65, :181-186). */
  int32_t x_cpitch;         /* channels per pixel of the tensor x lives in (>= Cin); x uses channels [0, Cin) */
  int32_t res_cpitch;       /* channels per pixel of the tensor the residual lives in (>= Cout)              */
  int32_t res_coffset;      /* first channel of the residual inside that tensor                              */
  int32_t relu_n;           /* apply ReLU only to output channels < relu_n (0 = use `relu` for all channels) */
  /* Residual projection inside the K loop (0 = off).  A BasicBlock whose shortcut is a stride-1 1x1 conv + BN
   * (drn.py:181-186, layers 5/6 of DRN-D) computes relu(bn2(conv2(h)) + bn_d(proj(x))).  With proj_cin > 0 the
   * `residual` argument of drnb200_conv_forward is NOT added in the epilogue: it is a second INPUT tensor
   * [N, H, W, res_cpitch] (act_dtype, same H x W as the output) whose channels [res_coffset, res_coffset + proj_cin)
   * enter the accumulator through extra K-blocks appended to every output tile's list:
   *     kblk entry = DRNB200_KB_PROJ + 3 * cib   (cib = 64-channel block of the projection input),
   * sorted after the conv's own entries, their packed 128x64 weight tiles following in the same order.  The caller
   * folds the two BatchNorms: proj weights pre-multiplied by scale_d[c]/scale_2[c], bn_shift = shift_2 + shift_d.
   * Supported by the row-halo tcgen05 kernel only (3x3, stride 1, dilation <= 4, tile 128x64, output rows wider
   * than 128 pixels); plan creation fails with DRNB200_E_ARG otherwise and the caller keeps the projection as a
   * separate launch. */
  int32_t proj_cin;
  /* Accumulator orientation of the row-halo kernel (mode 5/6 below): 0 = the library's choice (cout-major: the
   * pixel-major flavour measured 20-45 % slower on every layer of the benchmark and is never picked on its own),
   * 1 = cout-major (TMEM lane = cout, staged shared-memory epilogue with TMA stores), 2 = pixel-major (TMEM lane =
   * pixel, register epilogue with 32-byte global accesses).  Results are identical; ignored by the other kernels. */
  int32_t acc_layout;
} drnb200_conv_desc;

#define DRNB200_KB_PROJ (3 << 20)

typedef struct drnb200_conv_plan drnb200_conv_plan;

int  drnb200_conv_plan_create(drnb200_conv_plan** out, const drnb200_conv_desc* desc,
                              const int32_t* row_ptr, const int32_t* kblk, const uint16_t* w_packed,
                              const float* bn_scale, const float* bn_shift);
int  drnb200_conv_forward(drnb200_conv_plan* plan, const void* x_nhwc, const void* residual_or_null,
                          void* y_nhwc, void* stream);
/* what the plan resolved to: 1 = direct (CUDA cores), 2 = tcgen05 */
int  drnb200_conv_plan_impl(const drnb200_conv_plan* plan);
/* which tcgen05 kernel: 0 = conv_tc MODE_T (128-cout tiles, staged epilogue, one TMA box per K-block),
 * 1 = conv_tc MODE_P (pixels as M), 2 = conv_tc MODE_T with float32 output, 3 = conv_gather (im2col in smem),
 * 4 = conv_halo (shifted windows of one halo tile), 5 = conv_tc MODE_T ROW variant (3x3 stride 1, 64-channel K-blocks,
 * rows wider than 128 pixels: each input row loaded once, the three kx taps as shifted UMMA windows; the only kernel
 * that accepts proj_cin), 6 = the same with the pixel-major accumulator (acc_layout), 7 = conv_ty (3x3 stride 1,
 * 16 -> 16 channels, even W: pixel-pair operand rows, filter rows folded into the weight operand), 8 = conv_s2 (3x3
 * stride 2, 16 -> 32 channels, even W: the input viewed as pixel pairs so that the stride is part of the operand layout,
 * no im2col copy), 9 = conv_ys (3x3 stride 1, 64 -> 64 channels (+ residual), dilation 1: input rows streamed through a
 * ring of single-row slots, filter rows folded into the weight operand), 10 = conv_y2 (3x3 stride 2, 32 -> 128 channels,
 * even W, tightly packed input: the same streaming over pixel-pair rows); -1 for direct plans */
int  drnb200_conv_plan_mode(const drnb200_conv_plan* plan);
/* live multiply-accumulates of one forward (tile-list granularity) — the numerator of tensor-pipe
 * utilisation counted at block granularity; element-granularity MACs are computed by the host. */
int64_t drnb200_conv_plan_tile_macs(const drnb200_conv_plan* plan);
void drnb200_conv_plan_destroy(drnb200_conv_plan* plan);

/* Stem: nn.Conv2d(3,C0,7,stride 1,pad 3) -> BatchNorm2d -> ReLU (drn.py:132-137).
 * x: float32 NCHW [N,3,H,W] exactly as the callers feed DRNSeg.forward (semantic_seg.py:444);
 * w: float32 OIHW [C0,3,7,7]; y: act_dtype NHWC [N,H,W,C0].  C0 must be 16.
 * Two implementations:
 *   drnb200_stem_plan_*    tcgen05, no im2col: the x direction of the 7x7 window is folded into a banded (Toeplitz)
 *                          weight matrix built once at plan creation (weights rounded to act_dtype), the y
 *                          direction is a row shift of the A operand inside one 16-bit halo tile that TMA loads
 *                          from the fp32 frame and the convert warps round to act_dtype (21 MMAs of
 *                          M=128 rows x N=128 (8 columns x 16 couts) x K=16 per 1024 pixels); this is what
 *                          DRNSeg uses.
 *   drnb200_stem_forward   CUDA-core fp32 direct convolution (no input/weight rounding); kept as the
 *                          full-precision cross-check of the tensor-core stem. */
typedef struct drnb200_stem_plan drnb200_stem_plan;
int  drnb200_stem_plan_create(drnb200_stem_plan** out, const float* w_oihw, const float* bn_scale,
                              const float* bn_shift, int N, int H, int W, int C0, int act_dtype,
                              void* stream);
int  drnb200_stem_plan_forward(drnb200_stem_plan* plan, const float* x_nchw, void* y_nhwc, void* stream);
void drnb200_stem_plan_destroy(drnb200_stem_plan* plan);
/* Frame ingest fused into the stem: frames are uint8 HWC [N,H,W,3] as cv2 / PIL deliver them
 * (seg_video_old.py:122-139); ToTensorVideoImage's `.float().div(255)` (data_transforms.py:256-281) and
 * Normalize's `(x - mean) / std` (data_transforms.py:109-125, info.json) are applied through `lut`
 * ([3][256] act_dtype values, DEVICE pointer, built by drnb200_ingest_lut and uploaded by the caller);
 * tensor channel c reads byte c of the pixel (byte 2-c when bgr != 0).  Padding pixels are 0 after
 * normalisation, exactly as Conv2d pads the normalised tensor.  Needs W % 16 == 0.  Output: same as
 * drnb200_stem_plan_forward on the normalised float32 NCHW tensor rounded to act_dtype. */
int  drnb200_stem_plan_forward_u8(drnb200_stem_plan* plan, const uint8_t* frames_nhwc, const uint16_t* lut,
                                  int bgr, void* y_nhwc, void* stream);
/* HOST function (no GPU): lut[c*256+b] = act_dtype(((float(b) / 255) - mean[c]) / std[c]), each step a
 * correctly rounded fp32 operation like the reference's torch CPU transforms.  lut: host, 768 entries. */
int  drnb200_ingest_lut(const float* mean, const float* std, int act_dtype, uint16_t* lut);
int drnb200_stem_forward(const float* x_nchw, const float* w_oihw, const float* bn_scale,
                         const float* bn_shift, int N, int H, int W, int C0, int act_dtype,
                         void* y_nhwc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (c) fused head: 1x1 classifier `seg` (semantic_seg.py:137-143) -> fixed bilinear grouped
 * ConvTranspose2d(k16,s8,p4) `up` (semantic_seg.py:115-124,147-152; zero padded borders, analytic
 * weights w[k] = 1-|2k-15|/16) -> [LogSoftmax, semantic_seg.py:139,158] -> argmax over classes with
 * first-max tie-break (torch.max(final,1), semantic_seg.py:445).
 * x: layer-8 output, act_dtype NHWC [N,h,w,C] (C % 16 == 0).  seg_w: float32 [classes, C]
 * (nn.Conv2d weight [classes,C,1,1]), seg_b: float32 [classes]; both are read once at plan creation
 * (rounded to act_dtype for the tensor-core classifier GEMM; the bias stays float32).
 * Any of the forward outputs may be NULL:
 *   labels      uint8  [N,8h,8w]            (fast path)
 *   seg_logits  float32 [N,classes,h,w]     (DRNSeg.forward()[1])
 *   logprob     float32 [N,classes,8h,8w]   (DRNSeg.forward()[0], parity mode only)
 * classes <= 32.
 * ------------------------------------------------------------------------------------------- */
typedef struct drnb200_head_plan drnb200_head_plan;

int  drnb200_head_plan_create(drnb200_head_plan** out, int N, int h, int w, int C, int classes,
                              int act_dtype, const float* seg_w, const float* seg_b, void* stream);
int  drnb200_head_forward(drnb200_head_plan* plan, const void* x_nhwc, uint8_t* labels,
                          float* seg_logits, float* logprob, void* stream);
void drnb200_head_plan_destroy(drnb200_head_plan* plan);
/* 1 when a labels-only drnb200_head_forward call (seg_logits == NULL, logprob == NULL) runs as ONE fused kernel
 * (classifier GEMM on tcgen05 -> x8 bilinear upsample -> argmax -> packed label stores; needs C % 64 == 0,
 * C <= 1024, classes <= 19); otherwise, and whenever logits or log-probs are requested, the classifier GEMM and
 * the upsample run as two launches over a float32 [N,h,w,32] scratch. */
int  drnb200_head_plan_fused(const drnb200_head_plan* plan);
/* Which `up` module the plan reproduces (semantic_seg.py:144-152):
 *   DRNB200_UP_TRANSPOSED  the fixed-bilinear grouped ConvTranspose2d(k16,s8,p4) (default, the reference's default);
 *   DRNB200_UP_ALIGNED     nn.UpsamplingBilinear2d(scale_factor=8) of `use_torch_up=True`: bilinear interpolation with
 *                          align_corners=True.  Always the two-launch form (drnb200_head_plan_fused returns 0). */
#define DRNB200_UP_TRANSPOSED 0
#define DRNB200_UP_ALIGNED    1
int  drnb200_head_plan_set_upsample(drnb200_head_plan* plan, int mode);

/* Confusion matrix accumulation: fast_hist (semantic_seg.py:293-296).
 * hist[label*classes + pred] += 1 for every pixel with 0 <= label < classes (255 = ignore falls out).
 * pred: uint8; label: uint8 if label_is_i64 == 0 else int64; hist: int64 [classes*classes], accumulates. */
int drnb200_confusion(const uint8_t* pred, const void* label, int label_is_i64, int64_t n_px,
                      int classes, int64_t* hist, void* stream);

/* uint8 labels -> int64 (the dtype torch.max returns, semantic_seg.py:445) */
/* Palette / overlay output: out[p] = palette[labels[p]] (`CITYSCAPE_PALETTE[pred]`,
 * semantic_seg.py:52-72, :101-112; seg_video.py:168); labels >= n_colors take the last palette row.
 * With frames != NULL (uint8 RGB, same pixels): out = rint(alpha*colour + (1-alpha)*frame) in separately
 * rounded fp32 steps — the alpha=0.6 overlay seg_video.py:200-203 draws with matplotlib (matplotlib is not
 * available here, so the blend rule is this library's definition, restated in oracle/frameio_oracle.py).
 * labels [n_px] uint8, palette [n_colors][3] uint8 (device), out [n_px][3] uint8; n_px % 4 == 0. */
int drnb200_colorize(const uint8_t* labels, int64_t n_px, const uint8_t* palette, int n_colors,
                     const uint8_t* frames_or_null, float alpha, uint8_t* out_rgb, void* stream);
int drnb200_labels_to_i64(const uint8_t* labels, int64_t n, int64_t* out, void* stream);

/* Frame resize, the step before the ingest in the video caller (seg_video_old.py:125-128: torchvision T.Resize on the PIL
 * frame = Pillow's 8-bit BILINEAR resampler, src/libImaging/Resample.c).  src uint8 [N,Hs,Ws,3] -> dst uint8 [N,H,W,3],
 * bit-identical to PIL: horizontal pass into the uint8 scratch `tmp` [N,Hs,W,3], then the vertical pass; per pass
 * ss = 2^21 + sum_t pixel * k[t], out = clip8(ss >> 22).  The caller passes Pillow's coefficient tables per axis as
 * int32 (normalize_coeffs_8bpc: (int)(k * 2^22 +- 0.5)): xmin[W], xcnt[W], xk[W*kx]; kx == 0 means "no horizontal pass"
 * (Ws == W), ky == 0 likewise; `tmp` may be NULL unless both passes run.  All pointers are device pointers. */
int drnb200_resize_u8(const uint8_t* src, int N, int Hs, int Ws, uint8_t* dst, int H, int W,
                      const int32_t* xmin, const int32_t* xcnt, const int32_t* xk, int kx,
                      const int32_t* ymin, const int32_t* ycnt, const int32_t* yk, int ky,
                      uint8_t* tmp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pinned HOST staging buffers for frames (H2D) and label maps (D2H) — the two PCIe ends of the path.  Replaces the
 * pageable tensors + `.cuda()` / `.cpu()` of the reference's callers (semantic_seg.py:440-455, seg_video_new.py).
 * HOST functions; the pointer is valid on every device (portable).  mode:
 *   DRNB200_HOST_PINNED  cudaHostAlloc                                   (== torch pin_memory())
 *   DRNB200_HOST_WC      write-combined: DMA reads skip CPU-cache snooping; CPU reads of it are very slow -> frames only
 *   DRNB200_HOST_HUGE    2 MiB-aligned anonymous mapping, MADV_HUGEPAGE, cudaHostRegister: fewer IOMMU entries per copy
 * drnb200_host_free accepts only pointers returned by drnb200_host_alloc (DRNB200_E_ARG otherwise).
 * ------------------------------------------------------------------------------------------- */
#define DRNB200_HOST_PINNED 0
#define DRNB200_HOST_WC     1
#define DRNB200_HOST_HUGE   2
int drnb200_host_alloc(void** out, uint64_t bytes, int mode);
int drnb200_host_free(void* p);

/* ---------------------------------------------------------------------------------------------
 * Multi-scale test.  Replaces resize_4d_tensor (semantic_seg.py:471-504: every float32 plane through
 * PIL `Image.resize((w, h), Image.BILINEAR)` on CPU threads) and the sum over scales of test_ms
 * (semantic_seg.py:540: `final = sum([resize_4d_tensor(out, w, h) for out in outputs])`):
 *   acc[N,C,H,W] = (first ? 0 : acc) + resize(src[N,C,Hs,Ws])          float32 NCHW, device pointers
 * The resampling is Pillow's (src/libImaging/Resample.c): horizontal pass into float32, then vertical pass, each
 * output sample `ss = 0.0; ss += pixel * k[t]` in double over its taps in order.  The caller passes Pillow's
 * precompute_coeffs() tables per axis: xmin[W] (first tap), xcnt[W] (taps used), xk[W*kx] (double weights, row
 * stride kx); kx == 0 means "no horizontal pass" and requires Ws == W (same for y).  Bit-identical to the reference. */
int drnb200_ms_accumulate(const float* src, int N, int C, int Hs, int Ws, float* acc, int H, int W,
                          const int32_t* xmin, const int32_t* xcnt, const double* xk, int kx,
                          const int32_t* ymin, const int32_t* ycnt, const double* yk, int ky,
                          int first, void* stream);
/* `pred = final.argmax(axis=1)` (semantic_seg.py:543; first maximum wins): acc float32 [N,C,H,W] -> uint8 [N,H,W] */
int drnb200_ms_argmax(const float* acc, int N, int C, int H, int W, uint8_t* labels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRNB200_H_ */
