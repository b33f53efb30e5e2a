// conv_plan.cu — C-ABI entry points for (b): plan creation / forward / destruction.
#include "conv_internal.cuh"
#include <new>

using namespace drnb200;

extern "C" int drnb200_conv_plan_create(drnb200_conv_plan** out, const drnb200_conv_desc* desc,
                                        const int32_t* row_ptr, const int32_t* kblk,
                                        const uint16_t* w_packed, const float* bn_scale,
                                        const float* bn_shift) {
  DRN_REQUIRE(out && desc && row_ptr && kblk && w_packed && bn_scale && bn_shift,
              "conv_plan_create: null pointer");
  const drnb200_conv_desc& d = *desc;
  DRN_REQUIRE(d.N > 0 && d.H > 0 && d.W > 0 && d.Cin > 0 && d.Cout > 0,
              "conv_plan_create: bad shape N=%d H=%d W=%d Cin=%d Cout=%d", d.N, d.H, d.W, d.Cin, d.Cout);
  DRN_REQUIRE(d.ksize == 1 || d.ksize == 3, "conv_plan_create: ksize must be 1 or 3 (got %d)", d.ksize);
  DRN_REQUIRE(d.stride >= 1 && d.dilation >= 1, "conv_plan_create: bad stride/dilation");
  DRN_REQUIRE(d.act_dtype == DRNB200_BF16 || d.act_dtype == DRNB200_F16,
              "conv_plan_create: bad act_dtype %d", d.act_dtype);
  DRN_REQUIRE(d.tile_ci == 16 || d.tile_ci == 32 || d.tile_ci == 64,
              "conv_plan_create: tile_ci must be 16/32/64 (got %d)", d.tile_ci);
  DRN_REQUIRE(d.tile_o > 0 && d.tile_o % 8 == 0 && d.Cout % d.tile_o == 0 && d.Cin % d.tile_ci == 0,
              "conv_plan_create: Cout=%d / Cin=%d not divisible by tile %dx%d", d.Cout, d.Cin,
              d.tile_o, d.tile_ci);
  DRN_REQUIRE(d.impl >= DRNB200_IMPL_AUTO && d.impl <= DRNB200_IMPL_TCGEN05,
              "conv_plan_create: bad impl %d", d.impl);
  DRN_REQUIRE(d.acc_layout >= 0 && d.acc_layout <= 2, "conv_plan_create: bad acc_layout %d", d.acc_layout);

  drnb200_conv_plan* plan = new (std::nothrow) drnb200_conv_plan();
  if (!plan) { set_error("conv_plan_create: out of host memory"); return DRNB200_E_NOMEM; }
  plan->d = d;
  plan->d_ot_order = nullptr;
  plan->tmap_ptr = nullptr;
  plan->tmap_y_ptr = nullptr;
  plan->tmap_r_ptr = nullptr;
  ConvParams& p = plan->p;
  p = ConvParams{};
  p.row_ptr = row_ptr; p.kblk = kblk; p.w_packed = reinterpret_cast<const uint8_t*>(w_packed);
  p.scale = bn_scale; p.shift = bn_shift;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.Cout = d.Cout;
  p.OH = (d.H - 1) / d.stride + 1;   // padding == dilation*(k/2)  (drn.py:27-29, :205-207)
  p.OW = (d.W - 1) / d.stride + 1;
  p.taps = d.ksize * d.ksize; p.stride = d.stride; p.dil = d.dilation;
  p.tile_o = d.tile_o; p.tile_ci = d.tile_ci;
  p.n_ot = d.Cout / d.tile_o; p.n_cib = d.Cin / d.tile_ci;
  p.relu = d.relu; p.has_res = d.has_residual; p.out_f32 = d.out_f32;
  p.x_cpitch = d.x_cpitch > 0 ? d.x_cpitch : d.Cin;
  p.res_pitch = d.res_cpitch > 0 ? d.res_cpitch : d.Cout;
  p.res_coff = d.res_coffset;
  p.relu_n = d.relu_n > 0 ? d.relu_n : (d.relu ? d.Cout : 0);
  p.proj = d.proj_cin > 0 ? d.proj_cin : 0;
  if (p.proj && (d.has_residual || d.ksize != 3 || d.stride != 1 || d.proj_cin % 64 || d.tile_o != 128 ||
                 d.tile_ci != 64 || d.res_cpitch <= 0 || d.impl == DRNB200_IMPL_DIRECT)) {
    set_error("conv_plan_create: proj_cin=%d needs a 3x3 stride-1 conv with 128x64 tiles, no epilogue residual, "
              "res_cpitch set and a tcgen05 plan", d.proj_cin);
    delete plan;
    return DRNB200_E_ARG;
  }
  const int res_width = p.proj ? p.proj : d.Cout;      // channels read from the residual / projection tensor
  if (p.x_cpitch < d.Cin || p.x_cpitch % 8 || p.res_pitch % 8 || p.res_coff % 8 || p.res_coff < 0 ||
      p.res_coff + res_width > p.res_pitch || p.relu_n > d.Cout) {
    set_error("conv_plan_create: bad channel sub-range (x_cpitch=%d res_cpitch=%d res_coffset=%d relu_n=%d)",
              d.x_cpitch, d.res_cpitch, d.res_coffset, d.relu_n);
    delete plan;
    return DRNB200_E_ARG;
  }

  plan->h_row_ptr.resize(p.n_ot + 1);
  cudaError_t e = cudaMemcpy(plan->h_row_ptr.data(), row_ptr, sizeof(int32_t) * (p.n_ot + 1),
                             cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { delete plan; return cuda_fail(e, "cudaMemcpy(row_ptr)"); }
  const int64_t n_live = plan->h_row_ptr[p.n_ot];
  plan->tile_macs = n_live * (int64_t)p.tile_o * p.tile_ci * (int64_t)p.N * p.OH * p.OW;

  plan->impl = 0;
  // narrow stride-1 3x3 layers (Cin, Cout <= 64): shifted-window halo kernel (no im2col copy);
  // 16-channel stride-2 layers: im2col-gather kernel; everything else: per-tap TMA implicit GEMM
  if (d.impl != DRNB200_IMPL_DIRECT && conv_ty_supported(d)) {
    plan->impl = DRNB200_IMPL_TCGEN05;
    plan->tc_mode = TC_MODE_TY;
  } else if (d.impl != DRNB200_IMPL_DIRECT && conv_s2_supported(d)) {
    plan->impl = DRNB200_IMPL_TCGEN05;
    plan->tc_mode = TC_MODE_S2;
  } else if (d.impl != DRNB200_IMPL_DIRECT && conv_y2_supported(d)) {
    plan->impl = DRNB200_IMPL_TCGEN05;
    plan->tc_mode = TC_MODE_Y2;
  } else if (d.impl != DRNB200_IMPL_DIRECT && conv_ys_supported(d)) {
    plan->impl = DRNB200_IMPL_TCGEN05;
    plan->tc_mode = TC_MODE_YS;
  } else if (d.impl != DRNB200_IMPL_DIRECT && conv_halo_supported(d)) {
    plan->impl = DRNB200_IMPL_TCGEN05;
    plan->tc_mode = TC_MODE_HALO;
  } else if (d.impl != DRNB200_IMPL_DIRECT && conv_gather_supported(d)) {
    plan->impl = DRNB200_IMPL_TCGEN05;
    plan->tc_mode = TC_MODE_GATHER;
  } else if (d.impl != DRNB200_IMPL_DIRECT) {
    int rc = conv_tc_setup(plan);
    if (rc == DRNB200_OK && p.proj && !(plan->tc_mode == 0 && p.row_mode)) {
      set_error("conv_plan_create: proj_cin needs the row-halo kernel (output rows wider than 128 pixels, "
                "dilation <= 4); got OW=%d dil=%d", p.OW, p.dil);
      drnb200_conv_plan_destroy(plan);
      return DRNB200_E_ARG;
    }
    if (rc == DRNB200_OK) plan->impl = DRNB200_IMPL_TCGEN05;
    else if (d.impl == DRNB200_IMPL_TCGEN05 || rc != DRNB200_E_ARG || p.proj) {
      drnb200_conv_plan_destroy(plan);
      return rc;
    }
  }
  if (plan->impl == 0) {
    if (d.Cout % 8 != 0) {
      set_error("conv_plan_create: direct kernel needs Cout %% 8 == 0 (got %d)", d.Cout);
      drnb200_conv_plan_destroy(plan);
      return DRNB200_E_ARG;
    }
    plan->impl = DRNB200_IMPL_DIRECT;
  }
  *out = plan;
  return DRNB200_OK;
}

extern "C" int drnb200_conv_forward(drnb200_conv_plan* plan, const void* x_nhwc,
                                    const void* residual_or_null, void* y_nhwc, void* stream) {
  DRN_REQUIRE(plan && x_nhwc && y_nhwc, "conv_forward: null pointer");
  const bool wants_res = plan->d.has_residual || plan->p.proj;
  if (wants_res && !residual_or_null) {
    set_error("conv_forward: plan was built with has_residual / proj_cin but residual is NULL");
    return DRNB200_E_STATE;
  }
  plan->p.x = x_nhwc;
  plan->p.residual = wants_res ? residual_or_null : nullptr;
  plan->p.y = y_nhwc;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->impl != DRNB200_IMPL_TCGEN05) return conv_direct_launch(plan, st);
  if (plan->tc_mode == TC_MODE_TY) return conv_ty_launch(plan, st);
  if (plan->tc_mode == TC_MODE_S2) return conv_s2_launch(plan, st);
  if (plan->tc_mode == TC_MODE_YS) return conv_ys_launch(plan, st);
  if (plan->tc_mode == TC_MODE_Y2) return conv_y2_launch(plan, st);
  if (plan->tc_mode == TC_MODE_HALO) return conv_halo_launch(plan, st);
  return plan->tc_mode == TC_MODE_GATHER ? conv_gather_launch(plan, st) : conv_tc_launch(plan, st);
}

extern "C" int drnb200_conv_plan_impl(const drnb200_conv_plan* plan) { return plan ? plan->impl : 0; }

extern "C" int drnb200_conv_plan_mode(const drnb200_conv_plan* plan) {
  if (!plan || plan->impl != DRNB200_IMPL_TCGEN05) return -1;
  return (plan->tc_mode == 0 && plan->p.row_mode) ? (plan->p.pix_mode ? 6 : 5) : plan->tc_mode;
}

extern "C" int64_t drnb200_conv_plan_tile_macs(const drnb200_conv_plan* plan) {
  return plan ? plan->tile_macs : 0;
}

extern "C" void drnb200_conv_plan_destroy(drnb200_conv_plan* plan) {
  if (!plan) return;
  if (plan->d_ot_order) cudaFree(plan->d_ot_order);
  delete plan;
}
