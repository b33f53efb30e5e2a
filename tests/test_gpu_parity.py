"""Parity of the CUDA path (through the C ABI / DRNSeg mirror) against the oracle and the golden fixtures.

Gates (BASELINE.json north_star): mask compaction bit-exact; label-map argmax agreement >= 99.9 % of pixels;
logits within 2e-2 relative; mIoU within 0.1 point.  Run on the GPU box: pytest -m gpu
"""
import collections
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import drnb200
from drnb200 import ffi
from conftest import ROOT
from helpers import fixture_frames, fixture_state_dict, gate_case_cpu, golden, load_keys
from oracle import compact_oracle, drn_oracle, recipe

pytestmark = pytest.mark.gpu

LOGIT_RTOL = 2e-2        # north_star: "logits within 2e-2 relative" (relative to the logit range)
LABEL_AGREE = 0.999      # north_star: ">= 99.9 % of pixels" — asserted on ALL pixels for fp16 storage (the default)
MIOU_TOL = 0.1           # north_star: "mIoU within 0.1 point"

PARITY_LOG = os.path.join(ROOT, "gpurun_out", "parity_table.jsonl")


def record(case, **values):
    """append one measured parity row (pytest -q hides prints): gpurun_out/parity_table.jsonl on the GPU box, summarised
    into the tracked profiles/r02_parity_table.txt by tools/parity_table.py"""
    os.makedirs(os.path.dirname(PARITY_LOG), exist_ok=True)
    with open(PARITY_LOG, "a") as fh:
        fh.write(json.dumps(dict(case=case, **{k: (float(v) if isinstance(v, (float, np.floating)) else v)
                                              for k, v in values.items()})) + "\n")


def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def rel_err(got, ref):
    ref = ref.double()
    return float((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


# ------------------------------------------------------------------------------------------------ (a)
def _compact_on_gpu(mask, tile_o, tile_ci):
    lib = ffi.lib()
    O, I, kh, kw = mask.shape
    m = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.float32)).to(dev())
    n_ot, n_kb = O // tile_o, (I // tile_ci) * kh * kw
    rp = torch.full((n_ot + 1,), -1, dtype=torch.int32, device=dev())
    kb = torch.full((max(1, n_ot * n_kb),), -1, dtype=torch.int32, device=dev())
    nl = torch.zeros(1, dtype=torch.int32, device=dev())
    ffi.check(lib.drnb200_compact_mask(ffi.ptr(m), O, I, kh, kw, tile_o, tile_ci, ffi.ptr(rp), ffi.ptr(kb),
                                       ffi.ptr(nl), ffi.stream_ptr()))
    torch.cuda.synchronize()
    n = int(nl.item())
    return rp.cpu().numpy(), kb.cpu().numpy()[:n], rp, kb


def _masks_for_compaction():
    fx = np.load(golden("pruner_masks.npz"))
    out = []
    for name, shape in (("block_a_uncollapsed", (64, 32, 3, 3)), ("block_a_collapsed", (64, 32, 3, 3)),
                        ("block_a_unstructured", (64, 32, 3, 3)), ("hb_a", (64, 32, 3, 3)),
                        ("rmb_a", (64, 32, 3, 3)), ("rmcdb_a", (64, 32, 3, 3)), ("group_a", (64, 32, 3, 3)),
                        ("block_b_1x1conv", (32, 64, 1, 1)), ("srmb_RAMANUJAN_1", (64, 32, 3, 3)),
                        ("srmb_optimal_entry10", (256, 256, 3, 3))):
        out.append((name, recipe.unpack_mask_bits(fx[name], shape)))
    rng = np.random.RandomState(0)
    out.append(("empty", np.zeros((32, 32, 3, 3), np.float32)))
    out.append(("full", np.ones((32, 32, 3, 3), np.float32)))
    lone = np.zeros((128, 128, 3, 3), np.float32)
    lone[77, 99, 2, 0] = 3.0                      # one surviving element, value != 1 (Hb-style)
    out.append(("lone", lone))
    out.append(("random_blocks", np.kron((rng.rand(4, 8) < 0.25), np.ones((128, 64))).reshape(512, 512, 1, 1)
                .astype(np.float32).repeat(9, axis=2).reshape(512, 512, 3, 3)))
    return out


@pytest.mark.parametrize("tiles", [(16, 16), (32, 32), (8, 16), (128, 64)])
def test_compaction_bit_exact(tiles):
    tile_o, tile_ci = tiles
    ran = 0
    for name, mask in _masks_for_compaction():
        O, I = mask.shape[:2]
        if O % tile_o or I % tile_ci:
            continue
        rp_ref, kb_ref = compact_oracle.compact_mask(mask, tile_o, tile_ci)
        rp, kb, _, _ = _compact_on_gpu(mask, tile_o, tile_ci)
        assert np.array_equal(rp, rp_ref), name
        assert np.array_equal(kb, kb_ref), name
        ran += 1
    assert ran >= 3


@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
@pytest.mark.parametrize("tiles", [(16, 16), (32, 32), (64, 64), (128, 64)])
def test_weight_packing_bit_exact(act, tiles):
    tile_o, tile_ci = tiles
    lib = ffi.lib()
    rng = np.random.RandomState(1)
    O, I = 256, 128
    for k in (1, 3):
        w = (rng.randn(O, I, k, k) * 0.05).astype(np.float32)
        mask = np.kron(rng.rand(O // tile_o, I // tile_ci) < 0.5, np.ones((tile_o, tile_ci)))[:, :, None, None] \
            * np.ones((1, 1, k, k))
        mask = mask.astype(np.float32)
        mask[0, 0, 0, 0] = 1.0
        rp_ref, kb_ref = compact_oracle.compact_mask(mask, tile_o, tile_ci)
        _, _, rp, kb = _compact_on_gpu(mask, tile_o, tile_ci)
        packed = torch.zeros(len(kb_ref) * tile_o * tile_ci, dtype=torch.int16, device=dev())
        wd, md = torch.from_numpy(w).to(dev()), torch.from_numpy(mask).to(dev())
        ffi.check(lib.drnb200_pack_weights(ffi.ptr(wd), ffi.ptr(md), O, I, k, k, tile_o, tile_ci, ffi.ptr(rp),
                                           ffi.ptr(kb), act, ffi.ptr(packed), ffi.stream_ptr()))
        torch.cuda.synchronize()
        ref = compact_oracle.pack_weights(w, mask, tile_o, tile_ci, rp_ref, kb_ref, act)
        assert np.array_equal(packed.cpu().numpy().view(np.uint16), ref)


# ------------------------------------------------------------------------------------------------ (b)
def _conv_case(N, H, W, cin, cout, k, stride, dil, relu, res, act, impl, density, seed, out_f32=True, acc_layout=0,
               expect_mode=None, live_taps=None):
    """one conv+BN(+res)(+ReLU) through the C ABI vs torch fp32 on the same 16-bit-representable operands"""
    lib = ffi.lib()
    g = torch.Generator().manual_seed(seed)
    tdt = torch.bfloat16 if act == ffi.BF16 else torch.float16
    tile_ci = 64 if cin % 64 == 0 else 32 if cin % 32 == 0 else 16
    tile_o = 128 if cout % 128 == 0 else cout
    x = torch.randn(N, cin, H, W, generator=g).to(tdt)
    w = recipe.round_bf16(torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5)
    blocks = (torch.rand(cout // tile_o, cin // tile_ci, generator=g) < density).float()
    mask = torch.kron(blocks, torch.ones(tile_o, tile_ci))[:, :, None, None].expand(-1, -1, k, k).contiguous()
    if live_taps is not None:                     # whole filter taps pruned (the tile list is per (cin block, ky, kx))
        keep = torch.zeros(k, k)
        for ky, kx in live_taps:
            keep[ky, kx] = 1.0
        mask = mask * keep
    w = w * mask
    scale = 0.5 + torch.rand(cout, generator=g)
    shift = 0.2 * torch.randn(cout, generator=g)
    OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
    r = torch.randn(N, cout, OH, OW, generator=g).to(tdt) if res else None
    # fp32 reference (weights as the kernel sees them: rounded to the activation dtype)
    ref = torch.nn.functional.conv2d(x.float(), w.to(tdt).float(), None, stride, dil * (k // 2), dil)
    ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    d = dev()
    xd = x.permute(0, 2, 3, 1).contiguous().to(d)
    rd = r.permute(0, 2, 3, 1).contiguous().to(d) if res else None
    wd, md = w.to(d), mask.to(d)
    n_ot, n_kb = cout // tile_o, (cin // tile_ci) * k * k
    rp = torch.empty(n_ot + 1, dtype=torch.int32, device=d)
    kb = torch.empty(n_ot * n_kb, dtype=torch.int32, device=d)
    nl = torch.zeros(1, dtype=torch.int32, device=d)
    st = ffi.stream_ptr()
    ffi.check(lib.drnb200_compact_mask(ffi.ptr(md), cout, cin, k, k, tile_o, tile_ci, ffi.ptr(rp), ffi.ptr(kb),
                                       ffi.ptr(nl), st))
    packed = torch.empty(max(1, int(nl.item())) * tile_o * tile_ci, dtype=torch.int16, device=d)
    ffi.check(lib.drnb200_pack_weights(ffi.ptr(wd), ffi.ptr(md), cout, cin, k, k, tile_o, tile_ci, ffi.ptr(rp),
                                       ffi.ptr(kb), act, ffi.ptr(packed), st))
    desc = ffi.ConvDesc(N=N, H=H, W=W, Cin=cin, Cout=cout, ksize=k, stride=stride, dilation=dil, relu=int(relu),
                        has_residual=int(res), act_dtype=act, out_f32=int(out_f32), tile_o=tile_o, tile_ci=tile_ci,
                        impl=impl, acc_layout=acc_layout)
    plan = C.c_void_p()
    sc, sh = scale.to(d), shift.to(d)
    ffi.check(lib.drnb200_conv_plan_create(C.byref(plan), C.byref(desc), ffi.ptr(rp), ffi.ptr(kb), ffi.ptr(packed),
                                           ffi.ptr(sc), ffi.ptr(sh)))
    assert lib.drnb200_conv_plan_impl(plan) == impl
    if expect_mode is not None:
        assert lib.drnb200_conv_plan_mode(plan) == expect_mode
    # guard bands around the output: the kernels write nothing outside their tensor
    G = 4096
    flat = torch.full((N * OH * OW * cout + 2 * G,), float("nan"), dtype=torch.float32 if out_f32 else tdt, device=d)
    y = flat[G:G + N * OH * OW * cout].view(N, OH, OW, cout)
    ffi.check(lib.drnb200_conv_forward(plan, ffi.ptr(xd), ffi.ptr(rd), ffi.ptr(y), st))
    torch.cuda.synchronize()
    lib.drnb200_conv_plan_destroy(plan)
    assert bool(torch.isnan(flat[:G]).all()) and bool(torch.isnan(flat[-G:]).all())
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    tol = 2e-3 if out_f32 else (2.0 ** -8 if act == ffi.BF16 else 2.0 ** -10)      # 16-bit output: one rounding
    assert err <= tol * max(1.0, ref.abs().max().item()), "max abs err %g" % err
    return got


CONV_CASES = [
    # N  H   W   cin cout k  s  d  relu res  density
    (2, 24, 40, 16, 16, 3, 1, 1, True, False, 1.0),      # layer1-like   (MODE_P, SWIZZLE_32B)
    (1, 33, 47, 16, 32, 3, 2, 1, True, False, 1.0),      # layer2-like   (stride 2, odd size)
    (1, 32, 64, 32, 64, 3, 2, 1, True, False, 1.0),      # layer3.0.conv1 (SWIZZLE_64B)
    (1, 32, 64, 32, 64, 1, 2, 1, False, False, 1.0),     # layer3.0.downsample
    (2, 16, 32, 64, 64, 3, 1, 1, True, True, 0.5),       # layer3.x.conv2 + residual
    (1, 32, 64, 64, 128, 3, 2, 1, True, False, 1.0),     # layer4.0.conv1 (MODE_T, stride 2)
    (1, 16, 32, 128, 256, 3, 1, 2, True, False, 0.5),    # layer5.0.conv1 (dilation 2)
    (1, 16, 32, 256, 256, 1, 1, 1, False, False, 0.5),   # 1x1 projection
    (2, 16, 32, 256, 512, 3, 1, 4, True, True, 0.25),    # layer6 (dilation 4, residual, 75 % sparse)
    (1, 8, 16, 512, 512, 3, 1, 1, True, False, 0.0),     # everything pruned: y = relu(shift)
    (1, 24, 24, 2048, 512, 3, 1, 2, True, False, 0.1),   # D-54 layer7: K = 18432
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
def test_conv_tcgen05_vs_fp32(case, act):
    N, H, W, cin, cout, k, s, d, relu, res, dens = case
    _conv_case(N, H, W, cin, cout, k, s, d, relu, res, act, ffi.IMPL_TCGEN05, dens, seed=hash(case) & 0xFFFF)


ROW_CASES = [
    # N  H   W   cin cout d  relu  res   density      3x3 stride-1 convs over 64-channel K-blocks, rows of > 128 pixels
    (1, 5, 300, 128, 256, 2, True, True, 0.5),        # ragged second row tile (300 = 256 + 44), dilation 2, residual
    (2, 3, 257, 64, 128, 1, True, False, 1.0),        # one pixel in the second tile
    (1, 2, 512, 256, 256, 4, True, True, 0.25),       # dilation 4, 75 % sparse, two full tiles
    (1, 3, 256, 128, 512, 1, False, True, 0.5),       # no ReLU, four cout tiles
    (1, 2, 264, 128, 128, 2, True, False, 0.0),       # everything pruned: y = relu(shift)
]


@pytest.mark.parametrize("case", ROW_CASES)
@pytest.mark.parametrize("layout", [1, 2])
@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
def test_conv_row_kernel_both_accumulator_layouts(case, layout, act):
    """the row-halo kernel with the cout-major accumulator + staged epilogue (mode 5) and with the pixel-major
    accumulator + register epilogue (mode 6, drnb200_conv_desc.acc_layout) against torch fp32, and against each other:
    both accumulate the same products in fp32 and round once, so their outputs must be bit-identical"""
    N, H, W, cin, cout, dil, relu, res, dens = case
    outs = []
    for lay in (layout, 3 - layout):
        outs.append(_conv_case(N, H, W, cin, cout, 3, 1, dil, relu, res, act, ffi.IMPL_TCGEN05, dens,
                               seed=W + cin + cout, out_f32=False, acc_layout=lay, expect_mode=4 + lay))
    assert torch.equal(outs[0], outs[1])


TY_CASES = [
    # N  H    W   relu  live taps (None = all nine)          3x3 stride-1 16 -> 16 (DRN layer1), 16-bit output
    (2, 24, 300, True, None),                                # two row tiles, the second one ragged (300 = 256 + 44)
    (1, 19, 256, True, None),                                # H not a multiple of the 8-row tile, one full row tile
    (1, 11, 770, True, None),                                # four row tiles, one pixel pair in the last
    (1, 8, 40, False, None),                                 # narrower than a tile, no ReLU
    (3, 9, 258, True, None),                                 # one pixel pair / one row in the second tiles
    (1, 37, 260, True, [(0, 0), (1, 1), (2, 2), (0, 2)]),    # pruned taps (zero slots of the folded weight stack)
    (1, 16, 136, True, [(2, 1)]),                            # a single live tap, and not the ky = 0 one that
                                                             # initialises the accumulator columns
    (2, 264, 640, True, None),                               # 330 tiles: two or three per CTA
    (4, 400, 512, True, None),                               # 800 tiles: the halo ring and the accumulators wrap
]


@pytest.mark.parametrize("case", TY_CASES)
@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
def test_conv_ty_layer1_kernel(case, act):
    """conv_ty (plan mode 7: pixel-pair operand rows, filter rows folded into the weight operand, 8 output rows x 2
    pixels x 16 couts as accumulator columns) against torch fp32 on the same 16-bit operands, through the C ABI; guard
    bands checked by _conv_case"""
    N, H, W, relu, taps = case
    _conv_case(N, H, W, 16, 16, 3, 1, 1, relu, False, act, ffi.IMPL_TCGEN05, 1.0, seed=H * W, out_f32=False,
               expect_mode=7, live_taps=taps)


def test_conv_ty_and_s2_leave_odd_widths_to_the_older_kernels():
    """the pixel-pair view needs an even W: odd widths run on conv_halo (mode 4) / conv_gather (mode 3) as before"""
    _conv_case(1, 9, 129, 16, 16, 3, 1, 1, True, False, ffi.F16, ffi.IMPL_TCGEN05, 1.0, seed=3, out_f32=False, expect_mode=4)
    _conv_case(1, 9, 129, 16, 32, 3, 2, 1, True, False, ffi.F16, ffi.IMPL_TCGEN05, 1.0, seed=4, out_f32=False, expect_mode=3)


S2_CASES = [
    # N  H    W   relu  live taps (None = all nine)          3x3 stride-2 16 -> 32 (DRN layer2), 16-bit output, even W
    (2, 24, 300, True, None),                                # 150 output pixels: a ragged second row tile
    (1, 19, 256, True, None),                                # odd H (10 output rows: 2.5 row tiles), one full tile
    (1, 8, 40, False, None),                                 # narrower than a tile, no ReLU
    (3, 18, 258, True, None),                                # one output pixel / one output row in the second tiles
    (1, 37, 520, True, [(0, 0), (1, 1), (2, 2), (0, 2)]),    # pruned taps (zero slots of the weight stacks)
    (1, 16, 272, True, [(2, 0)]),                            # a single live tap, not one of the accumulate-off ones
    (1, 16, 272, True, [(1, 0)]),                            # only the tap whose MMA initialises the columns
    (2, 528, 1280, True, None),                              # 1320 tiles: slots, accumulators and barriers wrap
]


@pytest.mark.parametrize("case", S2_CASES)
@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
def test_conv_s2_layer2_kernel(case, act):
    """conv_s2 (plan mode 8: stride 2 through pixel-pair operand rows, filter rows folded into the weight operand)
    against torch fp32 on the same 16-bit operands, through the C ABI; guard bands checked by _conv_case"""
    N, H, W, relu, taps = case
    _conv_case(N, H, W, 16, 32, 3, 2, 1, relu, False, act, ffi.IMPL_TCGEN05, 1.0, seed=H * W + 1, out_f32=False,
               expect_mode=8, live_taps=taps)


YS_CASES = [
    # N  H    W   relu  res    live taps (None = all nine)      3x3 stride-1 64 -> 64 (DRN layer3 blocks), 16-bit output
    (2, 24, 300, True, True, None),                           # three row tiles, the last one ragged (300 = 256 + 44)
    (1, 19, 128, True, False, None),                          # H not a multiple of the 8-row segment
    (1, 8, 40, False, True, None),                            # narrower than a tile, residual without ReLU
    (3, 9, 129, True, True, None),                            # one pixel / one row in the second tiles
    (1, 37, 260, True, True, [(0, 0), (1, 1), (2, 2), (0, 2)]),   # pruned taps (zero blocks of the weight stacks)
    (1, 16, 136, True, False, [(2, 1)]),                      # a single live tap, not the accumulate-off one
    (2, 136, 520, True, True, None),                          # 170 items: more than one per CTA, rings and slots wrap
    (4, 264, 512, True, False, None),                         # 528 items
]


@pytest.mark.parametrize("case", YS_CASES)
@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
def test_conv_ys_layer3_kernel(case, act):
    """conv_ys (plan mode 9: input rows streamed through single-row slots, filter rows folded into the weight
    operand, one TMEM slot per output row of the 8-row segment) against torch fp32 on the same 16-bit operands,
    through the C ABI; guard bands checked by _conv_case"""
    N, H, W, relu, res, taps = case
    _conv_case(N, H, W, 64, 64, 3, 1, 1, relu, res, act, ffi.IMPL_TCGEN05, 1.0, seed=H * W + 2, out_f32=False,
               expect_mode=9, live_taps=taps)


Y2_CASES = [
    # N  H    W   relu  live taps (None = all nine)          3x3 stride-2 32 -> 128 (DRN block 3.0 conv1 + downsample)
    (2, 24, 300, True, None),                                # 150 output pixels: a ragged second row tile
    (1, 19, 256, True, None),                                # odd H (10 output rows: 2.5 segments), one full tile
    (1, 8, 40, False, None),                                 # narrower than a tile, no ReLU
    (3, 18, 258, True, None),                                # one output pixel / one output row in the second tiles
    (1, 37, 520, True, [(0, 0), (1, 1), (2, 2), (0, 2)]),    # pruned taps (zero blocks of the weight stacks)
    (1, 16, 272, True, [(2, 0)]),                            # a single live tap, not one of the accumulate-off ones
    (1, 16, 272, True, [(1, 0)]),                            # only the tap whose MMA initialises the columns
    (2, 528, 1280, True, None),                              # 1320 items: ring, slots and barriers wrap
]


@pytest.mark.parametrize("case", Y2_CASES)
@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
def test_conv_y2_block30_kernel(case, act):
    """conv_y2 (plan mode 10: stride 2 through streamed pixel-pair rows, filter rows folded into the weight operand,
    staged TMA stores) against torch fp32 on the same 16-bit operands, through the C ABI"""
    N, H, W, relu, taps = case
    _conv_case(N, H, W, 32, 128, 3, 2, 1, relu, False, act, ffi.IMPL_TCGEN05, 1.0, seed=H * W + 3, out_f32=False,
               expect_mode=10, live_taps=taps)


def test_conv_ty_many_launches_of_hbm_sized_batches():
    """regression: with one barrier per halo slot, an MMA warp could ask for the NEXT fill of a slot whose current
    fill was still in flight (TMA boxes complete out of order once the batch no longer fits in L2) and
    mbarrier.try_wait.parity answered "done": 1-2 % of the launches of two 1024x2048 frames faulted.  300 launches of
    one plan over two alternating buffer pairs (the tensor map is re-encoded at every switch) must all give the
    first launch's bits."""
    lib = ffi.lib()
    d, N, H, W = dev(), 2, 1024, 2048
    g = torch.Generator().manual_seed(7)
    wd = (torch.randn(16, 16, 3, 3, generator=g) * 0.1).to(d)
    md = torch.ones_like(wd)
    rp = torch.empty(2, dtype=torch.int32, device=d)
    kb = torch.empty(9, dtype=torch.int32, device=d)
    nl = torch.zeros(1, dtype=torch.int32, device=d)
    st = ffi.stream_ptr()
    ffi.check(lib.drnb200_compact_mask(ffi.ptr(md), 16, 16, 3, 3, 16, 16, ffi.ptr(rp), ffi.ptr(kb), ffi.ptr(nl), st))
    packed = torch.empty(9 * 256, dtype=torch.int16, device=d)
    ffi.check(lib.drnb200_pack_weights(ffi.ptr(wd), ffi.ptr(md), 16, 16, 3, 3, 16, 16, ffi.ptr(rp), ffi.ptr(kb),
                                       ffi.F16, ffi.ptr(packed), st))
    desc = ffi.ConvDesc(N=N, H=H, W=W, Cin=16, Cout=16, ksize=3, stride=1, dilation=1, relu=1, has_residual=0,
                        act_dtype=ffi.F16, out_f32=0, tile_o=16, tile_ci=16, impl=ffi.IMPL_TCGEN05, acc_layout=0)
    plan = C.c_void_p()
    sc, sh = torch.ones(16, device=d), torch.zeros(16, device=d)
    ffi.check(lib.drnb200_conv_plan_create(C.byref(plan), C.byref(desc), ffi.ptr(rp), ffi.ptr(kb), ffi.ptr(packed),
                                           ffi.ptr(sc), ffi.ptr(sh)))
    assert lib.drnb200_conv_plan_mode(plan) == 7
    xs = [torch.randn(N, H, W, 16, device=d, generator=torch.Generator(d).manual_seed(k)).half() for k in range(2)]
    ys = [torch.empty(N, H, W, 16, device=d, dtype=torch.half) for _ in range(2)]
    first = []
    for k in range(2):
        ffi.check(lib.drnb200_conv_forward(plan, ffi.ptr(xs[k]), None, ffi.ptr(ys[k]), st))
        first.append(ys[k].clone())
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(xs[0][:1, :64].permute(0, 3, 1, 2).float(), wd.half().float(), None, 1, 1).relu()
    assert (first[0][:1, :63].permute(0, 3, 1, 2).float() - ref[:, :, :63]).abs().max().item() <= 2.0 ** -9 * max(1.0, ref.max().item())
    for r in range(300):
        k = r % 2
        ys[k].zero_()
        ffi.check(lib.drnb200_conv_forward(plan, ffi.ptr(xs[k]), None, ffi.ptr(ys[k]), st))
        if r % 50 >= 48:
            assert torch.equal(ys[k], first[k]), r
    torch.cuda.synchronize()
    lib.drnb200_conv_plan_destroy(plan)


@pytest.mark.parametrize("case", CONV_CASES[:9])
def test_conv_direct_vs_fp32(case):
    N, H, W, cin, cout, k, s, d, relu, res, dens = case
    _conv_case(N, H, W, cin, cout, k, s, d, relu, res, ffi.BF16, ffi.IMPL_DIRECT, dens, seed=7)


@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
@pytest.mark.parametrize("layout", [1, 2])
@pytest.mark.parametrize("geom", [(1, 5, 300, 128, 256, 64, 64, 0, 2), (2, 3, 257, 64, 128, 128, 192, 64, 4),
                                  (1, 2, 512, 256, 256, 128, 128, 0, 1)])
def test_conv_projection_k_blocks_vs_fp32(geom, layout, act):
    """drnb200_conv_desc.proj_cin through the C ABI: 3x3 conv over h plus 1x1 projection of a second tensor x inside
    the same K loop (DRNB200_KB_PROJ entries), ragged row tiles, tiles with only-conv / only-projection / no entries,
    projection input as a channel sub-range; vs torch fp32 on the same 16-bit-representable operands"""
    N, H, W, cin, cout, pcin, ppitch, poff, dil = geom
    lib, d, st = ffi.lib(), dev(), ffi.stream_ptr()
    g = torch.Generator().manual_seed(W + cin)
    tdt = torch.bfloat16 if act == ffi.BF16 else torch.float16
    n_ot = cout // 128
    h = torch.randn(N, cin, H, W, generator=g).to(tdt)
    xfull = torch.randn(N, ppitch, H, W, generator=g).to(tdt)
    w = recipe.round_bf16(torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5)
    wp = recipe.round_bf16(torch.randn(cout, pcin, 1, 1, generator=g) * (2.0 / pcin) ** 0.5)
    blocks = (torch.rand(n_ot, cin // 64, generator=g) < 0.6).float()
    pblocks = (torch.rand(n_ot, pcin // 64, generator=g) < 0.7).float()
    blocks[0] = 0                                  # output tile 0: projection entries only
    pblocks[0, 0] = 1
    if n_ot > 1:
        pblocks[1] = 0                             # output tile 1: conv entries only
        blocks[1, 0] = 1
    mask = torch.kron(blocks, torch.ones(128, 64))[:, :, None, None].expand(-1, -1, 3, 3).contiguous()
    pmask = torch.kron(pblocks, torch.ones(128, 64))[:, :, None, None].contiguous()
    w, wp = w * mask, wp * pmask
    scale = 0.5 + torch.rand(cout, generator=g)
    shift = 0.2 * torch.randn(cout, generator=g)
    xs = xfull[:, poff:poff + pcin]
    ref = torch.nn.functional.conv2d(h.float(), w.to(tdt).float(), None, 1, dil, dil) \
        + torch.nn.functional.conv2d(xs.float(), wp.to(tdt).float())
    ref = (ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).relu()

    def compact_pack(wt, mk, k):
        O, I = wt.shape[:2]
        wd_, md_ = wt.contiguous().to(d), mk.contiguous().to(d)
        rp = torch.empty(O // 128 + 1, dtype=torch.int32, device=d)
        kb = torch.empty(max(1, (O // 128) * (I // 64) * k * k), dtype=torch.int32, device=d)
        nl = torch.zeros(1, dtype=torch.int32, device=d)
        ffi.check(lib.drnb200_compact_mask(ffi.ptr(md_), O, I, k, k, 128, 64, ffi.ptr(rp), ffi.ptr(kb), ffi.ptr(nl), st))
        n = int(nl.item())
        pk = torch.empty((max(1, n), 128 * 64), dtype=torch.int16, device=d)
        ffi.check(lib.drnb200_pack_weights(ffi.ptr(wd_), ffi.ptr(md_), O, I, k, k, 128, 64, ffi.ptr(rp), ffi.ptr(kb),
                                           act, ffi.ptr(pk), st))
        return rp.cpu().tolist(), kb[:n], pk[:n]

    rp2, kb2, pk2 = compact_pack(w, mask, 3)
    rpd, kbd, pkd = compact_pack(wp, pmask, 1)
    kbd = kbd * 3 + ffi.KB_PROJ
    kparts, wparts, row_ptr = [], [], [0]
    for ot in range(n_ot):
        kparts += [kb2[rp2[ot]:rp2[ot + 1]], kbd[rpd[ot]:rpd[ot + 1]]]
        wparts += [pk2[rp2[ot]:rp2[ot + 1]], pkd[rpd[ot]:rpd[ot + 1]]]
        row_ptr.append(row_ptr[-1] + rp2[ot + 1] - rp2[ot] + rpd[ot + 1] - rpd[ot])
    kblk = torch.cat(kparts + [torch.zeros(1, dtype=torch.int32, device=d)]).contiguous()
    packed = torch.cat(wparts + [torch.zeros((1, 128 * 64), dtype=torch.int16, device=d)]).contiguous()
    rp = torch.tensor(row_ptr, dtype=torch.int32, device=d)
    desc = ffi.ConvDesc(N=N, H=H, W=W, Cin=cin, Cout=cout, ksize=3, stride=1, dilation=dil, relu=1, has_residual=0,
                        act_dtype=act, out_f32=0, tile_o=128, tile_ci=64, impl=ffi.IMPL_TCGEN05, res_cpitch=ppitch,
                        res_coffset=poff, proj_cin=pcin, acc_layout=layout)
    plan = C.c_void_p()
    sc, sh = scale.to(d), shift.to(d)
    ffi.check(lib.drnb200_conv_plan_create(C.byref(plan), C.byref(desc), ffi.ptr(rp), ffi.ptr(kblk), ffi.ptr(packed),
                                           ffi.ptr(sc), ffi.ptr(sh)))
    assert lib.drnb200_conv_plan_mode(plan) == 4 + layout              # row-halo kernel, either accumulator layout
    hd = h.permute(0, 2, 3, 1).contiguous().to(d)
    xd = xfull.permute(0, 2, 3, 1).contiguous().to(d)
    y = torch.full((N, H, W, cout), float("nan"), dtype=tdt, device=d)
    with pytest.raises(ffi.Drnb200Error):
        ffi.check(lib.drnb200_conv_forward(plan, ffi.ptr(hd), None, ffi.ptr(y), st))     # projection input missing
    ffi.check(lib.drnb200_conv_forward(plan, ffi.ptr(hd), ffi.ptr(xd), ffi.ptr(y), st))
    torch.cuda.synchronize()
    lib.drnb200_conv_plan_destroy(plan)
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert torch.isfinite(got).all()
    tol = (2.0 ** -8 if act == ffi.BF16 else 2.0 ** -10) * max(1.0, ref.abs().max().item())
    assert (got - ref).abs().max().item() <= tol
    # unsupported geometries are refused (the caller keeps the projection as its own launch)
    bad = ffi.ConvDesc(N=N, H=H, W=64, Cin=cin, Cout=cout, ksize=3, stride=1, dilation=dil, relu=1, has_residual=0,
                       act_dtype=act, out_f32=0, tile_o=128, tile_ci=64, impl=ffi.IMPL_TCGEN05, res_cpitch=ppitch,
                       res_coffset=poff, proj_cin=pcin)
    assert lib.drnb200_conv_plan_create(C.byref(plan), C.byref(bad), ffi.ptr(rp), ffi.ptr(kblk), ffi.ptr(packed),
                                        ffi.ptr(sc), ffi.ptr(sh)) != 0


# ------------------------------------------------------------------------------------------ end to end
def _build(arch, sd, masks, act):
    m = drnb200.DRNSeg(arch, 19, pretrained_model=None, pretrained=False, act_dtype=act)
    missing = m.load_state_dict(sd, strict=False)
    assert set(missing.missing_keys) <= {"up.weight"} and not missing.unexpected_keys
    m = m.to(dev()).eval()
    if masks:
        m.set_masks(masks)
    return m


E2E = [("fwd_drn_d_22_64x128_dense.npz", "drn_d_22"), ("fwd_drn_d_22_64x128_block75.npz", "drn_d_22"),
       ("fwd_drn_d_38_32x64_block75.npz", "drn_d_38"), ("fwd_drn_d_54_32x64_dense.npz", "drn_d_54"),
       ("fwd_drn_c_26_32x64_dense.npz", "drn_c_26")]


@pytest.mark.parametrize("name,arch", E2E)
@pytest.mark.parametrize("act", ["fp16", "bf16"])
def test_forward_against_golden_fixture(name, arch, act):
    """small frames: CUDA path vs the outputs the REAL reference produced (tests/golden).  These maps hold 2048-8192
    pixels, so ONE flipped pixel is 0.012-0.05 %: the label gate is asserted as "at most max(2, 0.1 %) pixels differ"
    for fp16; bf16 storage is the reported exception (profiles/r02_precision_budget.txt)."""
    fx = np.load(golden(name))
    sd, masks = fixture_state_dict(arch, fx)
    model = _build(arch, sd, masks, act)
    x = fixture_frames(fx).to(dev())
    with torch.no_grad():
        logprob, seg = model(x)
        labels = model.predict(x)
    torch.cuda.synchronize()
    ref_seg = torch.from_numpy(fx["seg"])
    assert seg.shape == ref_seg.shape and logprob.shape[2:] == x.shape[2:]
    # forward()[0] and predict() agree with each other exactly
    assert torch.equal(torch.max(logprob, 1)[1].to(torch.uint8), labels)
    differ = int((labels.cpu().numpy() != fx["labels"]).sum())
    npx = fx["labels"].size
    e_seg = rel_err(seg.cpu(), ref_seg)
    sample = logprob[0, :, ::7, ::13].cpu().numpy()
    e_lp = float(np.abs(sample - fx["logprob_sample"]).max() / np.abs(fx["seg"]).max())
    record("real-reference fixture %s" % name.replace(".npz", ""), act=act, frames="1x%dx%d" % tuple(x.shape[2:]),
           label_agreement=1.0 - differ / npx, pixels_differ=differ, pixels=npx, logits_rel_err=e_seg,
           logprob_rel_err=e_lp)
    assert e_seg <= LOGIT_RTOL and e_lp <= LOGIT_RTOL
    if act == "fp16":
        assert differ <= max(2, int(npx * (1 - LABEL_AGREE))), (differ, npx)
    else:
        assert differ <= int(0.03 * npx), (differ, npx)


def test_real_video_frames_against_the_real_reference():
    """SURVEY 8(d) inputs: three frames of the reference's sample.mp4, decoded / resized / normalised by the reference's
    own FrameCapture transforms (tests/golden/gen_golden_frames.py), block-pruned DRN-D-22; labels and low-res logits
    the REAL reference produced are in the fixture.  Both ingest routes: float32 NCHW and the fused uint8 path."""
    from oracle import frameio_oracle
    fx = np.load(golden("real_frames.npz"))
    model, sd, _ = _gate_case("drn_d_22", 64, 128, 1, True, "fp16", seed=int(fx["seed"]))
    x = frameio_oracle.ingest(fx["frames_u8"], fx["mean"], fx["std"])
    assert np.array_equal(x[0].numpy(), fx["x0"])                       # the reference's own transform, bit for bit
    ref_lab = torch.from_numpy(fx["labels"].astype(np.int64))
    assert float((ref_lab[0] != ref_lab[1]).float().mean()) > 0.01      # the network looks at the frames
    lab, oracle_lab = _gates("real frames sample.mp4 D-22 BlockPruner 75% 256x448", model, sd, x)
    assert torch.equal(oracle_lab, ref_lab)                             # oracle == real reference on these frames
    agree = (lab == ref_lab).float().mean().item()
    with torch.no_grad():
        seg = model(x.to(dev()))[1]
        model.set_ingest(fx["mean"], fx["std"])
        lab_u8 = model.predict(torch.from_numpy(fx["frames_u8"]).to(dev()))
    assert rel_err(seg.cpu(), torch.from_numpy(fx["seg"])) <= LOGIT_RTOL
    assert torch.equal(lab_u8.cpu().long(), lab)                        # fused uint8 ingest: identical labels
    assert agree >= LABEL_AGREE
    # bf16 storage on the same frames: reported
    model.set_act_dtype("bf16")
    model.set_masks(model._mask_dict)
    with torch.no_grad():
        lab_b = model.predict(x.to(dev())).cpu().long()
    record("real frames sample.mp4 D-22 BlockPruner 75% 256x448", act="bf16", frames="3x256x448",
           label_agreement=(lab_b == ref_lab).float().mean().item())


def _gate_case(arch, h, w, n, pruned, act, seed, cfg=None):
    model, sd, masks = gate_case_cpu(arch, pruned, seed, cfg)
    model.set_act_dtype(act)
    model = model.to(dev()).eval()
    model.set_masks(masks)
    x = recipe.make_frames(n, h, w, seed=99 + seed)
    return model, sd, x


def _gates(tag, model, sd, x, act="fp16", min_agree=LABEL_AGREE, logit_tol=LOGIT_RTOL, ref=None, frames_differ=True,
           miou_tol=MIOU_TOL):
    """the four north-star gates of one case against the fp32 oracle, on ALL pixels: low-res logits and log-probs
    within `logit_tol` of the logit range, argmax agreement >= min_agree, predict() == argmax(forward()[0]) bit for
    bit, mIoU of both label maps against a synthetic ground truth within 0.1 point.  Returns (labels, ref labels)."""
    ref_lp, ref_seg = ref if ref is not None else drn_oracle.drnseg_forward(sd, x)
    ref_lab = torch.max(ref_lp, 1)[1]
    xd = x.to(dev())
    with torch.no_grad():
        lp, seg = model(xd)
        lab = model.predict(xd)
    torch.cuda.synchronize()
    assert torch.equal(torch.max(lp, 1)[1].to(torch.uint8), lab)
    e_seg, e_lp = rel_err(seg.cpu(), ref_seg), rel_err(lp.cpu(), ref_lp)
    labc = lab.cpu().long()
    agree = (labc == ref_lab).float().mean().item()
    top2 = ref_lp.topk(2, dim=1)[0]
    confident = (top2[:, 0] - top2[:, 1]) > LOGIT_RTOL * ref_seg.abs().max()
    agree_conf = (labc == ref_lab)[confident].float().mean().item()
    gt = torch.randint(0, 19, ref_lab.shape, generator=torch.Generator().manual_seed(3))
    gt[ref_lab % 5 == 0] = 255
    gt = torch.where(torch.rand(gt.shape, generator=torch.Generator().manual_seed(4)) < 0.5, ref_lab, gt)
    ref_hist = drn_oracle.fast_hist(ref_lab.numpy().flatten(), gt.numpy().flatten(), 19)
    meter = drnb200.ConfusionMeter(19, dev())
    meter.update(lab, gt.to(dev()))
    assert np.array_equal(meter.hist.cpu().numpy(),
                          drn_oracle.fast_hist(labc.numpy().flatten(), gt.numpy().flatten(), 19))
    d_miou = abs(meter.miou() - drn_oracle.miou(ref_hist))
    # the synthetic network must actually look at its input, otherwise label parity says nothing about the front
    # kernels (round 1's 2x2-block recipe usually produced networks whose output ignored the frame)
    differ = float((ref_lab[0] != ref_lab[-1]).float().mean()) if x.shape[0] > 1 else None
    record(tag, act=act, frames="%dx%dx%d" % (x.shape[0], x.shape[2], x.shape[3]), label_agreement=agree,
           label_agreement_confident=agree_conf, confident_fraction=confident.float().mean().item(),
           logits_rel_err=e_seg, logprob_rel_err=e_lp, miou_delta=d_miou, classes=len(ref_lab.unique()),
           ref_labels_differ_between_frames=differ, asserted_min_agreement=min_agree)
    assert e_seg <= logit_tol and e_lp <= logit_tol, (tag, e_seg, e_lp)
    assert agree >= min_agree, (tag, agree)
    assert agree_conf >= LABEL_AGREE, (tag, agree_conf)
    assert d_miou <= miou_tol, (tag, d_miou)
    if frames_differ and differ is not None:
        assert differ > 0.01, (tag, "the synthetic network ignores its input", differ)
    return labc, ref_lab


@pytest.mark.parametrize("pruned", [False, True])
def test_parity_gates_drn_d_22(pruned):
    """the north-star gates at 256x512 (oracle finishes in seconds): fp16 activation storage"""
    model, sd, x = _gate_case("drn_d_22", 256, 512, 2, pruned, "fp16", seed=5)
    _gates("config 2 D-22 %s 256x512" % ("BlockPruner 75%" if pruned else "dense"), model, sd, x)


def test_parity_gates_at_benchmark_size():
    """BASELINE config 2 at ITS OWN size: two 1024x2048 frames of block-pruned DRN-D-22 through the launch list the
    benchmark times (13 ROW launches, projection in K, fused head) against the fp32 oracle on the same frames — all
    four gates on all pixels.  The oracle takes ~1 s per frame on the GPU box's host cores."""
    from oracle import frameio_oracle
    model, sd, x = _gate_case("drn_d_22", 1024, 2048, 1, True, "fp16", seed=12)
    # frame 0: white noise (torch.randn, the benchmark's input statistics); frame 1: a video-like frame (smooth field
    # + sensor noise) through the reference's ToTensor + Normalize, so that both input regimes are gated at this size
    fx = np.load(golden("frameio.npz"))
    x = torch.cat([x, frameio_oracle.ingest(recipe.make_u8_frames(1, 1024, 2048, seed=12).numpy(), fx["mean"], fx["std"])])
    lab, ref_lab = _gates("config 2 D-22 BlockPruner 75% 1024x2048 (benchmark size)", model, sd, x)
    eng = model.engine()
    modes = [ffi.lib().drnb200_conv_plan_mode(p) for op in eng.last_ops for p in op.plans.values()]
    assert eng.last_ops is eng.ops_proj and modes.count(5) + modes.count(6) == 13     # the launch list bench.py measures
    # forcing either accumulator layout everywhere gives the same labels bit for bit
    from drnb200.engine import ConvLayer
    base = model.predict(x.to(dev()))
    try:
        for lay in (1, 2):
            ConvLayer.acc_layout = lay
            assert torch.equal(model.predict(x.to(dev())), base), lay
    finally:
        ConvLayer.acc_layout = 0
    # uint8 labels straight from the fused head equal the int64 labels of torch.max on the host, frame by frame
    assert lab.shape == (2, 1024, 2048)


def _check_config_case(model, sd, x, min_agree, tag):
    return _gates(tag, model, sd, x, min_agree=min_agree, frames_differ=False)


def test_config3_drn_d_38_rmb_masks():
    """BASELINE config 3: DRN-D-38 with RmbPruner masks (row-wise outer sparsity 0.5 + blocklets)"""
    shapes = load_keys("drn_d_38")
    model, sd, x = _gate_case("drn_d_38", 128, 256, 2, True, "fp16", seed=21,
                              cfg=recipe.rmb_pruner_config(shapes, 0.5))
    _check_config_case(model, sd, x, LABEL_AGREE, "config 3 D-38 RmbPruner 128x256")
    dense, live, tile = model.engine().mac_counts(1, 128, 256)
    assert live < 0.2 * dense and tile < 0.62 * dense        # dead outer blocks are skipped as tiles


def test_config4_drn_d_54_srmbrep_masks():
    """BASELINE config 4: DRN-D-54 (Bottleneck) with per-layer srmbrep masks shaped like optimal_configs/drn_d_54:
    sparsity is finer than a tile, so every tile stays live and the result must still be exact in the mask"""
    shapes = load_keys("drn_d_54")
    model, sd, x = _gate_case("drn_d_54", 128, 256, 1, True, "fp16", seed=22,
                              cfg=recipe.srmbrep_config(shapes, 0.75))
    _check_config_case(model, sd, x, LABEL_AGREE, "config 4 D-54 srmbrep 128x256")
    dense, live, tile = model.engine().mac_counts(1, 128, 256)
    assert live < 0.3 * dense and tile > 0.95 * dense


def test_config5_drn_d_22_unstructured_90():
    """BASELINE config 5: torch.nn.utils.prune.l1_unstructured(amount=0.9) on EVERY Conv2d incl. stem and seg
    (semseg_unstructured.py:769-774); masks come from the weight_mask buffers, compacted to block tiles"""
    import torch.nn.utils.prune as prune
    model, sd, x = _gate_case("drn_d_22", 128, 256, 1, False, "fp16", seed=23)
    for _, module in model.named_modules():
        if isinstance(module, torch.nn.Conv2d):
            prune.l1_unstructured(module, name="weight", amount=0.9)
    eff = drnb200.checkpoint.normalize_state_dict(model.state_dict())
    eff = eff[0] if isinstance(eff, tuple) else eff
    for k in sd:
        if k in eff and k.endswith(".weight") and sd[k].dim() == 4 and not k.startswith("up."):
            nz = float((eff[k] != 0).float().mean())
            assert abs(nz - 0.1) < 0.01, (k, nz)
            sd[k] = eff[k].detach().cpu().clone()
    _check_config_case(model, sd, x, LABEL_AGREE, "config 5 D-22 unstructured 90% 128x256")
    dense, live, tile = model.engine().mac_counts(1, 128, 256)
    assert live < 0.11 * dense and tile > 0.9 * dense        # unstructured zeros leave (almost) every tile live


@pytest.mark.parametrize("arch", ["drn_d_22", "drn_d_38"])
def test_parity_gates_wide_frames_projection_in_k(arch):
    """frames wider than 1024 pixels switch blocks with a stride-1 1x1 shortcut (layers 5/6) to conv2 with the
    shortcut inside its K loop (DRNB200_KB_PROJ entries, engine.ProjResidualConv); same gates, 64 x 2048 frame"""
    from drnb200.engine import ProjResidualConv
    model, sd, x = _gate_case(arch, 64, 2048, 1, True, "fp16", seed=41)
    lab, ref_lab = _gates("wide frames %s BlockPruner 75%% 64x2048 (projection in K)" % arch, model, sd, x)
    lab = lab.to(torch.uint8).to(dev())
    eng = model.engine()
    assert eng.last_ops is eng.ops_proj and sum(isinstance(o, ProjResidualConv) for o in eng.last_ops) == 2
    for o in eng.last_ops:
        if isinstance(o, ProjResidualConv):
            kb = o.kblk.cpu().numpy()[:o.n_live]
            assert (kb >= ffi.KB_PROJ).sum() > 0 and ffi.lib().drnb200_conv_plan_mode(list(o.plans.values())[0]) in (5, 6)
    # the two launch lists give the same labels up to the re-rounded projection weights
    eng.proj_in_k = False
    lab_b = model.predict(x.to(dev()))
    assert eng.last_ops is eng.ops
    assert (lab_b == lab).float().mean().item() >= 0.999
    dense, live, tile = eng.mac_counts(1, 64, 2048)
    eng.proj_in_k = True
    model.predict(x.to(dev()))
    assert eng.mac_counts(1, 64, 2048) == (dense, live, tile)      # same work counted either way


@pytest.mark.parametrize("hw", [(300, 300), (100, 156), (77, 204)])
def test_frame_sizes_that_are_not_multiples_of_8(hw):
    """seg_video_old.py:127 resizes frames to 300x300: the stride-2 stages round up (150, 75, 38) and `up` returns
    8*38 = 304 rows, exactly like the reference; logits and labels against the oracle at such sizes"""
    model, sd, x = _gate_case("drn_d_22", hw[0], hw[1], 1, True, "fp16", seed=61)
    ref_lp, ref_seg = drn_oracle.drnseg_forward(sd, x)
    h8, w8 = -(-hw[0] // 8), -(-hw[1] // 8)
    assert tuple(ref_seg.shape[2:]) == (h8, w8) and tuple(ref_lp.shape[2:]) == (8 * h8, 8 * w8)
    lab, ref_lab = _gates("odd frame size D-22 BlockPruner 75%% %dx%d" % hw, model, sd, x, ref=(ref_lp, ref_seg))
    assert tuple(lab.shape) == (1, 8 * h8, 8 * w8) == tuple(ref_lab.shape)


def test_random_frame_shapes_sweep():
    """seeded sweep over frame shapes and batch sizes (any H >= 8, W % 4 == 0): output shapes follow the reference's
    rounding (8*ceil(H/8) x 8*ceil(W/8)), logits within tolerance, predict() == argmax(forward()[0]) bit for bit, and
    labels equal to the oracle's except for near-tie pixels.  This is a SHAPE-robustness test (ragged tiles of every
    kernel: halo / gather / row-halo / per-tap convs, the head's 15x7-cell tiles with half cells at all four borders),
    so the label bound is what 16-bit storage rounding alone explains: at most max(3 pixels, 1.5 x the count the CPU
    emulation of fp16 storage gives on the same frame) — e.g. 121x1028: 135 pixels on the GPU, 127 emulated (99.898 %
    vs 99.904 %); pixels whose fp32 margin exceeds the logit tolerance must agree exactly."""
    rng = np.random.RandomState(2026)
    model, sd, _ = _gate_case("drn_d_22", 64, 128, 1, True, "fp16", seed=33)
    worst = 1.0
    for case in range(10):
        n = int(rng.randint(1, 4))
        h = int(rng.randint(8, 200))
        w = 4 * int(rng.randint(2, 100))
        if case == 0:
            h, w = 8, 8                                       # the smallest accepted frame
        if case == 1:
            h, w = 121, 1028                                  # one row-halo tile plus 1 pixel at 1/8 resolution
        x = recipe.make_frames(n, h, w, seed=500 + case)
        ref_lp, ref_seg = drn_oracle.drnseg_forward(sd, x)
        with torch.no_grad():
            lp, seg = model(x.to(dev()))
            lab = model.predict(x.to(dev()))
        h8, w8 = -(-h // 8), -(-w // 8)
        assert tuple(lab.shape) == (n, 8 * h8, 8 * w8) == tuple(ref_lp.shape[0:1] + ref_lp.shape[2:]), (n, h, w)
        assert torch.equal(torch.max(lp, 1)[1].to(torch.uint8), lab), (n, h, w)
        assert rel_err(seg.cpu(), ref_seg) <= LOGIT_RTOL and rel_err(lp.cpu(), ref_lp) <= LOGIT_RTOL, (n, h, w)
        ref_lab = torch.max(ref_lp, 1)[1]
        differ = int((lab.cpu().long() != ref_lab).sum())
        npx = lab.numel()
        worst = min(worst, 1.0 - differ / npx)
        emu_lp, _ = drn_oracle.drnseg_forward(sd, x, quant=lambda role, key, t: t.half().float())
        emu = int((emu_lp.argmax(1) != ref_lab).sum())
        assert differ <= max(3, int(1.5 * emu) + 3), (n, h, w, differ, emu, npx)
        top2 = ref_lp.topk(2, dim=1)[0]
        confident = (top2[:, 0] - top2[:, 1]) > LOGIT_RTOL * ref_seg.abs().max()
        assert torch.equal(lab.cpu().long()[confident], ref_lab[confident]), (n, h, w)
    record("random frame shapes sweep (10 shapes, D-22 BlockPruner 75%): worst case", act="fp16", frames="8x8 .. 200x400",
           label_agreement=worst)


def test_bf16_storage_is_the_measured_exception():
    """bf16 activation storage (north_star's nominal layout).  Logits, log-probs and mIoU gates hold; the 99.9 % label
    gate does NOT on random-init networks and cannot with bf16 conv operands: replaying the engine's roundings inside
    the fp32 oracle gives 99.5-99.6 % for bf16 everywhere and still only 99.85 % with an fp32 residual stream and an
    fp32 hand-off into the head (profiles/r02_precision_budget.txt); the mIoU gate moves with it (0.3 point on the
    synthetic ground truth).  Asserted here: the measured floor (>= 99.3 %
    of all pixels), 99.9 % on pixels whose fp32 top-1/top-2 margin exceeds the logit tolerance, and that the
    disagreement is what the emulation predicts (bf16 rounding, not a kernel defect): <= 2x the emulated count."""
    model, sd, x = _gate_case("drn_d_22", 256, 512, 2, True, "bf16", seed=5)
    ref = drn_oracle.drnseg_forward(sd, x)
    # the synthetic ground truth copies the reference labels on half of the pixels, so 0.6 % flipped labels move the
    # mIoU by ~0.3 point: bf16 misses that gate as well on this network (recorded; bound asserted at 0.5)
    lab, ref_lab = _gates("config 2 D-22 BlockPruner 75% 256x512", model, sd, x, act="bf16", min_agree=0.993, ref=ref,
                          miou_tol=0.5)
    emu_lp, _ = drn_oracle.drnseg_forward(sd, x, quant=lambda role, key, t: t.to(torch.bfloat16).float())
    emu_dis = (emu_lp.argmax(1) != ref_lab).float().mean().item()
    dis = (lab != ref_lab).float().mean().item()
    record("config 2 D-22 BlockPruner 75% 256x512: CPU emulation of bf16 storage", act="bf16(emulated)",
           frames="2x256x512", label_agreement=1.0 - emu_dis)
    assert dis <= 2.0 * emu_dis + 1e-4, (dis, emu_dis)


@pytest.mark.parametrize("act", ["fp16", "bf16"])
@pytest.mark.parametrize("case", [("drn_d_22", 64, 128, True), ("drn_d_22", 64, 2048, True), ("drn_d_54", 64, 128, False)])
def test_per_layer_outputs_track_the_oracle(case, act):
    """EVERY stored activation of the engine (stem, each conv+BN(+res)+ReLU output, each stored shortcut) against
    the oracle tap of the same layer, NHWC 16-bit -> NCHW fp32.  Bound per layer: max |diff| over ALL elements <=
    min(depth + 4, 8) x half-ulp(storage type) x the layer's range, i.e. this layer's own rounding plus the propagated
    roundings of the layers before it (measured worst case on B200: 3.2 half-ulps for fp16, 3.4 for bf16 —
    profiles/r02_parity_table.txt); a wrong tap, tile or K-block shows up as an error of the order of the range."""
    arch, h, w, pruned = case
    model, sd, x = _gate_case(arch, h, w, 1, pruned, act, seed=8)
    ref = {}
    drn_oracle.drnseg_forward(sd, x, taps=ref)
    got = {}
    with torch.no_grad():
        model.engine().run(x.to(dev()), want_labels=True, taps=got)
    torch.cuda.synchronize()
    eng = model.engine()
    assert eng.launches_per_forward == 1 + len(eng.last_ops) + 1      # stem + convs + ONE fused head launch
    ulp = 2.0 ** -8 if act == "bf16" else 2.0 ** -11
    rows, depth = [], 0
    # shortcuts folded into conv2's K loop (projection in K) are never stored: every OTHER oracle tap must be present
    folded = {k for op in eng.last_ops for k in op.keys[1:] if not isinstance(op, drnb200.engine.FusedFirstConv)}
    assert set(ref) - folded == set(got), (sorted(set(ref) - folded - set(got)), sorted(set(got) - set(ref)))
    for key in ref:
        if key not in got:
            continue
        depth += 1
        a, b = got[key].cpu(), ref[key]
        assert a.shape == b.shape, (key, a.shape, b.shape)
        rng = float(b.abs().max())
        err = float((a - b).abs().max())
        rms = float((a - b).pow(2).mean().sqrt())
        rows.append((key, err / rng, rms / rng))
        assert err <= min(depth + 4, 8) * ulp * rng, (key, err, rng, depth)
    record("per-layer max/rms error relative to the layer's range: %s %dx%d %s" % (
        arch, h, w, "BlockPruner 75%" if pruned else "dense"), act=act,
        layers=[{"layer": k, "max": round(e, 6), "rms": round(r, 7)} for k, e, r in rows])
    dense, live, tile = eng.mac_counts(1, h, w)
    assert live <= tile <= dense
    if pruned:
        assert abs(live / dense - 0.27) < 0.03          # 75 % of the 24 prunable layers + dense stem/seg


def test_masks_from_zeros_and_from_torch_prune_give_the_same_tiles():
    """mask ingestion: Pruner.mask_dict == weight != 0 == torch.nn.utils.prune buffers"""
    import torch.nn.utils.prune as prune
    model, sd, x = _gate_case("drn_d_22", 64, 128, 1, True, "fp16", seed=9)
    xd = x.to(dev())
    lab_masks = model.predict(xd)
    lists_a = [(op.row_ptr.cpu().tolist(), op.kblk.cpu().tolist()[:op.n_live]) for op in model.engine().ops]
    model.set_masks(None)                          # liveness from zeros of the already-masked weights
    model.engine().ops[0].version = None
    for op in model.engine().ops:
        op.version = None
    lab_zeros = model.predict(xd)
    lists_b = [(op.row_ptr.cpu().tolist(), op.kblk.cpu().tolist()[:op.n_live]) for op in model.engine().ops]
    assert lists_a == lists_b and torch.equal(lab_masks, lab_zeros)
    conv = model.layer[8][0]
    prune.l1_unstructured(conv, "weight", amount=0.9)     # semseg_unstructured.py:770-773
    lab_pruned = model.predict(xd)
    op = [o for o in model.engine().ops if o.key == "layer.8.0"][0]
    assert op.live_elems <= int(0.1 * conv.weight_orig.numel()) + 1
    assert lab_pruned.shape == lab_masks.shape


def test_apply_masks_invalidates_the_cache():
    model, sd, x = _gate_case("drn_d_22", 64, 128, 1, False, "fp16", seed=10)
    xd = x.to(dev())
    a = model.predict(xd)
    with torch.no_grad():
        model.state_dict()["layer.8.0.weight"].mul_(0.0)     # what Pruner.apply_masks does in place
    b = model.predict(xd)
    op = [o for o in model.engine().ops if o.key == "layer.8.0"][0]
    assert op.n_live == 0 and not torch.equal(a, b)


def test_data_writes_need_invalidate_or_verify_weights():
    """`p.data` writes bypass the autograd version counter the cache is keyed on: stale until invalidate() (documented
    in drnb200/model.py), caught automatically by verify_weights=True"""
    model, sd, x = _gate_case("drn_d_22", 64, 128, 1, False, "fp16", seed=10)
    xd = x.to(dev())
    a = model.predict(xd)
    v0 = model.layer[8][0].weight._version
    model.layer[8][0].weight.data.mul_(0.0)
    assert model.layer[8][0].weight._version == v0                 # the premise: no version bump
    assert torch.equal(model.predict(xd), a)                       # stale by design
    model.invalidate()
    b = model.predict(xd)
    op = [o for o in model.engine().ops if o.key == "layer.8.0"][0]
    assert op.n_live == 0 and not torch.equal(a, b)
    # BN edits through .data and prepare(force=True): visible in the low-res logits
    seg_b = model(xd)[1]
    model.layer[8][1].bias.data.add_(0.5)
    assert torch.equal(model(xd)[1], seg_b)                        # stale
    model.prepare(force=True)
    assert not torch.equal(model(xd)[1], seg_b)
    model.layer[8][1].bias.data.sub_(0.5)
    model.invalidate()
    # verify_weights=True: no invalidate() needed
    m2 = drnb200.DRNSeg("drn_d_22", 19, pretrained=False, verify_weights=True)
    m2.load_state_dict(sd, strict=False)
    m2 = m2.to(dev()).eval()
    a2 = m2.predict(xd)
    assert torch.equal(a2, a)
    m2.layer[8][0].weight.data.mul_(0.0)
    assert torch.equal(m2.predict(xd), b)


def test_graph_replay_matches_eager_and_follows_the_weights():
    """DRNSeg.enable_graphs(): predict() replays one CUDA graph per input buffer — same kernels, same labels; a cache
    rebuild (new masks, apply_masks-style in-place writes, invalidate) drops the graphs; uint8 ingest + FramePipeline
    (fixed staging buffers, PIL-exact resize into a per-slot buffer) run through the graphs as well."""
    model, sd, x = _gate_case("drn_d_22", 128, 256, 2, True, "fp16", seed=31)
    x = x.to(dev())
    buf = torch.empty_like(x)
    frames = [x, x.flip(3).contiguous(), (x * 0.5 + 0.1).contiguous(), x.flip(2).contiguous()]
    eager = []
    for f in frames:
        buf.copy_(f)
        eager.append(model.predict(buf).clone())
    assert not torch.equal(eager[0], eager[1])
    model.enable_graphs()
    eng = model.engine(dev())
    for rep in range(2):
        for f, ref in zip(frames, eager):
            buf.copy_(f)
            got = model.predict(buf)
            assert torch.equal(got, ref)
    assert len(eng._graphs) == 1                       # one input buffer, one graph
    assert torch.equal(model.predict(frames[1]), eager[1])       # another buffer: its own graph (eager on first use)
    assert torch.equal(model.predict(frames[1]), eager[1])
    assert len(eng._graphs) == 2
    # in-place weight write (what Pruner.apply_masks does): version counters move, graphs are dropped and rebuilt
    w = model.state_dict()["layer.8.0.weight"]
    w.mul_(-1.0)
    buf.copy_(frames[0])
    changed = model.predict(buf).clone()
    assert len(eng._graphs) == 1
    model.enable_graphs(False)
    assert torch.equal(model.predict(buf), changed)
    assert not torch.equal(changed, eager[0])
    w.mul_(-1.0)
    assert torch.equal(model.predict(buf), eager[0])
    # uint8 frames through the streaming pipeline, with the resize: graphs on == graphs off
    model.set_ingest([0.29, 0.33, 0.28], [0.18, 0.19, 0.18])
    batches = [_u8_frames(2, 150, 280, seed=s) for s in range(5)]
    outs = {}
    for on in (False, True):
        model.enable_graphs(on)
        pipe = drnb200.FramePipeline(model, (2, 150, 280, 3), torch.uint8, resize_to=(96, 160), output="overlay")
        outs[on] = [o.clone() for o in pipe.run(batches + batches)]
        pipe.close()
    assert len(outs[True]) == 10
    for a, b in zip(outs[False], outs[True]):
        assert torch.equal(a, b)
    assert not torch.equal(outs[True][0], outs[True][1])
    model.enable_graphs(False)


_CHAIN_SCRIPT = """
import hashlib, sys
sys.path[:0] = [%r, %r, %r]
import torch
from helpers import gate_case_cpu
from oracle import recipe
model, sd, masks = gate_case_cpu("drn_d_22", True, 21)
model = model.cuda().eval(); model.set_masks(masks)
x = recipe.make_frames(2, 128, 512, seed=77).cuda()
h = hashlib.sha256()
for _ in range(3):                       # back-to-back forwards: every kernel follows another one of the chain
    final, seg = model(x)
    h.update(model.predict(x).cpu().numpy().tobytes()); h.update(seg.cpu().numpy().tobytes())
print("DIGEST", h.hexdigest())
"""


def test_chained_launches_are_bit_identical_to_plain_stream_order():
    """programmatic dependent launch: the persistent kernels start their prologue under the previous
    kernel's tail and wait before their first global read; DRNB200_PDL=0 launches in plain stream order.  Same labels
    and same low-res logits, bit for bit (the knob is read once per process, hence the two subprocesses)."""
    import subprocess
    import sys
    script = _CHAIN_SCRIPT % (os.path.join(ROOT, "video-seg-model-compress_b200"), os.path.join(ROOT, "tests"), ROOT)
    digests = []
    for knob in ("1", "0"):
        env = dict(os.environ, DRNB200_PDL=knob)
        out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([l for l in out.stdout.splitlines() if l.startswith("DIGEST")][-1])
    assert digests[0] == digests[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_frames_on_another_device_and_data_parallel():
    """frames on cuda:1 while cuda:0 is the current device run on cuda:1 (per-device engines, device guard), and
    nn.DataParallel(DRNSeg) — the reference's multi-GPU wrapper, semantic_seg.py:812 — gives the single-GPU result"""
    model, sd, x = _gate_case("drn_d_22", 64, 128, 4, True, "fp16", seed=13)
    x0 = x.to("cuda:0")
    with torch.no_grad():
        ref_lp, ref_seg = model(x0)
        ref_lab = model.predict(x0)
        assert torch.cuda.current_device() == 0
        lab1 = model.predict(x.to("cuda:1"))
        assert lab1.device.index == 1 and torch.equal(lab1.cpu(), ref_lab.cpu())
        assert set(model._engines) == {0, 1}
        dp = torch.nn.DataParallel(model, device_ids=[0, 1])
        lp, seg = dp(x0)
    assert torch.equal(lp.cpu(), ref_lp.cpu()) and torch.equal(seg.cpu(), ref_seg.cpu())
    assert set(model._engines) == {0, 1}                           # replicas reused the original's engines


@pytest.mark.parametrize("mode", ["pinned", "wc", "huge"])
def test_host_staging_buffers(mode):
    """drnb200_host_alloc / HostBuffer: pinned, write-combined and huge-page-registered frame buffers are valid sources
    of asynchronous H2D copies (and, for the cacheable modes, targets of D2H copies); foreign pointers are refused"""
    from drnb200.frameio import HostBuffer
    hb = HostBuffer((3, 64, 128, 3), torch.uint8, mode)
    ref = recipe.make_u8_frames(3, 64, 128, seed=4)
    hb.tensor.copy_(ref)
    d = torch.empty_like(ref, device=dev())
    d.copy_(hb.tensor, non_blocking=True)
    torch.cuda.synchronize()
    assert torch.equal(d.cpu(), ref)
    if mode != "wc":                                       # write-combined memory is for H2D sources only
        hb.tensor.zero_()
        hb.tensor.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        assert torch.equal(hb.tensor, ref)
    hb.close()
    hb.close()                                             # idempotent
    assert ffi.lib().drnb200_host_free(C.c_void_p(0x1000)) != 0      # not ours
    assert ffi.lib().drnb200_host_free(None) == 0
    with pytest.raises(ffi.Drnb200Error):
        HostBuffer((4,), torch.uint8, "mapped")


def test_input_validation():
    model, sd, x = _gate_case("drn_d_22", 64, 128, 1, False, "fp16", seed=11)
    with pytest.raises(ffi.Drnb200Error):
        model.predict(torch.zeros(1, 3, 60, 62, device=dev()))       # W % 4 != 0
    with pytest.raises(ffi.Drnb200Error):
        model.predict(torch.zeros(1, 3, 64, 64, device=dev(), dtype=torch.float16))
    out = model.predict(torch.zeros(3, 3, 72, 200, device=dev()))    # ragged map sizes (9 x 25 at 1/8)
    assert out.shape == (3, 72, 200)


def test_head_borders_against_conv_transpose():
    """the zero-padded ConvTranspose2d rule (not F.interpolate) incl. the outer 4 pixels, and first-max ties"""
    lib = ffi.lib()
    d = dev()
    g = torch.Generator().manual_seed(2)
    N, h, w, Cc, classes = 2, 6, 11, 64, 19
    feat = torch.randn(N, Cc, h, w, generator=g).half()
    sw = recipe.round_bf16(torch.randn(classes, Cc, generator=g) * 0.2)
    sb = 0.1 * torch.randn(classes, generator=g)
    sd = {"seg.weight": sw.half().float().view(classes, Cc, 1, 1), "seg.bias": sb}
    ref_lp, ref_seg = drn_oracle.head_forward(sd, feat.float())
    plan = C.c_void_p()
    swd, sbd = sw.to(d), sb.to(d)
    ffi.check(lib.drnb200_head_plan_create(C.byref(plan), N, h, w, Cc, classes, ffi.F16, ffi.ptr(swd), ffi.ptr(sbd),
                                           ffi.stream_ptr()))
    xd = feat.permute(0, 2, 3, 1).contiguous().to(d)
    lab = torch.empty(N, 8 * h, 8 * w, dtype=torch.uint8, device=d)
    seg = torch.empty(N, classes, h, w, dtype=torch.float32, device=d)
    lp = torch.empty(N, classes, 8 * h, 8 * w, dtype=torch.float32, device=d)
    ffi.check(lib.drnb200_head_forward(plan, ffi.ptr(xd), ffi.ptr(lab), ffi.ptr(seg), ffi.ptr(lp), ffi.stream_ptr()))
    torch.cuda.synchronize()
    lib.drnb200_head_plan_destroy(plan)
    assert (seg.cpu() - ref_seg).abs().max() <= 1e-3
    assert (lp.cpu() - ref_lp).abs().max() <= 2e-3
    ref_lab = torch.max(ref_lp, 1)[1]
    top2 = ref_lp.topk(2, dim=1)[0]
    clear = (top2[:, 0] - top2[:, 1]) > 1e-3
    assert torch.equal(lab.cpu().long()[clear], ref_lab[clear])
    # all-equal logits: every class ties, torch.max returns index 0
    plan = C.c_void_p()
    zw, zb = torch.zeros(classes, Cc, device=d), torch.zeros(classes, device=d)
    ffi.check(lib.drnb200_head_plan_create(C.byref(plan), N, h, w, Cc, classes, ffi.F16, ffi.ptr(zw), ffi.ptr(zb),
                                           ffi.stream_ptr()))
    ffi.check(lib.drnb200_head_forward(plan, ffi.ptr(xd), ffi.ptr(lab), None, None, ffi.stream_ptr()))
    torch.cuda.synchronize()
    lib.drnb200_head_plan_destroy(plan)
    assert int(lab.max()) == 0


def test_use_torch_up_against_the_real_reference():
    """DRNSeg(use_torch_up=True): nn.UpsamplingBilinear2d(scale_factor=8) (align_corners=True) instead of the fixed
    ConvTranspose2d (semantic_seg.py:144-145).  Fixture produced by the real reference; no `up.weight` key exists."""
    fx = np.load(golden("fwd_drn_d_22_40x72_torch_up.npz"))
    shapes = collections.OrderedDict((k, v) for k, v in load_keys("drn_d_22").items() if k != "up.weight")
    sd = recipe.make_state_dict(shapes, seed=int(fx["seed"]))
    model = drnb200.DRNSeg("drn_d_22", 19, pretrained=False, use_torch_up=True)
    assert "up.weight" not in model.state_dict()
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to(dev()).eval()
    x = recipe.make_frames(1, int(fx["hw"][0]), int(fx["hw"][1]), seed=1234 + int(fx["seed"])).to(dev())
    with torch.no_grad():
        lp, seg = model(x)
        lab = model.predict(x)
    assert not ffi.lib().drnb200_head_plan_fused(list(model.engine().head_plans.values())[0])
    assert torch.equal(torch.max(lp, 1)[1].to(torch.uint8), lab)
    rng = np.abs(fx["seg"]).max()
    e_seg = float(np.abs(seg.cpu().numpy() - fx["seg"]).max() / rng)
    e_lp = float(np.abs(lp.cpu().numpy() - fx["logprob"]).max() / rng)
    differ = int((lab.cpu().numpy() != fx["labels"]).sum())
    record("real-reference fixture use_torch_up=True drn_d_22 40x72", act="fp16", frames="1x40x72",
           label_agreement=1.0 - differ / fx["labels"].size, pixels_differ=differ, pixels=int(fx["labels"].size),
           logits_rel_err=e_seg, logprob_rel_err=e_lp)
    assert e_seg <= LOGIT_RTOL and e_lp <= LOGIT_RTOL
    assert differ <= max(2, int(fx["labels"].size * (1 - LABEL_AGREE)))


def test_confusion_matrix_and_ignore_label():
    fx = np.load(golden("metrics.npz"))
    pred = torch.from_numpy(fx["pred"].astype(np.uint8)).to(dev())
    lab64 = torch.from_numpy(fx["label"].astype(np.int64)).to(dev())
    lab8 = torch.from_numpy(fx["label"].astype(np.uint8)).to(dev())
    for lab in (lab64, lab8):
        meter = drnb200.ConfusionMeter(19, dev())
        meter.update(pred, lab)
        assert np.array_equal(meter.hist.cpu().numpy(), fx["hist"])
        assert meter.miou() == float(fx["miou"])
    empty = drnb200.ConfusionMeter(19, dev())
    empty.update(pred[:0], lab8[:0])
    assert int(empty.hist.sum()) == 0
    tiny = drnb200.fast_hist(torch.tensor([0, 1, 1, 2], dtype=torch.uint8, device=dev()),
                             torch.tensor([0, 1, 2, 255], dtype=torch.int64, device=dev()), 3)
    assert np.array_equal(tiny.cpu().numpy(), fx["tiny"])


def test_eval_loop_mirrors_of_test_and_val_miou(tmp_path):
    """evalops.test / evalops.val_miou (semantic_seg.py:429-468, :638-671) on a two-batch loader: mIoU equals the
    reference formula on the label maps predict() returns; save_vis writes the label and palette PNGs"""
    from drnb200 import evalops
    model, sd, x = _gate_case("drn_d_22", 64, 128, 4, True, "fp16", seed=51)
    gt = torch.randint(0, 19, (4, 64, 128), generator=torch.Generator().manual_seed(52))
    gt[:, :3] = 255
    loader = [(x[:2], gt[:2], ["a/f0.png", "a/f1.png"]), (x[2:], gt[2:], ["b/f2.png", "b/f3.png"])]
    with torch.no_grad():
        pred = model.predict(x.to(dev())).cpu().numpy().astype(np.int64)
    want = drn_oracle.miou(drn_oracle.fast_hist(pred.flatten(), gt.numpy().flatten(), 19))
    out = str(tmp_path / "pred")
    assert evalops.test(loader, model, 19, output_dir=out, has_gt=True, save_vis=True) == want
    assert evalops.val_miou([(b[0], b[1]) for b in loader], model, 19) == want
    assert evalops.test(loader, model, 19, has_gt=False) is None
    from PIL import Image
    lab = np.asarray(Image.open(out + "/b/f2.png"))
    col = np.asarray(Image.open(out + "_color/b/f2.png"))
    assert np.array_equal(lab, pred[2]) and np.array_equal(col, drnb200.CITYSCAPE_PALETTE[pred[2]])


def test_full_size_properties():
    """BASELINE size (1024x2048), beside test_parity_gates_at_benchmark_size: size-independent properties
    (i) tcgen05 and CUDA-core direct kernels agree on the same tile lists,
    (ii) frames are independent: predict(batch)[i] == predict(frame i),
    (iii) determinism: two runs are bit-identical, (iv) the label histogram sums to the pixel count."""
    model, sd, x = _gate_case("drn_d_22", 1024, 2048, 2, True, "fp16", seed=12)
    xd = x.to(dev())
    a = model.predict(xd)
    b = model.predict(xd)
    assert torch.equal(a, b)
    assert torch.equal(model.predict(xd[1:2])[0], a[1])
    meter = drnb200.ConfusionMeter(19, dev())
    meter.update(a, a)
    h = meter.hist.cpu().numpy()
    assert h.sum() == a.numel() and np.count_nonzero(h - np.diag(np.diag(h))) == 0
    eng = model.engine()
    eng.conv_impl = ffi.IMPL_DIRECT
    c = model.predict(xd[:1])
    eng.conv_impl = ffi.IMPL_AUTO
    agree = (c[0] == a[0]).float().mean().item()
    print("tcgen05 vs direct label agreement at 1024x2048: %.6f" % agree)
    assert agree >= 0.9995


# ------------------------------------------------------------------ rows either side of the path
def _u8_frames(n, h, w, seed, smooth=False):
    rng = np.random.RandomState(seed)
    if smooth:
        yy, xx = np.mgrid[0:h, 0:w]
        base = (np.sin(yy / 9.0)[..., None] * 60 + np.cos(xx / 13.0)[..., None] * 60 + 128
                + rng.randn(1, 1, 3) * 20)
        f = np.clip(base[None] + rng.randn(n, h, w, 3) * 3, 0, 255)
        return f.astype(np.uint8)
    return rng.randint(0, 256, size=(n, h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("act", ["fp16", "bf16"])
@pytest.mark.parametrize("hw", [(64, 128), (136, 48), (72, 208)])
def test_uint8_ingest_is_bit_identical_to_normalised_float_frames(act, hw):
    """frame ingest fused into the stem: predict(uint8 HWC) == predict(ToTensor+Normalize of the same frames)
    bit for bit (labels AND logits), because the table holds exactly the act_dtype rounding of the fp32 transform"""
    from oracle import frameio_oracle
    fx = np.load(golden("frameio.npz"))
    mean, std = fx["mean"], fx["std"]
    model, sd, _ = _gate_case("drn_d_22", 64, 128, 1, True, act, seed=21)
    frames = _u8_frames(2, hw[0], hw[1], seed=hw[0], smooth=(hw[1] == 48))
    x = frameio_oracle.ingest(frames, mean, std)                       # the reference's transform (CPU)
    model.set_ingest(mean, std)
    fd = torch.from_numpy(frames).to(dev())
    with torch.no_grad():
        lab_u8 = model.predict(fd)
        lab_f32 = model.predict(x.to(dev()))
        lp_u8, seg_u8 = model(fd)
        lp_f32, seg_f32 = model(x.to(dev()))
    assert torch.equal(lab_u8, lab_f32) and torch.equal(seg_u8, seg_f32) and torch.equal(lp_u8, lp_f32)
    # BGR frames (cv2 order) with bgr=True give the same result as the RGB frames
    model.set_ingest(mean, std, bgr=True)
    lab_bgr = model.predict(torch.from_numpy(np.ascontiguousarray(frames[..., ::-1])).to(dev()))
    assert torch.equal(lab_bgr, lab_u8)
    # and the whole thing still tracks the oracle on the normalised frames
    ref_lp, ref_seg = drn_oracle.drnseg_forward(sd, x)
    assert rel_err(seg_u8.cpu(), ref_seg) <= (LOGIT_RTOL if act == "fp16" else 2 * LOGIT_RTOL)


@pytest.mark.parametrize("act", [ffi.BF16, ffi.F16])
@pytest.mark.parametrize("shape", [(2, 150, 72), (1, 128, 16), (1, 261, 304), (3, 40, 20)])
def test_stem_kernels_vs_fp32(shape, act):
    """the 7x7 stem in isolation, through the C ABI: the tcgen05 Toeplitz kernel (drnb200_stem_plan_*: input and weights
    rounded to the activation dtype, fp32 accumulation, BN affine + ReLU, one output rounding) and the CUDA-core fp32
    cross-check (drnb200_stem_forward: no operand rounding) against torch conv2d; tile edges in both directions (rows
    beyond a 128-row tile, a single 8-column tile, widths that are not a multiple of the 16-column tile pair)"""
    lib = ffi.lib()
    N, H, W = shape
    d = dev()
    g = torch.Generator().manual_seed(H * W + N)
    tdt = torch.bfloat16 if act == ffi.BF16 else torch.float16
    x = torch.randn(N, 3, H, W, generator=g)
    w = torch.randn(16, 3, 7, 7, generator=g) * (2.0 / 147) ** 0.5
    scale = 0.5 + torch.rand(16, generator=g)
    shift = 0.2 * torch.randn(16, generator=g)
    aff = lambda t: (t * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).relu()      # noqa: E731
    ref16 = aff(torch.nn.functional.conv2d(x.to(tdt).float(), w.to(tdt).float(), None, 1, 3))
    ref32 = aff(torch.nn.functional.conv2d(x, w, None, 1, 3))
    xd, wd, sc, sh = x.to(d), w.to(d), scale.to(d), shift.to(d)
    st = ffi.stream_ptr()
    G = 4096
    flat = torch.full((N * H * W * 16 + 2 * G,), float("nan"), dtype=tdt, device=d)
    y = flat[G:G + N * H * W * 16].view(N, H, W, 16)
    plan = C.c_void_p()
    ffi.check(lib.drnb200_stem_plan_create(C.byref(plan), ffi.ptr(wd), ffi.ptr(sc), ffi.ptr(sh), N, H, W, 16, act, st))
    ffi.check(lib.drnb200_stem_plan_forward(plan, ffi.ptr(xd), ffi.ptr(y), st))
    torch.cuda.synchronize()
    lib.drnb200_stem_plan_destroy(plan)
    assert bool(torch.isnan(flat[:G]).all()) and bool(torch.isnan(flat[-G:]).all())      # nothing written outside y
    got = y.float().permute(0, 3, 1, 2).cpu()
    tol = 2.0 ** -8 if act == ffi.BF16 else 2.0 ** -10                                   # one output rounding
    assert (got - ref16).abs().max().item() <= tol * max(1.0, ref16.abs().max().item())
    y2 = torch.empty(N, H, W, 16, dtype=tdt, device=d)
    ffi.check(lib.drnb200_stem_forward(ffi.ptr(xd), ffi.ptr(wd), ffi.ptr(sc), ffi.ptr(sh), N, H, W, 16, act, ffi.ptr(y2), st))
    torch.cuda.synchronize()
    got2 = y2.float().permute(0, 3, 1, 2).cpu()
    assert (got2 - ref32).abs().max().item() <= tol * max(1.0, ref32.abs().max().item())


def test_uint8_ingest_validation():
    model, sd, _ = _gate_case("drn_d_22", 64, 128, 1, False, "fp16", seed=22)
    frames = torch.zeros(1, 64, 128, 3, dtype=torch.uint8, device=dev())
    with pytest.raises(ffi.Drnb200Error):
        model.predict(frames)                                           # set_ingest() not called
    model.set_ingest((0.3, 0.3, 0.3), (0.2, 0.2, 0.2))
    assert model.predict(frames).shape == (1, 64, 128)
    with pytest.raises(ffi.Drnb200Error):
        model.predict(torch.zeros(1, 64, 72, 3, dtype=torch.uint8, device=dev()))    # W % 16 != 0
    with pytest.raises(ffi.Drnb200Error):
        model.predict(torch.zeros(1, 3, 64, 128, dtype=torch.uint8, device=dev()))   # not HWC


def test_colorize_and_overlay_bit_exact():
    from oracle import frameio_oracle
    fx = np.load(golden("frameio.npz"))
    pred = fx["pred"].astype(np.uint8)
    got = drnb200.colorize(torch.from_numpy(pred).to(dev()))
    assert np.array_equal(got.cpu().numpy(), fx["color"])               # the real reference's CITYSCAPE_PALETTE[pred]
    rng = np.random.RandomState(3)
    lab = rng.randint(0, 19, size=(3, 40, 52)).astype(np.uint8)
    lab[0, 0, :7] = [255, 19, 20, 0, 18, 200, 7]                        # ignore / out-of-palette labels -> last row
    frames = rng.randint(0, 256, size=(3, 40, 52, 3), dtype=np.uint8)
    ld, fd = torch.from_numpy(lab).to(dev()), torch.from_numpy(frames).to(dev())
    assert np.array_equal(drnb200.colorize(ld).cpu().numpy(), frameio_oracle.colorize(lab))
    for alpha in (0.6, 0.0, 1.0, 0.37):
        assert np.array_equal(drnb200.overlay(ld, fd, alpha).cpu().numpy(),
                              frameio_oracle.overlay(lab, frames, alpha)), alpha
    pal = rng.randint(0, 256, size=(7, 3), dtype=np.uint8)
    assert np.array_equal(drnb200.colorize(ld, pal).cpu().numpy(), frameio_oracle.colorize(lab, pal))
    assert drnb200.colorize(torch.zeros(0, 4, dtype=torch.uint8, device=dev())).shape == (0, 4, 3)


def test_frame_resize_bit_exact_against_pil():
    """drnb200.resize_frames == T.Resize on the PIL frame (seg_video_old.py:125-128), bit for bit: files produced by
    real torchvision / PIL for down-, up- and single-axis scaling of a real video frame, plus the oracle on a batch
    of random frames at the caller's own size (640x1138 -> 300x300), plus guard bands around the output"""
    from oracle import frameio_oracle
    fx = np.load(golden("frame_resize.npz"))
    src = torch.from_numpy(fx["src"]).to(dev())[None]
    for i, (h, w) in enumerate(fx["sizes"]):
        got = drnb200.resize_frames(src, (int(h), int(w)))
        assert np.array_equal(got[0].cpu().numpy(), fx["dst%d" % i]), (h, w)
    assert torch.equal(drnb200.resize_frames(src, (160, 284)), src)                  # same size: a copy
    rng = np.random.RandomState(9)
    frames = rng.randint(0, 256, size=(2, 640, 1138, 3), dtype=np.uint8)
    ref = frameio_oracle.resize_u8(frames, 300, 300)
    got = drnb200.resize_frames(torch.from_numpy(frames).to(dev()), (300, 300))
    assert np.array_equal(got.cpu().numpy(), ref)
    # straight into the fused ingest: resized uint8 frames -> labels (304x304 like the reference, seg_video_old.py:127)
    model, sd, _ = _gate_case("drn_d_22", 64, 128, 1, True, "fp16", seed=24)
    fio = np.load(golden("frameio.npz"))
    model.set_ingest(fio["mean"], fio["std"])
    small = drnb200.resize_frames(torch.from_numpy(frames).to(dev()), (300, 304))     # W % 16 == 0 for the uint8 ingest
    lab = model.predict(small)
    assert lab.shape == (2, 304, 304)
    assert torch.equal(lab, model.predict(frameio_oracle.ingest(small.cpu().numpy(), fio["mean"], fio["std"]).to(dev())))
    with pytest.raises(ffi.Drnb200Error):
        drnb200.resize_frames(torch.from_numpy(frames), (300, 300))                  # CPU tensor: no fallback


def test_frame_pipeline_matches_direct_calls():
    """drnb200.FramePipeline (the reference's FrameCapture data flow, seg_video_old.py:110-203, as a three-stage
    streaming object): per-batch results arrive in order and equal resize_frames -> predict -> overlay called directly;
    more batches than buffers; float32 frames; evaluation mode returns the running confusion matrix"""
    from oracle import frameio_oracle
    model, sd, _ = _gate_case("drn_d_22", 64, 128, 1, True, "fp16", seed=25)
    fio = np.load(golden("frameio.npz"))
    model.set_ingest(fio["mean"], fio["std"])
    rng = np.random.RandomState(11)
    batches = [rng.randint(0, 256, size=(2, 90, 150, 3), dtype=np.uint8) for _ in range(7)]
    direct_lab, direct_ov = [], []
    for b in batches:
        small = drnb200.resize_frames(torch.from_numpy(b).to(dev()), (64, 96))
        lab = model.predict(small)
        direct_lab.append(lab.cpu().clone())
        direct_ov.append(drnb200.overlay(lab, small, 0.6).cpu().clone())
    pipe = drnb200.FramePipeline(model, (2, 90, 150, 3), torch.uint8, resize_to=(64, 96), output="labels", depth=3)
    got = [r.clone() for r in pipe.run(iter(batches))]
    assert len(got) == 7 and all(torch.equal(a, b) for a, b in zip(got, direct_lab))
    # frames decoded straight into the staging buffers, fewer batches than buffers, then reuse of the same object
    def feed():
        for k in range(2):
            pipe.staging(k).copy_(torch.from_numpy(batches[k]))
            yield pipe.staging(k)
    got = [r.clone() for r in pipe.run(feed())]
    assert len(got) == 2 and torch.equal(got[0], direct_lab[0]) and torch.equal(got[1], direct_lab[1])
    assert list(pipe.run(iter([]))) == []
    pipe.close()
    ov = drnb200.FramePipeline(model, (2, 90, 150, 3), torch.uint8, resize_to=(64, 96), output="overlay", depth=2,
                               host_mode="wc")
    got = [r.clone() for r in ov.run(iter(batches[:5]))]
    assert all(torch.equal(a, b) for a, b in zip(got, direct_ov[:5]))
    ov.close()
    # float32 NCHW frames (the reference's tensor) and the evaluation flow
    x = recipe.make_frames(2, 64, 128, seed=77)
    gt = torch.randint(0, 19, (2, 64, 128), generator=torch.Generator().manual_seed(5)).to(torch.uint8).to(dev())
    meter = drnb200.ConfusionMeter(19, dev())
    ev = drnb200.FramePipeline(model, (2, 3, 64, 128), torch.float32, output="hist", meter=(meter, gt))
    hists = [r.clone() for r in ev.run(iter([x, x, x]))]
    want = drnb200.ConfusionMeter(19, dev())
    want.update(model.predict(x.to(dev())), gt)
    assert torch.equal(hists[0], want.hist.cpu()) and torch.equal(hists[2], 3 * want.hist.cpu())
    ev.close()
    with pytest.raises(ffi.Drnb200Error):
        list(drnb200.FramePipeline(model, (2, 3, 64, 128), torch.float32).run(iter([x[:1]])))     # wrong batch shape
    with pytest.raises(ffi.Drnb200Error):
        drnb200.FramePipeline(model, (2, 3, 64, 128), torch.float32, resize_to=(32, 32))          # resize needs uint8
    del frameio_oracle


def test_full_size_uint8_ingest():
    """1024x2048: the fused uint8 ingest equals the float path fed with the reference's transform of the same frames"""
    from oracle import frameio_oracle
    fx = np.load(golden("frameio.npz"))
    model, sd, _ = _gate_case("drn_d_22", 64, 128, 1, True, "fp16", seed=23)
    frames = _u8_frames(1, 1024, 2048, seed=5, smooth=True)
    model.set_ingest(fx["mean"], fx["std"])
    a = model.predict(torch.from_numpy(frames).to(dev()))
    b = model.predict(frameio_oracle.ingest(frames, fx["mean"], fx["std"]).to(dev()))
    assert torch.equal(a, b)
    col = drnb200.overlay(a, torch.from_numpy(frames).to(dev()))
    assert np.array_equal(col.cpu().numpy(), frameio_oracle.overlay(a.cpu().numpy(), frames))


# ------------------------------------------------------------------------------------------------ 8f-4 multi-scale
def test_multiscale_matches_reference_fixture_bit_exact():
    """drnb200_ms_accumulate / drnb200_ms_argmax against what the REAL reference produced (resize_4d_tensor through
    PIL + sum + argmax, tests/golden/gen_golden_ms.py): float32 bit-exact, labels identical"""
    from helpers import ms_sources
    from drnb200 import multiscale
    fx = np.load(golden("multiscale.npz"))
    H, W = (int(v) for v in fx["target"])
    srcs = [t.to(dev()) for t in ms_sources()]
    for i, src in enumerate(srcs):
        acc = torch.full((src.shape[0], src.shape[1], H, W), float("nan"), device=dev())
        multiscale.resize_accumulate(src, acc, first=True)
        assert np.array_equal(acc.cpu().numpy(), fx["dst%d" % i]), tuple(src.shape)
    final, pred = multiscale.combine(srcs, H, W)
    assert np.array_equal(final.cpu().numpy(), fx["final"])
    assert np.array_equal(pred.cpu().numpy().astype(np.int64), fx["pred"])


@pytest.mark.parametrize("target", [(64, 136), (33, 57)])
def test_multiscale_against_oracle(target):
    """19 classes, two frames, the reference's five scales plus the frame itself; ragged target (scalar argmax path)"""
    from oracle import ms_oracle
    from drnb200 import multiscale
    H, W = target
    g = torch.Generator().manual_seed(H)
    srcs = [torch.log_softmax(2.0 * torch.randn(2, 19, int(H * s), int(W * s), generator=g), dim=1)
            for s in [1.0] + multiscale.SCALES]
    ref_final, ref_pred = ms_oracle.ms_combine([t.numpy() for t in srcs], W, H)
    final, pred = multiscale.combine([t.to(dev()) for t in srcs], H, W)
    assert np.array_equal(final.cpu().numpy(), ref_final)
    assert np.array_equal(pred.cpu().numpy().astype(np.int64), ref_pred)
    # guard bands around the accumulator and the label map: the kernels write nothing outside their tensors
    G = 4096
    flat = torch.full((2 * 19 * H * W + 2 * G,), -7.0, device=dev())
    acc = flat[G:G + 2 * 19 * H * W].view(2, 19, H, W)
    for i, t in enumerate(srcs):
        multiscale.resize_accumulate(t.to(dev()), acc, first=(i == 0))
    lflat = torch.full((2 * H * W + 2 * G,), 200, dtype=torch.uint8, device=dev())
    lab = lflat[G:G + 2 * H * W].view(2, H, W)
    ffi.check(ffi.lib().drnb200_ms_argmax(ffi.ptr(acc), 2, 19, H, W, ffi.ptr(lab), ffi.stream_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(acc.cpu().numpy(), ref_final) and np.array_equal(lab.cpu().numpy().astype(np.int64), ref_pred)
    assert bool((flat[:G] == -7.0).all()) and bool((flat[-G:] == -7.0).all())
    assert bool((lflat[:G] == 200).all()) and bool((lflat[-G:] == 200).all())
    # first maximum wins on ties (numpy argmax)
    tie = torch.zeros(1, 19, 8, 12, device=dev())
    tie[:, 7] = 1.0
    tie[:, 11] = 1.0
    assert int(multiscale.argmax_labels(tie).unique().item()) == 7
    with pytest.raises(ffi.Drnb200Error):
        multiscale.resize_accumulate(srcs[1], torch.zeros(2, 19, H, W), first=True)      # CPU tensor: no fallback


@pytest.mark.parametrize("src_hw", [(330, 57), (990, 40), (33, 570), (5, 7)])
def test_multiscale_large_factors(src_hw):
    """strong down/up-scaling: generic tap loops and the shorter row strips (scratch sized by the scale factor)"""
    from oracle import ms_oracle
    from drnb200 import multiscale
    H, W = 33, 57
    src = torch.randn(1, 3, src_hw[0], src_hw[1], generator=torch.Generator().manual_seed(src_hw[0]))
    ref = ms_oracle.resize_4d_tensor(src.numpy(), W, H)
    acc = torch.empty(1, 3, H, W, device=dev())
    multiscale.resize_accumulate(src.to(dev()), acc, first=True)
    assert np.array_equal(acc.cpu().numpy(), ref)
    multiscale.resize_accumulate(src.to(dev()), acc, first=False)
    assert np.array_equal(acc.cpu().numpy(), ref + ref)
    with pytest.raises(ffi.Drnb200Error):            # factor 100: beyond the scratch, rejected (never silently wrong)
        multiscale.resize_accumulate(torch.zeros(1, 1, 3300, 8, device=dev()), torch.zeros(1, 1, H, 8, device=dev()), True)


def test_multiscale_end_to_end_predict_ms_and_test_ms():
    """predict_ms / test_ms (mirror of semantic_seg.py:507-557) against the oracle's model + Pillow restatement"""
    import torch.nn.functional as F
    from oracle import ms_oracle
    from drnb200 import multiscale
    model, sd, x = _gate_case("drn_d_22", 64, 128, 2, True, "fp16", seed=31)
    images = [x] + [F.interpolate(x, size=(int(64 * s), int(128 * s)), mode="bicubic", align_corners=False)
                    for s in multiscale.SCALES]
    ref_out = [drn_oracle.drnseg_forward(sd, im)[0].numpy() for im in images]
    ref_final, ref_pred = ms_oracle.ms_combine(ref_out, 128, 64)
    pred = multiscale.predict_ms(model, [im.to(dev()) for im in images])
    agree = float((pred.cpu().numpy() == ref_pred).mean())
    print("multi-scale argmax agreement %.5f" % agree)
    assert agree >= 0.995
    gt = torch.from_numpy(ref_pred).clone()
    gt[:, ::5] = 255
    loader = [(images[0], gt, ["a.png", "b.png"]) + tuple(images[1:])]
    miou = multiscale.test_ms(loader, model, 19, multiscale.SCALES, has_gt=True)
    hist = drn_oracle.fast_hist(pred.cpu().numpy().flatten().astype(np.int64), gt.numpy().flatten(), 19)
    assert miou == drn_oracle.miou(hist)
    assert abs(miou - drn_oracle.miou(drn_oracle.fast_hist(ref_pred.flatten(), gt.numpy().flatten(), 19))) <= 0.5


def test_multiscale_full_size_properties():
    """size-independent properties at 1024x2048: a single same-size 'scale' reproduces predict(); constant planes
    survive the 1.75x -> 1x resample exactly (every coefficient row sums to 1)"""
    from drnb200 import multiscale
    model, sd, x = _gate_case("drn_d_22", 1024, 2048, 1, True, "fp16", seed=32)
    xd = x.to(dev())
    assert torch.equal(multiscale.predict_ms(model, [xd]), model.predict(xd))
    vals = torch.linspace(-7.0, -0.1, 19, device=dev()).view(1, 19, 1, 1)
    src = vals.expand(1, 19, 1792, 3584).contiguous()
    acc = torch.empty(1, 19, 1024, 2048, device=dev())
    multiscale.resize_accumulate(src, acc, first=True)
    assert torch.equal(acc, vals.expand_as(acc))
    multiscale.resize_accumulate(src[:, :, :512, :1024].contiguous(), acc, first=False)
    assert torch.equal(acc, (vals + vals).expand_as(acc))
    assert int(multiscale.argmax_labels(acc).unique().item()) == 18
