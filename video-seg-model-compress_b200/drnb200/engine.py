"""CUDA engine behind :class:`drnb200.DRNSeg` — turns the module tree into a list of kernel launches.

Per conv layer the engine keeps a *derived cache* of (tile list, packed 16-bit weights, folded BN affine,
C-ABI plan): the fp32 OIHW ``nn.Parameter`` stays the source of truth because the reference's pruners
read and mutate ``model.state_dict()[key]`` in place (pruners/Pruner.py:17-20).  The cache is keyed on
the tensors' ``_version`` counters and rebuilt when stale.

Data layout in HBM: activations NHWC 16-bit (bf16 or fp16, ``act_dtype``), one buffer per layer output
drawn from a size-keyed pool; packed weights = live K-blocks only, in the swizzled shared-memory image.
"""
import ctypes as C
import collections
import os

import torch
import torch.nn as nn

from . import ffi
from .drn import BasicBlock, Bottleneck

_DT = {"bf16": ffi.BF16, "fp16": ffi.F16, "f16": ffi.F16}


def _pick_tiles(cin, cout):
    """tile-list granularity the kernels consume: K-blocks of 64/32/16 input channels; 128-cout output
    tiles on the wide layers (MODE_T), the whole cout range on the narrow ones (MODE_P)."""
    tile_ci = 64 if cin % 64 == 0 else 32 if cin % 32 == 0 else 16 if cin % 16 == 0 else None
    if tile_ci is None:
        raise ffi.Drnb200Error("Cin=%d is not a multiple of 16" % cin)
    if cout % 128 == 0:
        tile_o = 128
    elif cout % 16 == 0 and cout <= 256:
        tile_o = cout
    elif cout % 8 == 0:
        tile_o = 8
    else:
        raise ffi.Drnb200Error("Cout=%d is not a multiple of 8" % cout)
    return tile_o, tile_ci



def _effective_weight(conv):
    """(weight tensor, version key) of a dense-path conv (stem, seg).  Under torch.nn.utils.prune the module's
    `weight` attribute is only refreshed by the forward pre-hook (semseg_unstructured.py:770-773 prunes the stem
    and `seg` too), which this path never runs: recompute `weight_orig * weight_mask` like the hook does."""
    if hasattr(conv, "weight_orig") and hasattr(conv, "weight_mask"):
        w, m = conv.weight_orig.detach(), conv.weight_mask.detach()
        return w * m.to(w.dtype), ("pruned", conv.weight_orig._version, conv.weight_mask._version)
    return conv.weight.detach(), conv.weight._version


class ConvLayer:
    """one conv (+BN) (+residual) (+ReLU) of the network and its derived device-side cache"""

    acc_layout = 0      # drnb200_conv_desc.acc_layout for new plans: 0 auto, 1 cout-major, 2 pixel-major (A/B runs, tests)

    def __init__(self, key, conv, bn, relu, residual_from=None, input_from=None):
        self.key = key                  # state_dict key prefix of the conv, e.g. 'layer.3.0.conv1'
        self.conv, self.bn, self.relu = conv, bn, relu
        self.input_from = input_from    # index of the op whose output feeds this conv (None = previous)
        self.residual_from = residual_from
        self.out_f32 = False
        self.keys = [key]               # state_dict conv keys this launch covers
        self.x_cpitch = 0               # input lives in a wider NHWC tensor (channels per pixel), 0 = packed
        self.res_cpitch = 0             # residual lives in a wider NHWC tensor
        self.res_coffset = 0            # first residual channel inside that tensor
        self.relu_n = 0                 # ReLU only on output channels < relu_n (0 = all, per self.relu)
        self.version = None
        self.row_ptr = self.kblk = self.w_packed = self.scale = self.shift = None
        self.n_live = 0
        self.tile_o = self.tile_ci = None
        self.live_elems = 0             # surviving weight elements (Pruner.print_stats numerator)
        self.plans = {}

    def destroy_plans(self):
        lib = ffi.lib()
        for p in self.plans.values():
            lib.drnb200_conv_plan_destroy(p)
        self.plans = {}

    @staticmethod
    def _conv_params(conv, key, mask_dict, device):
        """(w32 [O,I,kh,kw], mask32 {0,1}, version) of one nn.Conv2d under the three mask sources"""
        if hasattr(conv, "weight_orig") and hasattr(conv, "weight_mask"):   # torch.nn.utils.prune
            w, m = conv.weight_orig.detach(), conv.weight_mask.detach()
            ver = (w._version, m._version)
        else:
            w, m, ver = conv.weight.detach(), None, (conv.weight._version,)
            if mask_dict is not None:
                for k in (key + ".weight", "module." + key + ".weight"):
                    if k in mask_dict:
                        m = mask_dict[k]
                        ver = (w._version, id(m), m._version)
                        break
        w32 = w.to(device=device, dtype=torch.float32).contiguous()
        if m is None:
            mask32 = (w32 != 0).to(torch.float32)          # liveness from zeros (semantic_seg.py test path)
        else:
            mask32 = (m.to(device=device) != 0).to(torch.float32).contiguous()   # Hb masks may exceed 1
        return w32, mask32, ver

    @staticmethod
    def _bn_affine(bn, channels, device):
        """BatchNorm2d eval: y = (x-mean)/sqrt(var+eps)*gamma+beta (drn.py:7) as (scale, shift, version)"""
        if bn is None:
            return (torch.ones(channels, dtype=torch.float32, device=device),
                    torch.zeros(channels, dtype=torch.float32, device=device), ())
        inv = torch.rsqrt(bn.running_var.detach().to(device, torch.float32) + bn.eps)
        scale = bn.weight.detach().to(device, torch.float32) * inv
        shift = bn.bias.detach().to(device, torch.float32) - bn.running_mean.detach().to(device, torch.float32) * scale
        ver = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version)
        return scale.contiguous(), shift.contiguous(), ver

    def version_key(self, mask_dict):
        conv = self.conv
        if hasattr(conv, "weight_orig") and hasattr(conv, "weight_mask"):
            ver = (conv.weight_orig._version, conv.weight_mask._version)
        else:
            ver = (conv.weight._version,)
            if mask_dict is not None:
                for k in (self.key + ".weight", "module." + self.key + ".weight"):
                    if k in mask_dict:
                        ver = ver + (id(mask_dict[k]), mask_dict[k]._version)
                        break
        bn = self.bn
        if bn is not None:
            ver = ver + (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version)
        return ver

    def params(self, mask_dict, device):
        """-> (w32, mask32, scale, shift) of this launch"""
        w32, mask32, _ = self._conv_params(self.conv, self.key, mask_dict, device)
        scale, shift, _ = self._bn_affine(self.bn, w32.shape[0], device)
        return w32, mask32, scale, shift

    @property
    def out_channels(self):
        return self.conv.out_channels

    def dense_macs_per_pixel(self):
        c = self.conv
        return c.out_channels * c.in_channels * c.kernel_size[0] ** 2

    def refresh(self, mask_dict, act_dtype, device, stamp=()):
        """(re)build tile list, packed weights and BN affine if the parameters changed"""
        ver = self.version_key(mask_dict) + (act_dtype, str(device), stamp)
        if ver == self.version:
            return False
        lib = ffi.lib()
        self.destroy_plans()
        w32, mask32, self.scale, self.shift = self.params(mask_dict, device)
        O, I, kh, kw = w32.shape
        self.tile_o, self.tile_ci = _pick_tiles(I, O)
        n_ot, n_kb = O // self.tile_o, (I // self.tile_ci) * kh * kw
        self.row_ptr = torch.empty(n_ot + 1, dtype=torch.int32, device=device)
        self.kblk = torch.empty(max(1, n_ot * n_kb), dtype=torch.int32, device=device)
        n_live = torch.zeros(1, dtype=torch.int32, device=device)
        st = ffi.stream_ptr()
        ffi.check(lib.drnb200_compact_mask(ffi.ptr(mask32), O, I, kh, kw, self.tile_o, self.tile_ci,
                                           ffi.ptr(self.row_ptr), ffi.ptr(self.kblk), ffi.ptr(n_live), st),
                  "compact_mask(%s)" % self.key)
        self.n_live = int(n_live.item())
        self.w_packed = torch.empty(max(1, self.n_live) * self.tile_o * self.tile_ci, dtype=torch.int16,
                                    device=device)
        ffi.check(lib.drnb200_pack_weights(ffi.ptr(w32), ffi.ptr(mask32), O, I, kh, kw, self.tile_o,
                                           self.tile_ci, ffi.ptr(self.row_ptr), ffi.ptr(self.kblk),
                                           act_dtype, ffi.ptr(self.w_packed), st),
                  "pack_weights(%s)" % self.key)
        self.live_elems = int(torch.count_nonzero(w32 * mask32).item())
        self.version = ver
        return True

    def plan(self, N, H, W, act_dtype, impl):
        k = (N, H, W, act_dtype, impl, self.out_f32, ConvLayer.acc_layout)
        p = self.plans.get(k)
        if p is None:
            conv = self.conv
            d = ffi.ConvDesc(N=N, H=H, W=W, Cin=conv.in_channels, Cout=self.out_channels,
                             ksize=conv.kernel_size[0], stride=conv.stride[0], dilation=conv.dilation[0],
                             relu=int(self.relu), has_residual=int(self.residual_from is not None),
                             act_dtype=act_dtype, out_f32=int(self.out_f32), tile_o=self.tile_o,
                             tile_ci=self.tile_ci, impl=impl, x_cpitch=self.x_cpitch,
                             res_cpitch=self.res_cpitch, res_coffset=self.res_coffset, relu_n=self.relu_n,
                             acc_layout=ConvLayer.acc_layout)
            h = C.c_void_p()
            ffi.check(ffi.lib().drnb200_conv_plan_create(
                C.byref(h), C.byref(d), ffi.ptr(self.row_ptr), ffi.ptr(self.kblk), ffi.ptr(self.w_packed),
                ffi.ptr(self.scale), ffi.ptr(self.shift)), "conv_plan_create(%s)" % self.key)
            p = self.plans[k] = h
        return p

    def out_hw(self, H, W):
        s = self.conv.stride[0]
        return (H - 1) // s + 1, (W - 1) // s + 1


class FusedFirstConv(ConvLayer):
    """conv1 (3x3, stride s) of a BasicBlock and the block's downsample (1x1, stride s; drn.py:181-186) as ONE
    launch with output channels [conv1 | downsample]: the 1x1 projection reads exactly the centre tap of the
    3x3 window (padding == dilation), so it is embedded as a 3x3 filter whose other eight taps are zero — dead
    K-blocks that the tile list skips.  ReLU applies to the conv1 half only.  conv2 then takes its input from
    channels [0, C) and its residual from channels [C, 2C) of the 2C-channel result."""

    def __init__(self, key, block, input_from):
        super().__init__(key + ".conv1", block.conv1, block.bn1, True, input_from=input_from)
        self.block = block
        self.ds_conv, self.ds_bn = block.downsample[0], block.downsample[1]
        self.ds_key = key + ".downsample.0"
        self.keys = [key + ".conv1", self.ds_key]
        self.relu_n = block.conv1.out_channels

    @staticmethod
    def applicable(block):
        ds = block.downsample
        c1 = block.conv1
        return (ds is not None and getattr(block, "residual", True) and c1.kernel_size == (3, 3)
                and ds[0].kernel_size == (1, 1) and ds[0].stride == c1.stride
                and ds[0].out_channels == c1.out_channels and ds[0].in_channels == c1.in_channels)

    @property
    def out_channels(self):
        return 2 * self.conv.out_channels

    def dense_macs_per_pixel(self):
        c = self.conv
        return c.out_channels * c.in_channels * 9 + c.out_channels * c.in_channels

    def version_key(self, mask_dict):
        ver = super().version_key(mask_dict)
        ds, bn = self.ds_conv, self.ds_bn
        if hasattr(ds, "weight_orig") and hasattr(ds, "weight_mask"):
            ver = ver + (ds.weight_orig._version, ds.weight_mask._version)
        else:
            ver = ver + (ds.weight._version,)
            if mask_dict is not None:
                for k in (self.ds_key + ".weight", "module." + self.ds_key + ".weight"):
                    if k in mask_dict:
                        ver = ver + (id(mask_dict[k]), mask_dict[k]._version)
                        break
        return ver + (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version)

    def params(self, mask_dict, device):
        w1, m1, _ = self._conv_params(self.conv, self.key, mask_dict, device)
        wd, md, _ = self._conv_params(self.ds_conv, self.ds_key, mask_dict, device)
        s1, b1, _ = self._bn_affine(self.bn, w1.shape[0], device)
        sd, bd, _ = self._bn_affine(self.ds_bn, wd.shape[0], device)
        wd3, md3 = torch.zeros_like(w1), torch.zeros_like(m1)
        wd3[:, :, 1, 1] = wd[:, :, 0, 0]
        md3[:, :, 1, 1] = md[:, :, 0, 0]
        return (torch.cat([w1, wd3]).contiguous(), torch.cat([m1, md3]).contiguous(),
                torch.cat([s1, sd]).contiguous(), torch.cat([b1, bd]).contiguous())


class ProjFallback(Exception):
    """the BatchNorm fold of a residual projection is ill-conditioned: keep the projection as its own output"""


class ProjResidualConv(ConvLayer):
    """conv2 (3x3, stride 1) of a BasicBlock whose shortcut is a STRIDE-1 1x1 conv + BN (drn.py:181-186; layers 5
    and 6 of DRN-D-22/38), with the shortcut computed INSIDE conv2's K loop instead of being written to HBM by the
    previous launch and read back as a residual:

        relu(s2*conv2(h) + b2 + sd*proj(x) + bd) = relu(s2*(conv2(h) + proj'(x)) + (b2 + bd)),  proj' = (sd/s2) * proj

    The projection's live 128x64 weight blocks become extra entries (DRNB200_KB_PROJ + 3*cib) at the end of every
    output tile's K list; the kernel loads their activations from the block's input tensor (passed through the
    `residual` argument).  proj' is re-rounded to 16 bits after the per-channel rescale (<= 2^-11 relative per
    weight); when sd/s2 is not finite or exceeds 1024 in magnitude the fold is refused (ProjFallback) and the engine
    keeps the [conv1 | downsample] form.  Only the row-halo kernel runs this (feature maps wider than 128 pixels)."""

    MAX_RATIO = 1024.0
    MAX_SUBNORMAL_FRACTION = 1e-3     # of the live rescaled shortcut weights (random-init Gaussians: ~1e-4 at ratio 1)

    def __init__(self, key, block, input_from, proj_from, proj_pitch):
        super().__init__(key + ".conv2", block.conv2, block.bn2, True, residual_from=proj_from,
                         input_from=input_from)
        self.ds_conv, self.ds_bn = block.downsample[0], block.downsample[1]
        self.ds_key = key + ".downsample.0"
        self.keys = [key + ".conv2", self.ds_key]
        self.proj_cin = self.ds_conv.in_channels
        self.res_cpitch = proj_pitch

    @staticmethod
    def applicable(block):
        ds, c1, c2 = block.downsample, block.conv1, block.conv2
        return (ds is not None and getattr(block, "residual", True) and ds[0].kernel_size == (1, 1)
                and ds[0].stride == (1, 1) and c1.stride == (1, 1) and c2.kernel_size == (3, 3)
                and c2.stride == (1, 1) and c2.dilation[0] <= 4 and c2.out_channels % 128 == 0
                and c2.in_channels % 64 == 0 and ds[0].in_channels % 64 == 0
                and ds[0].out_channels == c2.out_channels)

    def dense_macs_per_pixel(self):
        c, d = self.conv, self.ds_conv
        return c.out_channels * c.in_channels * 9 + d.out_channels * d.in_channels

    def version_key(self, mask_dict):
        return super().version_key(mask_dict) + _conv_bn_version(self.ds_conv, self.ds_bn, self.ds_key, mask_dict)

    def _compact_and_pack(self, w32, mask32, act_dtype, device):
        lib, st = ffi.lib(), ffi.stream_ptr()
        O, I, kh, kw = w32.shape
        n_ot, n_kb = O // 128, (I // 64) * kh * kw
        row_ptr = torch.empty(n_ot + 1, dtype=torch.int32, device=device)
        kblk = torch.empty(max(1, n_ot * n_kb), dtype=torch.int32, device=device)
        n_live = torch.zeros(1, dtype=torch.int32, device=device)
        ffi.check(lib.drnb200_compact_mask(ffi.ptr(mask32), O, I, kh, kw, 128, 64, ffi.ptr(row_ptr), ffi.ptr(kblk),
                                           ffi.ptr(n_live), st), "compact_mask(%s)" % self.key)
        n = int(n_live.item())
        packed = torch.empty((max(1, n), 128 * 64), dtype=torch.int16, device=device)
        ffi.check(lib.drnb200_pack_weights(ffi.ptr(w32), ffi.ptr(mask32), O, I, kh, kw, 128, 64, ffi.ptr(row_ptr),
                                           ffi.ptr(kblk), act_dtype, ffi.ptr(packed), st), "pack_weights(%s)" % self.key)
        return row_ptr.cpu().tolist(), kblk[:n], packed[:n], n

    def refresh(self, mask_dict, act_dtype, device, stamp=()):
        ver = self.version_key(mask_dict) + (act_dtype, str(device), stamp)
        if ver == self.version:
            return False
        self.destroy_plans()
        w2, m2, _ = self._conv_params(self.conv, self.key, mask_dict, device)
        wd, md, _ = self._conv_params(self.ds_conv, self.ds_key, mask_dict, device)
        s2, b2, _ = self._bn_affine(self.bn, w2.shape[0], device)
        sd, bd, _ = self._bn_affine(self.ds_bn, wd.shape[0], device)
        ratio = sd / s2
        if not bool(torch.isfinite(ratio).all()) or float(ratio.abs().max()) > self.MAX_RATIO:
            raise ProjFallback(self.key)
        if act_dtype == ffi.F16:
            # fp16 has 5 exponent bits: rescaled shortcut weights below 2^-14 would be rounded as subnormals (or to
            # zero) and the lost relative precision is multiplied back by s2 in the epilogue; above 65504 they overflow
            folded = (wd * md.ne(0)).abs() * ratio.abs().view(-1, 1, 1, 1)
            live = folded[folded > 0]
            if live.numel() and (float(live.max()) > 6.0e4 or
                                 float((live < 2.0 ** -14).float().mean()) > self.MAX_SUBNORMAL_FRACTION):
                raise ProjFallback(self.key)
        self.tile_o, self.tile_ci = 128, 64
        rp2, kb2, pk2, n2 = self._compact_and_pack(w2, m2, act_dtype, device)
        rpd, kbd, pkd, nd = self._compact_and_pack((wd * ratio.view(-1, 1, 1, 1)).contiguous(), md, act_dtype, device)
        # merge per output tile: the conv's own entries, then the projection's (sorted by construction)
        kbd = kbd * 3 + ffi.KB_PROJ
        kparts, wparts, row_ptr = [], [], [0]
        for ot in range(len(rp2) - 1):
            kparts += [kb2[rp2[ot]:rp2[ot + 1]], kbd[rpd[ot]:rpd[ot + 1]]]
            wparts += [pk2[rp2[ot]:rp2[ot + 1]], pkd[rpd[ot]:rpd[ot + 1]]]
            row_ptr.append(row_ptr[-1] + (rp2[ot + 1] - rp2[ot]) + (rpd[ot + 1] - rpd[ot]))
        self.n_live = n2 + nd
        self.kblk = torch.cat(kparts + [torch.zeros(1, dtype=torch.int32, device=device)]).contiguous()
        self.w_packed = torch.cat(wparts + [torch.zeros((1, 128 * 64), dtype=torch.int16, device=device)]).contiguous()
        self.row_ptr = torch.tensor(row_ptr, dtype=torch.int32, device=device)
        self.scale, self.shift = s2.contiguous(), (b2 + bd).contiguous()
        self.live_elems = int(torch.count_nonzero(w2 * m2).item()) + int(torch.count_nonzero(wd * md).item())
        self.version = ver
        return True

    def plan(self, N, H, W, act_dtype, impl):
        k = (N, H, W, act_dtype, impl, ConvLayer.acc_layout)
        p = self.plans.get(k)
        if p is None:
            conv = self.conv
            d = ffi.ConvDesc(N=N, H=H, W=W, Cin=conv.in_channels, Cout=conv.out_channels, ksize=3, stride=1,
                             dilation=conv.dilation[0], relu=1, has_residual=0, act_dtype=act_dtype, out_f32=0,
                             tile_o=128, tile_ci=64, impl=ffi.IMPL_TCGEN05, x_cpitch=self.x_cpitch,
                             res_cpitch=self.res_cpitch, res_coffset=0, relu_n=0, proj_cin=self.proj_cin,
                             acc_layout=ConvLayer.acc_layout)
            h = C.c_void_p()
            ffi.check(ffi.lib().drnb200_conv_plan_create(
                C.byref(h), C.byref(d), ffi.ptr(self.row_ptr), ffi.ptr(self.kblk), ffi.ptr(self.w_packed),
                ffi.ptr(self.scale), ffi.ptr(self.shift)), "conv_plan_create(%s + projection)" % self.key)
            p = self.plans[k] = h
        return p


def _conv_bn_version(conv, bn, key, mask_dict):
    if hasattr(conv, "weight_orig") and hasattr(conv, "weight_mask"):
        ver = (conv.weight_orig._version, conv.weight_mask._version)
    else:
        ver = (conv.weight._version,)
        if mask_dict is not None:
            for k in (key + ".weight", "module." + key + ".weight"):
                if k in mask_dict:
                    ver = ver + (id(mask_dict[k]), mask_dict[k]._version)
                    break
    return ver + (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version)


def _check_conv(conv, key):
    k, s, d, p = conv.kernel_size, conv.stride, conv.dilation, conv.padding
    ok = (k[0] == k[1] and k[0] in (1, 3) and s[0] == s[1] and d[0] == d[1] and conv.groups == 1
          and conv.bias is None and p[0] == p[1] == d[0] * (k[0] // 2))
    if not ok:
        raise ffi.Drnb200Error("conv %s: unsupported configuration %r" % (key, conv))


class _Timed:
    """context manager recording a CUDA-event pair around a launch group (no-op when disabled)"""

    def __init__(self, name, sink):
        self.name, self.sink = name, sink

    def __enter__(self):
        if self.sink is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.sink is not None:
            self.e1.record()
            self.sink.append((self.name, self.e0, self.e1))
        return False


def ingest_lut(mean, std, act_dtype):
    """[3,256] int16 table of act_dtype bit patterns: ((b / 255) - mean[c]) / std[c] in correctly rounded fp32
    steps (the reference's torch CPU transforms), then one rounding to act_dtype.  Host function of the C ABI."""
    m = (C.c_float * 3)(*[float(v) for v in mean])
    sd = (C.c_float * 3)(*[float(v) for v in std])
    out = torch.empty((3, 256), dtype=torch.int16)
    ffi.check(ffi.lib().drnb200_ingest_lut(m, sd, int(act_dtype), out.data_ptr()), "ingest_lut")
    return out


class Engine:
    """builds and runs the launch list for one DRNSeg module"""

    def __init__(self, seg_module, act_dtype="fp16", conv_impl=ffi.IMPL_AUTO, fuse_downsample=True,
                 verify_weights=False):
        self.m = seg_module
        self.verify_weights = verify_weights      # fingerprint parameter CONTENTS on every call (catches `.data` writes)
        self.epoch = 0                            # bumped by invalidate(); part of every cache key
        self.fuse_downsample = fuse_downsample    # [conv1 | 1x1 downsample] of a BasicBlock in one launch
        self.act_dtype = _DT[act_dtype] if isinstance(act_dtype, str) else int(act_dtype)
        self.conv_impl = conv_impl
        self.mask_dict = None
        self.stem = None       # (conv, bn)
        self.ops = []          # ConvLayer list in execution order
        self.head_plans = {}
        self.head_version = None
        self.stem_plans = {}
        self.ingest = None     # (mean, std, bgr) for uint8 frames
        self._lut = None
        self.stem_impl = "tcgen05"   # or "direct": CUDA-core fp32 stem (no input rounding), cross-check
        self.launches_per_forward = 0
        self.proj_in_k = os.environ.get("DRNB200_PROJ", "1") != "0"    # A/B knob: "0" keeps [conv1 | shortcut] everywhere
        self.last_ops = None
        self._graphs = collections.OrderedDict()   # CUDA graphs of the label path, keyed by input buffer + shape
        self._graph_pool = None
        self._build_graph()

    # ---- module tree -> op list ----------------------------------------------------------------
    def _build_graph(self):
        root = getattr(self.m, "layer", None)
        prefix = "layer"
        if root is None:
            root, prefix = self.m.base, "base"        # seg_video.py:79 flavour
        ops = []
        pending = {"conv": None, "key": None}

        def flush_plain(bn, relu):
            conv, key = pending["conv"], pending["key"]
            pending["conv"] = None
            if conv.kernel_size[0] == 7:
                if conv.in_channels != 3 or conv.out_channels != 16 or self.stem is not None or ops:
                    raise ffi.Drnb200Error("unexpected 7x7 conv %s" % key)
                self.stem = (conv, bn, key)
                if not relu:
                    raise ffi.Drnb200Error("stem without ReLU is not supported")
                return
            _check_conv(conv, key)
            ops.append(ConvLayer(key, conv, bn, relu))

        def walk(mod, key):
            children = list(mod.named_children())
            if isinstance(mod, BasicBlock):
                src = len(ops) - 1          # output index feeding this block (-1 = stem output)
                _check_conv(mod.conv1, key + ".conv1"); _check_conv(mod.conv2, key + ".conv2")
                if self.fuse_downsample and FusedFirstConv.applicable(mod):
                    _check_conv(mod.downsample[0], key + ".downsample.0")
                    planes = mod.conv1.out_channels
                    ops.append(FusedFirstConv(key, mod, input_from=src))
                    f = len(ops) - 1
                    c2 = ConvLayer(key + ".conv2", mod.conv2, mod.bn2, True, residual_from=f, input_from=f)
                    c2.x_cpitch, c2.res_cpitch, c2.res_coffset = 2 * planes, 2 * planes, planes
                    ops.append(c2)
                    return
                ops.append(ConvLayer(key + ".conv1", mod.conv1, mod.bn1, True, input_from=src))
                res = None
                if getattr(mod, "residual", True):
                    res = src
                    if mod.downsample is not None:
                        dconv, dbn = mod.downsample[0], mod.downsample[1]
                        _check_conv(dconv, key + ".downsample.0")
                        ops.append(ConvLayer(key + ".downsample.0", dconv, dbn, False, input_from=src))
                        res = len(ops) - 1
                c1 = len(ops) - 1 if res is None or mod.downsample is None else len(ops) - 2
                ops.append(ConvLayer(key + ".conv2", mod.conv2, mod.bn2, True, residual_from=res,
                                     input_from=c1))
                return
            if isinstance(mod, Bottleneck):
                src = len(ops) - 1
                for n in ("conv1", "conv2", "conv3"):
                    _check_conv(getattr(mod, n), key + "." + n)
                ops.append(ConvLayer(key + ".conv1", mod.conv1, mod.bn1, True, input_from=src))
                ops.append(ConvLayer(key + ".conv2", mod.conv2, mod.bn2, True))
                c2 = len(ops) - 1
                res = src
                if mod.downsample is not None:
                    dconv, dbn = mod.downsample[0], mod.downsample[1]
                    _check_conv(dconv, key + ".downsample.0")
                    ops.append(ConvLayer(key + ".downsample.0", dconv, dbn, False, input_from=src))
                    res = len(ops) - 1
                ops.append(ConvLayer(key + ".conv3", mod.conv3, mod.bn3, True, residual_from=res,
                                     input_from=c2))
                return
            if isinstance(mod, nn.Conv2d):
                if pending["conv"] is not None:
                    raise ffi.Drnb200Error("conv %s is not followed by BatchNorm" % pending["key"])
                pending["conv"], pending["key"] = mod, key
                return
            if isinstance(mod, nn.BatchNorm2d):
                if pending["conv"] is None:
                    raise ffi.Drnb200Error("BatchNorm %s without a conv" % key)
                pending["bn"] = mod
                return
            if isinstance(mod, nn.ReLU):
                if pending["conv"] is not None:
                    flush_plain(pending.pop("bn"), True)
                return
            if isinstance(mod, nn.Sequential) or children:
                for name, child in children:
                    walk(child, key + "." + name)
                    # a conv+bn pair not followed by ReLU inside this container
                return
            raise ffi.Drnb200Error("unsupported module %s: %r" % (key, mod))

        walk(root, prefix)
        if pending["conv"] is not None:
            raise ffi.Drnb200Error("dangling conv %s" % pending["key"])
        if self.stem is None or not ops:
            raise ffi.Drnb200Error("could not find the DRN stem / conv stages")
        self.ops = ops
        # alternative launch list for wide frames: blocks with a stride-1 1x1 shortcut run as
        # [conv1] + [conv2 with the shortcut inside its K loop] instead of [conv1 | shortcut] + [conv2 + residual]
        # (same indices, all other ops shared).  Chosen per frame size by ops_for().
        self.ops_proj = None
        alt = list(ops)
        for i, op in enumerate(ops):
            if isinstance(op, FusedFirstConv) and ProjResidualConv.applicable(op.block):
                bkey = op.key[:-len(".conv1")]
                src = op.input_from
                pitch = self.stem[0].out_channels if src < 0 else ops[src].out_channels
                alt[i] = ConvLayer(bkey + ".conv1", op.block.conv1, op.block.bn1, True, input_from=src)
                alt[i + 1] = ProjResidualConv(bkey, op.block, input_from=i, proj_from=src, proj_pitch=pitch)
                self.ops_proj = alt

    def ops_for(self, H, W):
        """the launch list used for [*, 3, H, W] frames"""
        if self.ops_proj is None or self.conv_impl == ffi.IMPL_DIRECT or not self.proj_in_k:
            return self.ops
        shapes = {-1: (H, W)}
        for i, op in enumerate(self.ops_proj):
            src = op.input_from if op.input_from is not None else i - 1
            shapes[i] = op.out_hw(*shapes[src])
            if isinstance(op, ProjResidualConv) and shapes[i][1] <= 128:     # row-halo kernel: rows > 128 pixels
                return self.ops
        return self.ops_proj

    # ---- parameters -> device caches -----------------------------------------------------------
    def set_masks(self, mask_dict):
        """attach a Pruner.mask_dict (pruners/Pruner.py:13); None = derive liveness from zeros"""
        self.mask_dict = mask_dict

    def invalidate(self):
        """drop every derived cache on the next call (parameter writes through `.data` bypass the version counters)"""
        self.epoch += 1

    def _stamp(self):
        """cache-key component: the invalidate() epoch, plus (verify_weights) a content fingerprint of all parameters
        and buffers — one device reduction per tensor and ONE host synchronisation per call"""
        if not self.verify_weights:
            return (self.epoch,)
        sums = [t.detach().double().abs().sum() + t.detach().double().sum() * 0.5
                for t in list(self.m.parameters()) + list(self.m.buffers()) if t.is_floating_point()]
        if self.mask_dict is not None:
            sums += [m.detach().double().sum() for m in self.mask_dict.values()]
        by_dev = {}
        for v in sums:
            by_dev.setdefault(v.device, []).append(v)
        return (self.epoch,) + tuple(x for vs in by_dev.values() for x in torch.stack(vs).cpu().tolist())

    def refresh(self, device, ops=None):
        rebuilt = 0
        stamp = self._stamp()
        for op in (self.ops if ops is None else ops):
            rebuilt += bool(op.refresh(self.mask_dict, self.act_dtype, device, stamp))
        conv, bn, _ = self.stem
        stem_w, stem_wver = _effective_weight(conv)
        ver = (stem_wver, bn.weight._version, bn.bias._version, bn.running_mean._version,
               bn.running_var._version, str(device), stamp)
        if getattr(self, "_stem_version", None) != ver:
            inv = torch.rsqrt(bn.running_var.detach().to(device, torch.float32) + bn.eps)
            self.stem_scale = (bn.weight.detach().to(device, torch.float32) * inv).contiguous()
            self.stem_shift = (bn.bias.detach().to(device, torch.float32)
                               - bn.running_mean.detach().to(device, torch.float32) * self.stem_scale
                               ).contiguous()
            self.stem_w = stem_w.to(device, torch.float32).contiguous()
            self._stem_version = ver
            self._drop_stem_plans()
            rebuilt += 1
        seg = self.m.seg
        seg_w, seg_wver = _effective_weight(seg)
        hver = (seg_wver, seg.bias._version, self.act_dtype, str(device), stamp)
        if self.head_version != hver:
            lib = ffi.lib()
            for p in self.head_plans.values():
                lib.drnb200_head_plan_destroy(p)
            self.head_plans = {}
            self.seg_w = seg_w.to(device, torch.float32).reshape(seg.out_channels, -1).contiguous()
            self.seg_b = seg.bias.detach().to(device, torch.float32).contiguous()
            self.head_version = hver
            rebuilt += 1
        if rebuilt:
            self._graphs.clear()       # captured launches point into the plans / weights that were just replaced
        return rebuilt

    def set_ingest(self, mean, std, bgr=False):
        """normalisation of uint8 frames (data_transforms.py:109-125: (x/255 - mean) / std per channel)"""
        self.ingest = (tuple(float(m) for m in mean), tuple(float(v) for v in std), bool(bgr))
        self._lut = None
        self._graphs.clear()

    def _ingest_lut(self, device):
        if self._lut is None or self._lut[0] != (device, self.act_dtype):
            self._lut = ((device, self.act_dtype),
                         ingest_lut(self.ingest[0], self.ingest[1], self.act_dtype).to(device))
        return self._lut[1]

    def _drop_stem_plans(self):
        lib = ffi.lib()
        for p in self.stem_plans.values():
            lib.drnb200_stem_plan_destroy(p)
        self.stem_plans = {}

    def _stem_plan(self, N, H, W):
        k = (N, H, W, self.act_dtype)
        p = self.stem_plans.get(k)
        if p is None:
            hdl = C.c_void_p()
            ffi.check(ffi.lib().drnb200_stem_plan_create(
                C.byref(hdl), ffi.ptr(self.stem_w), ffi.ptr(self.stem_scale), ffi.ptr(self.stem_shift),
                N, H, W, self.stem[0].out_channels, self.act_dtype, ffi.stream_ptr()), "stem_plan_create")
            p = self.stem_plans[k] = hdl
        return p

    def _head_plan(self, N, h, w):
        k = (N, h, w)
        p = self.head_plans.get(k)
        if p is None:
            hdl = C.c_void_p()
            seg = self.m.seg
            ffi.check(ffi.lib().drnb200_head_plan_create(
                C.byref(hdl), N, h, w, seg.in_channels, seg.out_channels, self.act_dtype,
                ffi.ptr(self.seg_w), ffi.ptr(self.seg_b), ffi.stream_ptr()), "head_plan_create")
            if getattr(self.m, "use_torch_up", False):      # nn.UpsamplingBilinear2d(scale_factor=8), semantic_seg.py:144-145
                ffi.check(ffi.lib().drnb200_head_plan_set_upsample(hdl, 1), "head_plan_set_upsample")
            p = self.head_plans[k] = hdl
        return p

    # ---- statistics ----------------------------------------------------------------------------
    def mac_counts(self, N, H, W):
        """(dense MACs, unpruned-element MACs, live-tile MACs) of the conv stack incl. stem and seg,
        per forward of an [N,3,H,W] batch — numerators for tensor-pipe utilisation."""
        dense = live = tile = 0
        h, w = H, W
        sconv = self.stem[0]
        stem_macs = N * H * W * sconv.out_channels * 147
        dense += stem_macs; live += stem_macs; tile += stem_macs
        shapes = {-1: (H, W)}
        ops = self.last_ops if self.last_ops is not None and self.last_ops is self.ops_for(H, W) else self.ops
        for i, op in enumerate(ops):
            src = op.input_from if op.input_from is not None else i - 1
            ih, iw = shapes[src]
            oh, ow = op.out_hw(ih, iw)
            shapes[i] = (oh, ow)
            px = N * oh * ow
            dense += px * op.dense_macs_per_pixel()
            live += px * op.live_elems
            tile += px * op.n_live * op.tile_o * op.tile_ci
        oh, ow = shapes[len(self.ops) - 1]
        seg = N * oh * ow * self.m.seg.in_channels * self.m.seg.out_channels
        return dense + seg, live + seg, tile + seg

    # ---- forward ---------------------------------------------------------------------------------
    def _check_input(self, x):
        u8 = x.dtype == torch.uint8
        if u8:
            # frame ingest fused into the stem: uint8 HWC frames (cv2 / PIL order), normalisation via set_ingest()
            if not (x.is_cuda and x.dim() == 4 and x.shape[3] == 3):
                raise ffi.Drnb200Error("uint8 input must be a CUDA tensor [N,H,W,3] (got %s on %s); "
                                       "there is no CPU path" % (tuple(x.shape), x.device))
            if self.ingest is None:
                raise ffi.Drnb200Error("uint8 frames need set_ingest(mean, std) first (info.json of the dataset)")
            if self.stem_impl == "direct":
                raise ffi.Drnb200Error("the CUDA-core stem has no uint8 ingest")
            N, H, W, _ = x.shape
            if W % 16:
                raise ffi.Drnb200Error("uint8 ingest needs W %% 16 == 0 (got %d)" % W)
        else:
            if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
                raise ffi.Drnb200Error("input must be a float32 CUDA tensor [N,3,H,W] (got %s %s on %s); "
                                       "there is no CPU path" % (tuple(x.shape), x.dtype, x.device))
            N, _, H, W = x.shape
        # any H; W % 4 == 0 (the float32 frame rows must be 16-byte multiples for TMA).  Like the reference the
        # stride-2 stages round up, so the label map is 8*ceil(H/8) x 8*ceil(W/8) (300x300 -> 304x304,
        # seg_video_old.py:127)
        if W % 4 or H < 8 or W < 8:
            raise ffi.Drnb200Error("W must be a multiple of 4 and H, W >= 8 (got %dx%d)" % (H, W))
        return u8, N, H, W

    def run(self, x, want_labels=True, want_logprob=False, want_seg=False, timings=None, taps=None):
        """launch the whole path on the current stream.  `timings`, if a list, receives
        (name, start_event, end_event) per launch group (CUDA events on the launching stream).
        `taps`, if a dict, receives a float32 NCHW copy of every stored activation keyed by the state_dict prefix of
        the conv that produced it (the keys of the oracle's taps; a fused [conv1 | downsample] launch yields both)."""
        def timed(name):
            return _Timed(name, timings)

        tdt = torch.bfloat16 if self.act_dtype == ffi.BF16 else torch.float16

        def tap(op_keys, buf, n, oh, ow, ch):
            if taps is None:
                return
            t = buf[:n * oh * ow * ch].view(tdt).view(n, oh, ow, ch).float().permute(0, 3, 1, 2)
            step = ch // len(op_keys)
            for j, k in enumerate(op_keys):
                taps[k] = t[:, j * step:(j + 1) * step].contiguous()

        u8, N, H, W = self._check_input(x)
        # every launch, allocation and stream query below must target the input's device, whatever the caller's
        # current device is (frames on cuda:1 while cuda:0 is current, DataParallel worker threads)
        with torch.cuda.device(x.device):
            return self._run(x, u8, N, H, W, want_labels, want_logprob, want_seg, timed, tap)

    GRAPH_SLOTS = 8

    def run_graphed(self, x, static_output=False):
        """label map of `x` through a captured CUDA graph of the whole launch list (stem .. fused head).

        The 22-39 launches of a forward cost the Python host 0.3-0.5 ms; at batch 8 of 1024x2048 frames the GPU needs
        2.8 ms and hides that, at batch 1 (the reference's evaluation loop, semantic_seg.py test(): one frame per step)
        or on small video frames it does not.  A graph replays the same kernels, tensor maps and programmatic
        dependencies with one launch.  Graphs are keyed by the INPUT BUFFER (address, shape, dtype) because tensor maps
        are kernel parameters: feed frames through a fixed set of device buffers (FramePipeline does).  First call with
        a new key: one eager forward (creates the plans), then the capture; at most GRAPH_SLOTS graphs are kept (LRU),
        all in one memory pool.  Any rebuild of the derived caches (weights, masks, BN, ingest) drops every graph.
        static_output=True returns the graph's own label tensor, overwritten by the next replay of the same graph."""
        u8, N, H, W = self._check_input(x)
        if self.verify_weights or not x.is_contiguous():
            return self.run(x, want_labels=True)[0]      # content fingerprints synchronise: no capture
        with torch.cuda.device(x.device):
            key = (x.data_ptr(), tuple(x.shape), x.dtype)
            ent = self._graphs.get(key) if self.last_ops is not None else None
            if ent is not None:
                try:
                    if self.refresh(x.device, self.last_ops):
                        ent = None                       # refresh() dropped the graphs
                except ProjFallback:
                    ent = None
            if ent is None:
                labels = self.run(x, want_labels=True)[0]            # eager: plans, attributes, fallbacks
                if self._graph_pool is None:
                    self._graph_pool = torch.cuda.graph_pool_handle()
                g = torch.cuda.CUDAGraph()
                timed = lambda name: _Timed(name, None)
                with torch.cuda.graph(g, pool=self._graph_pool, capture_error_mode="thread_local"):
                    out = self._run(x, u8, N, H, W, True, False, False, timed, lambda *a: None)[0]
                self._graphs[key] = (g, out)
                while len(self._graphs) > self.GRAPH_SLOTS:
                    self._graphs.popitem(last=False)
                return labels                                        # the eager result of this very call
            self._graphs.move_to_end(key)
            g, out = ent
            g.replay()
            return out if static_output else out.clone()

    def _run(self, x, u8, N, H, W, want_labels, want_logprob, want_seg, timed, tap):
        x = x.contiguous()
        dev = x.device
        lib = ffi.lib()
        ops = self.ops_for(H, W)
        try:
            self.refresh(dev, ops)
        except ProjFallback:
            # ill-conditioned BatchNorm fold: keep the shortcut as its own output.  The alternative list's own ops
            # (everything not shared with self.ops) may already hold plans: destroy them before dropping the list
            shared = set(map(id, self.ops))
            for op in self.ops_proj:
                if id(op) not in shared:
                    op.destroy_plans()
            self.ops_proj = None
            ops = self.ops
            self.refresh(dev, ops)
        self.last_ops = ops
        st = ffi.stream_ptr()
        adt = self.act_dtype
        launches = 0
        # last conv hands float32? (kept 16-bit in round 1: the head GEMM consumes act_dtype)
        outs = {}
        shapes = {-1: (H, W)}
        # liveness of buffers for pooling
        last_use = {}
        for i, op in enumerate(ops):
            src = op.input_from if op.input_from is not None else i - 1
            last_use[src] = i
            if op.residual_from is not None:
                last_use[op.residual_from] = i
        last_use[len(ops) - 1] = len(ops)      # consumed by the head
        pool = {}

        def take(nelem):
            lst = pool.get(nelem)
            if lst:
                return lst.pop()
            return torch.empty(nelem, dtype=torch.int16, device=dev)

        def give(idx):
            t = outs.pop(idx, None)
            if t is not None:
                pool.setdefault(t.numel(), []).append(t)

        sconv = self.stem[0]
        c0 = sconv.out_channels
        y = take(N * H * W * c0)
        with timed("stem"):
            if self.stem_impl == "direct":
                ffi.check(lib.drnb200_stem_forward(ffi.ptr(x), ffi.ptr(self.stem_w), ffi.ptr(self.stem_scale),
                                                   ffi.ptr(self.stem_shift), N, H, W, c0, adt, ffi.ptr(y), st),
                          "stem_forward")
            elif u8:
                ffi.check(lib.drnb200_stem_plan_forward_u8(self._stem_plan(N, H, W), ffi.ptr(x),
                                                           ffi.ptr(self._ingest_lut(dev)), int(self.ingest[2]),
                                                           ffi.ptr(y), st), "stem_plan_forward_u8")
            else:
                ffi.check(lib.drnb200_stem_plan_forward(self._stem_plan(N, H, W), ffi.ptr(x), ffi.ptr(y), st),
                          "stem_plan_forward")
        launches += 1
        outs[-1] = y
        tap([self.stem[2]], y, N, H, W, c0)
        for i, op in enumerate(ops):
            src = op.input_from if op.input_from is not None else i - 1
            ih, iw = shapes[src]
            oh, ow = op.out_hw(ih, iw)
            shapes[i] = (oh, ow)
            plan = op.plan(N, ih, iw, adt, self.conv_impl)
            yo = take(N * oh * ow * op.out_channels)
            res = outs[op.residual_from] if op.residual_from is not None else None
            with timed(op.key):
                ffi.check(lib.drnb200_conv_forward(plan, ffi.ptr(outs[src]), ffi.ptr(res), ffi.ptr(yo), st),
                          "conv_forward(%s)" % op.key)
            launches += 1
            outs[i] = yo
            tap(op.keys if isinstance(op, FusedFirstConv) else op.keys[:1], yo, N, oh, ow, op.out_channels)
            for j in [k for k, last in last_use.items() if last == i]:
                give(j)
        last = len(ops) - 1
        h8, w8 = shapes[last]
        hp = self._head_plan(N, h8, w8)
        classes = self.m.seg.out_channels
        labels = torch.empty((N, 8 * h8, 8 * w8), dtype=torch.uint8, device=dev) if want_labels else None
        logprob = torch.empty((N, classes, 8 * h8, 8 * w8), dtype=torch.float32, device=dev) \
            if want_logprob else None
        seg = torch.empty((N, classes, h8, w8), dtype=torch.float32, device=dev) if want_seg else None
        with timed("head"):
            ffi.check(lib.drnb200_head_forward(hp, ffi.ptr(outs[last]), ffi.ptr(labels), ffi.ptr(seg),
                                               ffi.ptr(logprob), st), "head_forward")
        if want_labels and not want_logprob and not want_seg and lib.drnb200_head_plan_fused(hp):
            launches += 1                    # classifier + upsample + argmax in one kernel
        else:
            launches += 1 + int(want_labels or want_logprob) + int(want_seg)
        self.launches_per_forward = launches
        return labels, logprob, seg

    def close(self):
        self._graphs.clear()
        lib = ffi.lib()
        for op in self.ops + (self.ops_proj or []):
            op.destroy_plans()
        for p in self.head_plans.values():
            lib.drnb200_head_plan_destroy(p)
        self.head_plans = {}
        self._drop_stem_plans()
