"""ORACLE — test infrastructure, not product code.

CPU fp32 restatement (torch functional ops + numpy) of the reference's inference path
``DRNSeg.forward`` -> ``torch.max(final, 1)`` -> ``fast_hist`` (semantic_seg.py:126-164, :444-455,
:293-300; drn.py:32-259).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package; the product (video-seg-model-compress_b200/) never does.

Pinning: ``tests/golden/gen_golden.py`` imports the *real* reference from /root/reference in the build
container, runs it on seeded weights/inputs (oracle/recipe.py) and commits the outputs as fixtures under
``tests/golden/``; ``tests/test_oracle_pinned.py`` checks this restatement against those fixtures and against
the one golden file the reference ships (pruners/block_test.txt).  Parity of the CUDA path is then checked
against this oracle on the GPU box, where /root/reference does not exist.

The network structure is recovered from the state_dict keys alone, so the oracle works for every
DRN-C/D variant and both key flavours (``layer.`` of semantic_seg.py, ``base.`` of seg_video.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm2d default, drn.py:7


def _bn(x, sd, key):
    """BatchNorm2d in eval mode (drn.py:40,44,135,208)"""
    scale = sd[key + ".weight"] / torch.sqrt(sd[key + ".running_var"] + BN_EPS)
    shift = sd[key + ".bias"] - sd[key + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def backbone_forward(sd, x, prefix="layer", taps=None, raw=None, quant=None):
    """Run the DRN backbone (everything in DRNSeg.layer) on float32 NCHW `x`.

    `quant`, if a callable ``quant(role, key, tensor) -> tensor``, is applied wherever the CUDA path STORES a
    tensor in 16 bits (role "input": the frame; "act": a conv+bn(+res)+relu output; "shortcut": a downsample
    output) — a storage-precision emulation for the error-budget analysis (tests/precision_budget.py); None
    (the default) leaves this the plain fp32 restatement.

    `taps`, if a dict, receives every conv+bn(+res)(+relu) output keyed by the conv's state_dict prefix
    (NCHW float32) — used by the per-layer parity tests.  `raw`, if a dict, receives the bare nn.Conv2d
    outputs (what a forward hook on the reference's conv modules sees) for pinning against fixtures."""
    def conv2d(key, inp, **kw):
        out = F.conv2d(inp, sd[key + ".weight"], None, **kw)
        if raw is not None:
            raw[key] = out
        return out

    def store(role, key, t):
        return t if quant is None else quant(role, key, t)

    x = store("input", "frame", x)
    keys = [k for k in sd.keys() if k.startswith(prefix + ".")]
    top = sorted({int(k.split(".")[1]) for k in keys})
    arch_d = (prefix + ".0.0.weight") in sd           # arch D: child 0 is Sequential(conv7, bn, relu)
    # map top-level child index -> DRN stage number (layer<stage>)
    # arch D children: layer0..layer8 -> indices 0..8 ; arch C: conv1, bn1, relu, layer1..layer8 -> 0,1,(2),3..10
    if arch_d:
        x = conv2d(prefix + ".0.0", x, stride=1, padding=3)      # drn.py:132-137
        x = store("act", prefix + ".0.0", F.relu(_bn(x, sd, prefix + ".0.1")))
        if taps is not None:
            taps[prefix + ".0.0"] = x
        stage_of = {i: i for i in top if i >= 1}
    else:
        x = conv2d(prefix + ".0", x, stride=1, padding=3)        # drn.py:123-127
        x = store("act", prefix + ".0", F.relu(_bn(x, sd, prefix + ".1")))
        if taps is not None:
            taps[prefix + ".0"] = x
        stage_of = {i: i - 2 for i in top if i >= 3}

    for idx in sorted(stage_of):
        stage = stage_of[idx]
        base = "%s.%d" % (prefix, idx)
        stride = 2 if stage in (2, 3, 4) else 1                                     # drn.py:140-145
        dil = {5: 2, 6: 4, 7: 2}.get(stage, 1)                                      # drn.py:146-160
        sub = sorted({int(k.split(".")[2]) for k in keys if k.startswith(base + ".")})
        is_block = any(k.startswith(base + ".0.conv1.") for k in keys)
        if not is_block:
            # _make_conv_layers: [conv3x3(stride on first), BN, ReLU] * n  (drn.py:201-211)
            convs = [i for i in sub if (base + ".%d.weight" % i) in sd and sd[base + ".%d.weight" % i].dim() == 4]
            for n, ci in enumerate(convs):
                ck = base + ".%d" % ci
                x = conv2d(ck, x, stride=stride if n == 0 else 1, padding=dil, dilation=dil)
                x = store("act", ck, F.relu(_bn(x, sd, base + ".%d" % (ci + 1))))
                if taps is not None:
                    taps[ck] = x
            continue
        # residual stage (_make_layer, drn.py:177-199); C-arch layer7/8 have residual=False (drn.py:153-158)
        no_residual = (not arch_d) and stage in (7, 8)
        for bi in sub:
            bk = base + ".%d" % bi
            bstride = stride if bi == 0 else 1
            # first block: dilation (1,1) if dil==1 else (dil//2 if new_level else dil, dil);
            # layers 5..8 are built with new_level=False, so both convs use `dil` (drn.py:188-196)
            d1 = d2 = dil
            bottleneck = (bk + ".conv3.weight") in sd
            resid = x
            if bottleneck:                                                          # drn.py:86-106
                out = store("act", bk + ".conv1", F.relu(_bn(conv2d(bk + ".conv1", x), sd, bk + ".bn1")))
                if taps is not None:
                    taps[bk + ".conv1"] = out
                out = store("act", bk + ".conv2", F.relu(_bn(
                    conv2d(bk + ".conv2", out, stride=bstride, padding=d2, dilation=d2), sd, bk + ".bn2")))
                if taps is not None:
                    taps[bk + ".conv2"] = out
                out = _bn(conv2d(bk + ".conv3", out), sd, bk + ".bn3")
                last = bk + ".conv3"
            else:                                                                   # drn.py:49-65
                out = store("act", bk + ".conv1", F.relu(_bn(
                    conv2d(bk + ".conv1", x, stride=bstride, padding=d1, dilation=d1), sd, bk + ".bn1")))
                if taps is not None:
                    taps[bk + ".conv1"] = out
                out = _bn(conv2d(bk + ".conv2", out, padding=d2, dilation=d2), sd, bk + ".bn2")
                last = bk + ".conv2"
            if (bk + ".downsample.0.weight") in sd:                                 # drn.py:181-186
                resid = store("shortcut", bk + ".downsample.0",
                              _bn(conv2d(bk + ".downsample.0", x, stride=bstride), sd, bk + ".downsample.1"))
                if taps is not None:
                    taps[bk + ".downsample.0"] = resid
            if not no_residual:
                out = out + resid
            x = store("act", last, F.relu(out))
            if taps is not None:
                taps[last] = x
    return x


def up_weights(k=16):
    """fill_up_weights (semantic_seg.py:115-124): separable bilinear taps; for k=16: w[i] = 1-|2i-15|/16"""
    f = math.ceil(k / 2)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    return torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=torch.float64)


def head_forward(sd, feat, up_weight=None, use_torch_up=False):
    """seg 1x1 -> grouped ConvTranspose2d(k16,s8,p4) -> LogSoftmax(dim=1)   (semantic_seg.py:154-158);
    `use_torch_up`: nn.UpsamplingBilinear2d(scale_factor=8) instead (semantic_seg.py:144-145)"""
    seg = F.conv2d(feat, sd["seg.weight"], sd["seg.bias"])
    classes = seg.shape[1]
    if use_torch_up:
        y = F.interpolate(seg, scale_factor=8, mode="bilinear", align_corners=True)
        return F.log_softmax(y, dim=1), seg
    if up_weight is None:
        up_weight = sd.get("up.weight")
    if up_weight is None:
        ax = up_weights(16)
        up_weight = torch.outer(ax, ax).to(torch.float32).expand(classes, 1, 16, 16).contiguous()
    y = F.conv_transpose2d(seg, up_weight, None, stride=8, padding=4, groups=classes)
    return F.log_softmax(y, dim=1), seg


@torch.no_grad()
def drnseg_forward(sd, x, prefix=None, taps=None, raw=None, quant=None, use_torch_up=False):
    """DRNSeg.forward(x) -> (logprob [N,C,H,W], seg_logits [N,C,H/8,W/8])  fp32 on CPU"""
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items() if torch.is_floating_point(v)}
    if prefix is None:
        prefix = "layer" if any(k.startswith("layer.") for k in sd) else "base"
    feat = backbone_forward(sd, x.detach().to("cpu", torch.float32), prefix, taps, raw, quant)
    return head_forward(sd, feat, use_torch_up=use_torch_up)


def predict_labels(sd, x, prefix=None):
    """``_, pred = torch.max(final, 1)`` (semantic_seg.py:445): int64 labels, first maximum wins"""
    final, _ = drnseg_forward(sd, x, prefix)
    return torch.max(final, 1)[1]


# -------------------------------------------------------------------------------- metrics (numpy)

def fast_hist(pred, label, n):
    """semantic_seg.py:293-296"""
    k = (label >= 0) & (label < n)
    return np.bincount(n * label[k].astype(int) + pred[k], minlength=n ** 2).reshape(n, n)


def per_class_iu(hist):
    """semantic_seg.py:299-300"""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))


def miou(hist):
    """``round(np.nanmean(per_class_iu(hist) * 100), 2)`` (semantic_seg.py:466-468)"""
    return round(float(np.nanmean(per_class_iu(hist) * 100)), 2)
