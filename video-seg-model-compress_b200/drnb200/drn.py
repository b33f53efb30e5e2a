"""Parameter containers for the DRN backbones — the host-side mirror of the reference's ``drn.py``.

These modules exist so that ``state_dict()`` keys, parameter shapes and initialisation match the
reference exactly (drn.py:109-259 ``DRN``; :32-65 ``BasicBlock``; :68-106 ``Bottleneck``; factories
:361-397), which is what the reference's pruners (``model.state_dict()[key]``) and checkpoints
(``load_state_dict``) rely on.  They are *containers*: the arithmetic of the inference path is done by
the CUDA engine in :mod:`drnb200.engine`, which walks these modules; calling one of them directly
(``DRN.forward``) is not part of the accelerated path and raises.
"""
import math

import torch.nn as nn

# channel plan shared by every DRN-C / DRN-D variant (drn.py:112)
_CHANNELS = (16, 32, 64, 128, 256, 512, 512, 512)


def _conv_bn_relu_stack(cin, cout, count, stride=1, dilation=1):
    """`count` x [3x3 conv (stride on the first), BN, ReLU] as one nn.Sequential (drn.py:201-211)."""
    mods = []
    for i in range(count):
        mods += [nn.Conv2d(cin, cout, 3, stride=stride if i == 0 else 1, padding=dilation,
                           dilation=dilation, bias=False),
                 nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
        cin = cout
    return nn.Sequential(*mods)


class BasicBlock(nn.Module):
    """conv3x3-BN-ReLU-conv3x3-BN (+identity / 1x1 projection) - ReLU   (drn.py:32-65)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=(1, 1), residual=True):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=dilation[0],
                               dilation=dilation[0], bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=dilation[1], dilation=dilation[1], bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride
        self.residual = residual


class Bottleneck(nn.Module):
    """1x1-BN-ReLU-3x3(dil, stride)-BN-ReLU-1x1(x4)-BN (+residual) - ReLU   (drn.py:68-106)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=(1, 1), residual=True):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=dilation[1],
                               dilation=dilation[1], bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class DRN(nn.Module):
    """Dilated residual network, architectures 'C' and 'D' (drn.py:109-175)."""

    def __init__(self, block, layers, num_classes=1000, channels=_CHANNELS, out_map=False,
                 out_middle=False, pool_size=28, arch="D"):
        super().__init__()
        if arch not in ("C", "D"):
            raise ValueError("arch must be 'C' or 'D'")
        ch = channels
        self.inplanes = ch[0]
        self.out_map, self.out_middle, self.out_dim, self.arch = out_map, out_middle, ch[-1], arch

        if arch == "C":
            self.conv1 = nn.Conv2d(3, ch[0], 7, stride=1, padding=3, bias=False)
            self.bn1 = nn.BatchNorm2d(ch[0])
            self.relu = nn.ReLU(inplace=True)
            self.layer1 = self._residual_stage(BasicBlock, ch[0], layers[0], stride=1)
            self.layer2 = self._residual_stage(BasicBlock, ch[1], layers[1], stride=2)
        else:
            self.layer0 = nn.Sequential(nn.Conv2d(3, ch[0], 7, stride=1, padding=3, bias=False),
                                        nn.BatchNorm2d(ch[0]), nn.ReLU(inplace=True))
            self.layer1 = self._plain_stage(ch[0], layers[0], stride=1)
            self.layer2 = self._plain_stage(ch[1], layers[1], stride=2)

        self.layer3 = self._residual_stage(block, ch[2], layers[2], stride=2)
        self.layer4 = self._residual_stage(block, ch[3], layers[3], stride=2)
        self.layer5 = self._residual_stage(block, ch[4], layers[4], dilation=2, new_level=False)
        self.layer6 = None if layers[5] == 0 else \
            self._residual_stage(block, ch[5], layers[5], dilation=4, new_level=False)
        if arch == "C":
            self.layer7 = None if layers[6] == 0 else \
                self._residual_stage(BasicBlock, ch[6], layers[6], dilation=2, new_level=False,
                                     residual=False)
            self.layer8 = None if layers[7] == 0 else \
                self._residual_stage(BasicBlock, ch[7], layers[7], dilation=1, new_level=False,
                                     residual=False)
        else:
            self.layer7 = None if layers[6] == 0 else self._plain_stage(ch[6], layers[6], dilation=2)
            self.layer8 = None if layers[7] == 0 else self._plain_stage(ch[7], layers[7], dilation=1)

        if num_classes > 0:
            self.avgpool = nn.AvgPool2d(pool_size)
            self.fc = nn.Conv2d(self.out_dim, num_classes, 1, stride=1, padding=0, bias=True)

        # He initialisation over fan-out, BN to identity (drn.py:169-175)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _residual_stage(self, block, planes, blocks, stride=1, dilation=1, new_level=True,
                        residual=True):
        assert dilation == 1 or dilation % 2 == 0
        out_ch = planes * block.expansion
        proj = None
        if stride != 1 or self.inplanes != out_ch:
            proj = nn.Sequential(nn.Conv2d(self.inplanes, out_ch, 1, stride=stride, bias=False),
                                 nn.BatchNorm2d(out_ch))
        first_dil = (1, 1) if dilation == 1 else \
            (dilation // 2 if new_level else dilation, dilation)
        stage = [block(self.inplanes, planes, stride, proj, dilation=first_dil, residual=residual)]
        self.inplanes = out_ch
        stage += [block(self.inplanes, planes, residual=residual, dilation=(dilation, dilation))
                  for _ in range(1, blocks)]
        return nn.Sequential(*stage)

    def _plain_stage(self, channels, convs, stride=1, dilation=1):
        stage = _conv_bn_relu_stack(self.inplanes, channels, convs, stride, dilation)
        self.inplanes = channels
        return stage

    def forward(self, x):
        raise RuntimeError("drnb200.drn.DRN is a parameter container; run it through "
                           "drnb200.DRNSeg (CUDA engine). There is no eager fallback.")


_VARIANTS = {
    # name: (arch, block, layers)                                          drn.py:318-397
    "drn_c_26": ("C", BasicBlock, [1, 1, 2, 2, 2, 2, 1, 1]),
    "drn_c_42": ("C", BasicBlock, [1, 1, 3, 4, 6, 3, 1, 1]),
    "drn_c_58": ("C", Bottleneck, [1, 1, 3, 4, 6, 3, 1, 1]),
    "drn_d_22": ("D", BasicBlock, [1, 1, 2, 2, 2, 2, 1, 1]),
    "drn_d_24": ("D", BasicBlock, [1, 1, 2, 2, 2, 2, 2, 2]),
    "drn_d_38": ("D", BasicBlock, [1, 1, 3, 4, 6, 3, 1, 1]),
    "drn_d_40": ("D", BasicBlock, [1, 1, 3, 4, 6, 3, 2, 2]),
    "drn_d_54": ("D", Bottleneck, [1, 1, 3, 4, 6, 3, 1, 1]),
    "drn_d_56": ("D", Bottleneck, [1, 1, 3, 4, 6, 3, 2, 2]),
    "drn_d_105": ("D", Bottleneck, [1, 1, 3, 4, 23, 3, 1, 1]),
    "drn_d_107": ("D", Bottleneck, [1, 1, 3, 4, 23, 3, 2, 2]),
}


def build(model_name, pretrained=False, **kwargs):
    """``drn.__dict__[model_name](pretrained=..., **kwargs)`` of the reference (semantic_seg.py:130-131).

    ``pretrained=True`` downloads ImageNet weights in the reference (drn.py:364); this environment has
    no network access, so it is rejected explicitly instead of failing inside urllib."""
    if model_name not in _VARIANTS:
        raise KeyError("unknown DRN variant %r (have: %s)" % (model_name, ", ".join(sorted(_VARIANTS))))
    if pretrained:
        raise RuntimeError("pretrained=True needs the model-zoo download of the reference (drn.py:13-24); "
                           "pass pretrained=False and load a state_dict instead")
    arch, block, layers = _VARIANTS[model_name]
    return DRN(block, layers, arch=arch, **kwargs)


def _factory(name):
    def make(pretrained=False, **kwargs):
        return build(name, pretrained=pretrained, **kwargs)
    make.__name__ = name
    return make


for _name in _VARIANTS:
    globals()[_name] = _factory(_name)
