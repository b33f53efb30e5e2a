"""``DRNSeg`` — drop-in mirror of the reference's segmentation module, running on libdrnb200.so.

Reference interface kept verbatim (semantic_seg.py:126-164; the ``base`` flavour of seg_video.py:70-108
via ``backbone_attr='base'``):

    DRNSeg(model_name, classes, pretrained_model=None, pretrained=True, use_torch_up=False)
    forward(x: float32 [N,3,H,W]) -> (log_softmax(up(seg(layer(x)))) [N,classes,H,W],
                                      seg(layer(x))                [N,classes,H/8,W/8])
    optim_parameters(); sub-modules layer|base, seg, up, softmax; identical state_dict() keys.

So ``final = model(x)[0]; _, pred = torch.max(final, 1)`` of ``test()`` (semantic_seg.py:444-445) and
``FrameCapture`` (seg_video.py:161-164) work unchanged.  Added for the measured fast path:

    predict(x) -> uint8 [N,H,W]      labels straight from the fused head, no logits materialised
    set_pruner(pruner) / set_masks(mask_dict)   hand the Pruner.mask_dict to the tile-list builder
    prepare()                        force the (tile list, packed weights) cache to be rebuilt now
"""
import math

import torch
import torch.nn as nn

from . import drn as _drn
from . import ffi
from .engine import Engine


def fill_up_weights(up):
    """bilinear kernel of the grouped ConvTranspose2d: w[i,j] = (1-|i/f-c|)(1-|j/f-c|)
    with f = ceil(k/2), c = (2f-1-f%2)/(2f)   (semantic_seg.py:115-124)."""
    w = up.weight.data
    k = w.size(2)
    f = math.ceil(k / 2)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    ax = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=torch.float64)
    w[:, 0] = torch.outer(ax, ax[: w.size(3)]).to(w.dtype)


class DRNSeg(nn.Module):
    def __init__(self, model_name, classes, pretrained_model=None, pretrained=True,
                 use_torch_up=False, backbone_attr="layer", act_dtype="bf16"):
        super().__init__()
        model = _drn.build(model_name, pretrained=pretrained, num_classes=1000)
        if pretrained_model is not None:
            model.load_state_dict(pretrained_model)
        if backbone_attr not in ("layer", "base"):
            raise ValueError("backbone_attr must be 'layer' (semantic_seg.py) or 'base' (seg_video.py)")
        # drop avgpool + fc, exactly children()[:-2]  (semantic_seg.py:135)
        setattr(self, backbone_attr, nn.Sequential(*list(model.children())[:-2]))
        self._backbone_attr = backbone_attr

        self.seg = nn.Conv2d(model.out_dim, classes, kernel_size=1, bias=True)
        self.softmax = nn.LogSoftmax(dim=1)          # reference: implicit dim -> 1 for 4-D input
        fan_out = self.seg.kernel_size[0] * self.seg.kernel_size[1] * self.seg.out_channels
        self.seg.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        self.seg.bias.data.zero_()
        self.use_torch_up = bool(use_torch_up)
        if use_torch_up:
            self.up = nn.UpsamplingBilinear2d(scale_factor=8)
        else:
            up = nn.ConvTranspose2d(classes, classes, 16, stride=8, padding=4, output_padding=0,
                                    groups=classes, bias=False)
            fill_up_weights(up)
            up.weight.requires_grad = False
            self.up = up
        self._act_dtype = act_dtype
        self._engine = None
        self._mask_dict = None
        self._ingest = None

    # ---- reference API ---------------------------------------------------------------------------
    def forward(self, x):
        labels, logprob, seg = self._eng().run(x, want_labels=False, want_logprob=True, want_seg=True)
        return logprob, seg

    def optim_parameters(self, memo=None):
        for param in getattr(self, self._backbone_attr).parameters():
            yield param
        for param in self.seg.parameters():
            yield param

    # ---- fast path / mask plumbing ---------------------------------------------------------------
    @torch.no_grad()
    def predict(self, x):
        """uint8 label map [N,H,W] == torch.max(model(x)[0], 1)[1] of the reference, without ever
        writing the [N,classes,H,W] logits."""
        return self._eng().run(x, want_labels=True)[0]

    def set_ingest(self, mean, std, bgr=False):
        """enable uint8 HWC frames [N,H,W,3] as input of forward()/predict(): ToTensorVideoImage + Normalize
        (data_transforms.py:109-125, :256-281; mean/std from info.json) are fused into the stem kernel."""
        self._ingest = (mean, std, bgr)
        if self._engine is not None:
            self._engine.set_ingest(mean, std, bgr)
        return self

    def set_masks(self, mask_dict):
        self._mask_dict = mask_dict
        if self._engine is not None:
            self._engine.set_masks(mask_dict)
        return self

    def set_pruner(self, pruner):
        return self.set_masks(None if pruner is None else pruner.mask_dict)

    def set_act_dtype(self, act_dtype):
        if act_dtype != self._act_dtype:
            self._act_dtype = act_dtype
            if self._engine is not None:
                self._engine.close()
                self._engine = None
        return self

    def prepare(self, device=None):
        dev = device or next(self.parameters()).device
        return self._eng().refresh(torch.device(dev))

    def engine(self):
        return self._eng()

    def _eng(self):
        if self.use_torch_up:
            raise ffi.Drnb200Error("use_torch_up=True (UpsamplingBilinear2d, align_corners) is not part "
                                   "of the accelerated path; the reference's default is the fixed "
                                   "ConvTranspose2d (semantic_seg.py:147-152)")
        up = self.up.weight
        if tuple(up.shape[2:]) != (16, 16) or self.up.stride != (8, 8) or self.up.padding != (4, 4):
            raise ffi.Drnb200Error("`up` must be ConvTranspose2d(k=16, s=8, p=4)")
        if self._engine is None:
            ffi.lib()                      # fail loudly if the CUDA library is missing
            self._engine = Engine(self, act_dtype=self._act_dtype)
            self._engine.set_masks(self._mask_dict)
            if self._ingest is not None:
                self._engine.set_ingest(*self._ingest)
        return self._engine

    def __del__(self):
        try:
            if self._engine is not None:
                self._engine.close()
        except Exception:
            pass
