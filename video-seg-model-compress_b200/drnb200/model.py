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
    prepare()                        build the (tile list, packed weights) cache now
    invalidate()                     drop that cache: REQUIRED after writing parameters through ``.data``

Cache coherence.  The derived device cache is keyed on the parameters' autograd version counters, which every
in-place tensor op bumps (``state_dict()[k] *= mask`` of Pruner.apply_masks, ``load_state_dict``, optimizers
under ``no_grad``).  Writes through ``param.data`` do NOT bump the counter; call ``invalidate()`` after them, or
construct with ``verify_weights=True`` to re-fingerprint every parameter on every call (slow, debugging aid).

Devices.  One engine per CUDA device is kept (keyed by the input's device) and every call runs under
``torch.cuda.device(x.device)``, so ``nn.DataParallel(DRNSeg)`` (semantic_seg.py:812) works: replicas share
the original module's per-device engines, which read the original parameters.

Inference only: there is no autograd graph and BatchNorm always uses its running statistics; ``forward`` raises in
training mode with gradients enabled instead of silently returning eval-mode results.
"""
import math
import threading
import warnings
import weakref

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
    act_dtype_default = "fp16"       # the 16-bit storage type that passes the label gate

    def __init__(self, model_name, classes, pretrained_model=None, pretrained=True,
                 use_torch_up=False, backbone_attr="layer", act_dtype="fp16", verify_weights=False):
        """`act_dtype`: 16-bit activation storage, "fp16" (default: passes the label gate) or
        "bf16".  `pretrained=True` is the reference's signature default (ImageNet weights from its model zoo,
        drn.py:13-24); there is no download here, so it warns and keeps the random initialisation — load a
        checkpoint with load_state_dict / drnb200.checkpoint afterwards."""
        super().__init__()
        if pretrained:
            warnings.warn("DRNSeg(pretrained=True): the reference's model-zoo download (drn.py:13-24) is not available; "
                          "the backbone keeps its random initialisation - load a state_dict afterwards", stacklevel=2)
        model = _drn.build(model_name, pretrained=False, num_classes=1000)
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
        self._verify_weights = bool(verify_weights)
        # per-device engines; the dict, the lock and the weak reference to THIS module are shared by DataParallel
        # replicas (their __dict__ is a shallow copy), so replicas run on the original module's engines
        self._engines = {}
        self._engine_lock = threading.Lock()
        object.__setattr__(self, "_origin", weakref.ref(self))
        self._mask_dict = None
        self._ingest = None
        self.use_graphs = False

    # ---- reference API ---------------------------------------------------------------------------
    def forward(self, x):
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.optim_parameters()):
            raise ffi.Drnb200Error("drnb200.DRNSeg is an inference path (no autograd graph, BatchNorm running statistics); "
                                   "call .eval() or run under torch.no_grad() - train with the reference module")
        labels, logprob, seg = self._eng(x).run(x, want_labels=False, want_logprob=True, want_seg=True)
        return logprob, seg

    def optim_parameters(self, memo=None):
        for param in getattr(self, self._backbone_attr).parameters():
            yield param
        for param in self.seg.parameters():
            yield param

    # ---- fast path / mask plumbing ---------------------------------------------------------------
    @torch.no_grad()
    def predict(self, x, static_output=False):
        """uint8 label map [N,H,W] == torch.max(model(x)[0], 1)[1] of the reference, without ever
        writing the [N,classes,H,W] logits.  With ``use_graphs`` (see enable_graphs) the launch list is replayed as
        one CUDA graph per input buffer; static_output=True then returns the graph's own output tensor."""
        if self.use_graphs:
            return self._eng(x).run_graphed(x, static_output=static_output)
        return self._eng(x).run(x, want_labels=True)[0]

    def enable_graphs(self, on=True):
        """replay predict() as a captured CUDA graph (Engine.run_graphed): one launch instead of 22-39 per forward.
        Pays when the host, not the GPU, bounds the loop — one frame per step (the reference's test() loop) or small
        video frames; feed the frames through a fixed set of device buffers (FramePipeline does), graphs are keyed by
        the input buffer.  Same kernels, same results."""
        self.use_graphs = bool(on)
        return self

    def set_ingest(self, mean, std, bgr=False):
        """enable uint8 HWC frames [N,H,W,3] as input of forward()/predict(): ToTensorVideoImage + Normalize
        (data_transforms.py:109-125, :256-281; mean/std from info.json) are fused into the stem kernel."""
        self._ingest = (mean, std, bgr)
        for eng in self._engines.values():
            eng.set_ingest(mean, std, bgr)
        return self

    def set_masks(self, mask_dict):
        self._mask_dict = mask_dict
        for eng in self._engines.values():
            eng.set_masks(mask_dict)
        return self

    def set_pruner(self, pruner):
        return self.set_masks(None if pruner is None else pruner.mask_dict)

    def set_act_dtype(self, act_dtype):
        if act_dtype != self._act_dtype:
            self._act_dtype = act_dtype
            self._close_engines()
        return self

    def prepare(self, device=None, force=False):
        """build (or, with force=True, rebuild from scratch) the device-side cache for `device` now"""
        dev = torch.device(device or next(self.parameters()).device)
        if force:
            self.invalidate()
        with torch.cuda.device(dev):
            return self._eng(dev).refresh(dev)

    def invalidate(self):
        """forget every derived cache (tile lists, packed weights, BN affines, head/stem weights).  Needed after
        parameter writes that bypass the autograd version counter (``p.data.mul_()``, ``p.data = ...``)."""
        for eng in self._engines.values():
            eng.invalidate()
        return self

    def engine(self, device=None):
        return self._eng(device)

    def _close_engines(self):
        with self._engine_lock:
            for eng in self._engines.values():
                eng.close()
            self._engines.clear()

    def _eng(self, where=None):
        """engine of the CUDA device `where` (a tensor, a device, or None = the parameters' device)"""
        if not self.use_torch_up:
            up = self.up.weight
            if tuple(up.shape[2:]) != (16, 16) or self.up.stride != (8, 8) or self.up.padding != (4, 4):
                raise ffi.Drnb200Error("`up` must be ConvTranspose2d(k=16, s=8, p=4)")
        if isinstance(where, torch.Tensor):
            where = where.device
        dev = torch.device(where) if where is not None else next(self.parameters()).device
        key = dev.index if dev.type == "cuda" and dev.index is not None else (
            torch.cuda.current_device() if dev.type == "cuda" and torch.cuda.is_available() else -1)
        eng = self._engines.get(key)
        if eng is None:
            with self._engine_lock:
                eng = self._engines.get(key)
                if eng is None:
                    ffi.lib()                      # fail loudly if the CUDA library is missing
                    if not self.use_torch_up and not torch.equal(self.up.weight.detach()[:, 0].cpu().double(),
                                                                 _bilinear_kernel(self.up.weight.shape[0], 16)):
                        raise ffi.Drnb200Error(
                            "`up.weight` differs from fill_up_weights (semantic_seg.py:115-124): the fused head "
                            "hard-codes the analytic bilinear kernel and would ignore this checkpoint's values")
                    origin = self._origin() or self
                    eng = Engine(origin, act_dtype=self._act_dtype, verify_weights=self._verify_weights)
                    eng.set_masks(self._mask_dict)
                    if self._ingest is not None:
                        eng.set_ingest(*self._ingest)
                    self._engines[key] = eng
        return eng

    def __getstate__(self):
        # copy.deepcopy / pickling: engines hold raw plan handles and must not be duplicated
        state = dict(self.__dict__)
        for k in ("_engines", "_engine_lock", "_origin"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._engines = {}
        self._engine_lock = threading.Lock()
        object.__setattr__(self, "_origin", weakref.ref(self))

    def __del__(self):
        try:
            origin = self._origin()
            if origin is None or origin is self:       # replicas do not own the engines
                for eng in self._engines.values():
                    eng.close()
        except Exception:
            pass


_UP_CACHE = {}


def _bilinear_kernel(classes, k):
    """[classes, k, k] float64 -> what fill_up_weights writes (as float32 values)"""
    key = (classes, k)
    if key not in _UP_CACHE:
        f = math.ceil(k / 2)
        c = (2 * f - 1 - f % 2) / (2.0 * f)
        ax = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=torch.float64)
        _UP_CACHE[key] = torch.outer(ax, ax).to(torch.float32).double().expand(classes, k, k)
    return _UP_CACHE[key]
