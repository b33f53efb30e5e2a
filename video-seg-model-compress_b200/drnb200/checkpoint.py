"""Checkpoint and mask ingestion for the drop-in path: the on-disk formats the reference's callers
produce and consume around `DRNSeg`.

* `torch.save({'epoch':…, 'state_dict': model.state_dict(), …})` checkpoints (semantic_seg.py:286-290, loaded at
  :598-607) and bare state dicts (`--pretrained`, :570-571);
* key prefixes: `module.` from `nn.DataParallel` (optimal_configs/* layer names), `base.` instead of `layer.` in the
  video scripts (seg_video.py:79; the commented renaming loop in seg_video_old.py:286-292);
* `torch.nn.utils.prune` re-parametrisation: `<name>.weight_orig` + `<name>.weight_mask`
  (semseg_unstructured.py:770-773) -> a plain `<name>.weight` = orig * mask plus a mask dict for the tile lists.
Everything here is host-side dictionary work; the masks end up in `DRNSeg.set_masks` and from there in
`drnb200_compact_mask`.
"""
import collections

import torch

_BACKBONES = ("layer", "base")


def unwrap(obj):
    """checkpoint object -> flat {key: tensor} (accepts the dict saved by save_checkpoint or a bare state dict)"""
    if isinstance(obj, (str, bytes)) or hasattr(obj, "__fspath__"):
        obj = torch.load(obj, map_location="cpu")
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        obj = obj["state_dict"]
    if not isinstance(obj, dict):
        raise TypeError("expected a state dict or a checkpoint with a 'state_dict' entry, got %s" % type(obj))
    return obj


def normalize_state_dict(obj, backbone_attr="layer"):
    """-> (state_dict, masks): keys renamed to this model's (`module.` stripped, `layer.`/`base.` mapped to
    `backbone_attr`), pruning re-parametrisations folded (`weight = weight_orig * weight_mask`), and
    `masks[<name>.weight]` = the `weight_mask` buffers found (empty when the checkpoint has none)."""
    if backbone_attr not in _BACKBONES:
        raise ValueError("backbone_attr must be 'layer' or 'base'")
    flat = unwrap(obj)
    renamed = collections.OrderedDict()
    for key, val in flat.items():
        k = key
        while k.startswith("module."):
            k = k[len("module."):]
        for other in _BACKBONES:
            if k.startswith(other + "."):
                k = backbone_attr + k[len(other):]
                break
        if k in renamed:
            raise KeyError("two checkpoint entries map to '%s'" % k)
        renamed[k] = val
    out = collections.OrderedDict()
    masks = collections.OrderedDict()
    for k, val in renamed.items():
        if k.endswith("_orig"):
            name = k[:-len("_orig")]
            mask = renamed.get(name + "_mask")
            if mask is None:
                raise KeyError("'%s' has no matching '%s_mask'" % (k, name))
            out[name] = val.detach() * mask.to(val.dtype)
            masks[name] = mask.detach().to(torch.float32)
        elif k.endswith("_mask") and (k[:-len("_mask")] + "_orig") in renamed:
            continue
        else:
            out[k] = val
    return out, masks


def masks_from_zeros(state_dict, keys=None):
    """liveness masks (weight != 0) for conv weights: what a checkpoint of an already-pruned network implies when no
    pruner object and no weight_mask buffers exist (apply_masks multiplies zeros in, pruners/Pruner.py:17-20)"""
    masks = collections.OrderedDict()
    for k, v in state_dict.items():
        if keys is not None and k not in keys:
            continue
        if torch.is_tensor(v) and v.dim() == 4 and k.endswith(".weight") and not k.startswith("up."):
            masks[k] = (v != 0).to(torch.float32)
    return masks


def load_checkpoint(model, obj, strict=True, use_masks=True):
    """load a reference checkpoint into a drnb200.DRNSeg; returns the mask dict handed to the engine (may be empty:
    the engine then derives liveness from the zeros of the weights themselves)."""
    sd, masks = normalize_state_dict(obj, getattr(model, "_backbone_attr", "layer"))
    own = model.state_dict()
    missing = [k for k in own if k not in sd and k != "up.weight" and not k.endswith("num_batches_tracked")]
    unexpected = [k for k in sd if k not in own]
    if strict and (missing or unexpected):
        raise KeyError("checkpoint does not match the model: missing %s, unexpected %s" % (missing[:5], unexpected[:5]))
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    if use_masks and masks:
        model.set_masks(masks)
    return masks
