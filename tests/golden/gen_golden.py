"""Generate the golden fixtures in this directory by running the REAL reference (imported unmodified from
/root/reference) on the seeded weights / masks / frames of oracle/recipe.py.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/gen_golden.py
The fixtures are committed; tests never import the reference.
"""
import collections
import io
import json
import os
import sys
import tempfile
import types
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-seg-model-compress_b200"))
REF = "/root/reference"

from oracle import recipe  # noqa: E402


def import_reference():
    """bootstrap documented in SURVEY 8(c): stub the two absent third-party modules, fix argv"""
    sys.path.insert(0, REF)
    for name in ("torchsummary", "pthflops"):
        m = types.ModuleType(name)
        m.summary = lambda *a, **k: None
        m.count_ops = lambda *a, **k: None
        sys.modules[name] = m
    sys.argv = ["semantic_seg.py", "test", "-d", "/tmp"]
    import semantic_seg as S
    return S


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    torch.set_num_threads(8)
    S = import_reference()
    from pruners.BlockPruner import BlockPruner, BlockPrunerConfig
    from pruners.HbPruner import HbPruner, HbPrunerConfig
    from pruners.GroupingPruner import GroupingPruner, GroupingPrunerConfig
    from pruners.RmbPruner import RmbPruner, RmbPrunerConfig, BlockletType
    from pruners.RmcdbPruner import RmcdbPruner, RmcdbPrunerConfig
    from pruners.SRMBRepMasker import SRMBRepMasker, SRMBRepMaskerConfig

    # ------------------------------------------------------------------ 1. state_dict keys
    keys = {}
    for arch in ("drn_d_22", "drn_d_38", "drn_d_54", "drn_c_26"):
        model = S.DRNSeg(arch, 19, pretrained_model=None, pretrained=False).eval()
        keys[arch] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as fh:
        json.dump(keys, fh)

    # ------------------------------------------------------------------ 2. forward fixtures
    def run_forward(arch, h, w, pruned, seed):
        model = S.DRNSeg(arch, 19, pretrained_model=None, pretrained=False).eval()
        shapes = collections.OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())
        sd = recipe.make_state_dict(shapes, seed=seed)
        masks_bits = {}
        if pruned:
            with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
                json.dump(recipe.block_pruner_config(shapes, 0.75), fh)
            model.load_state_dict(sd, strict=False)
            pruner = quiet(BlockPruner, fh.name, on_gpu=False)
            quiet(pruner.generate_masks, model, is_static=False, verbose=False)
            os.unlink(fh.name)
            sd = recipe.sparse_reinit(sd, pruner.mask_dict, seed=seed)
            masks_bits = {k: recipe.pack_mask_bits(m.numpy()) for k, m in pruner.mask_dict.items()}
        model.load_state_dict(sd, strict=False)
        x = recipe.make_frames(1, h, w, seed=1234 + seed)
        tap_stats = {}
        hooks = []
        for name, mod in model.named_modules():
            if isinstance(mod, torch.nn.Conv2d) and name != "seg":
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, name=name: tap_stats.__setitem__(
                        name, [float(o.double().sum()), float(o.double().abs().sum())])))
        with torch.no_grad():
            final, seg = model(x)
            pred = torch.max(final, 1)[1]
        for hk in hooks:
            hk.remove()
        out = {"seg": seg.numpy(), "labels": pred.numpy().astype(np.uint8),
               "logprob_sample": final[0, :, ::7, ::13].numpy().copy(),
               "logprob_sum": np.float64(final.double().sum().item()),
               "tap_names": np.array(list(tap_stats.keys())),
               "tap_stats": np.array(list(tap_stats.values()), dtype=np.float64),
               "seed": seed, "hw": np.array([h, w])}
        for k, b in masks_bits.items():
            out["maskbits:" + k] = b
        tag = "%s_%dx%d_%s" % (arch, h, w, "block75" if pruned else "dense")
        np.savez_compressed(os.path.join(HERE, "fwd_%s.npz" % tag), **out)
        print("forward fixture", tag, "classes predicted:", len(np.unique(out["labels"])))

    run_forward("drn_d_22", 64, 128, False, 0)
    run_forward("drn_d_22", 64, 128, True, 1)
    run_forward("drn_d_38", 32, 64, True, 2)
    run_forward("drn_d_54", 32, 64, False, 3)
    run_forward("drn_c_26", 32, 64, False, 4)

    # ------------------------------------------------------------------ 3. pruner masks on seeded weights
    masks = {}

    def seeded_weight(shape, seed):
        g = torch.Generator().manual_seed(seed)
        return torch.randn(shape, generator=g).numpy()

    w_a = seeded_weight((64, 32, 3, 3), 11)
    w_b = seeded_weight((32, 64, 1, 1), 12)
    w_c = seeded_weight((24, 20, 3, 3), 13)
    # BlockPruner — pruning path, collapsed / uncollapsed / sub-matrices / ragged edge / unstructured
    cases = {
        "block_a_uncollapsed": (w_a, BlockPrunerConfig(0.75, 16, 8, -1, -1, False)),
        "block_a_collapsed": (w_a, BlockPrunerConfig(0.5, 8, 24, -1, -1, True)),
        "block_a_sub": (w_a, BlockPrunerConfig(0.5, 8, 4, 32, 16, False)),
        "block_b_1x1conv": (w_b, BlockPrunerConfig(0.75, 8, 16, -1, -1, False)),
        "block_c_ragged": (w_c, BlockPrunerConfig(0.6, 16, 8, -1, -1, False)),
        "block_a_unstructured": (w_a, BlockPrunerConfig(0.9, 1, 1, -1, -1, True)),
        "block_a_fullrow": (w_a, BlockPrunerConfig(0.5, -1, 4, -1, -1, False)),
    }
    for name, (w, cfg) in cases.items():
        masks[name] = recipe.pack_mask_bits(BlockPruner.generate_mask_by_pruning(w, cfg))
        np.random.seed(77)
        masks[name + "_static"] = recipe.pack_mask_bits(BlockPruner.generate_mask_by_construction(w, cfg))
    # HbPruner
    hb = HbPrunerConfig([BlockPrunerConfig(0.875, 16, 8, -1, -1, False), BlockPrunerConfig(0.875, 1, 1, -1, -1, True)])
    masks["hb_a"] = recipe.pack_mask_bits(HbPruner.generate_mask(w_a, hb, False))
    np.random.seed(78)
    hbm = HbPruner.generate_mask(w_a, hb, True)
    masks["hb_a_static"] = recipe.pack_mask_bits(hbm)
    masks["hb_a_static_max"] = np.array([hbm.max()])
    # GroupingPruner
    masks["group_a"] = recipe.pack_mask_bits(GroupingPruner.construct_mask(w_a, GroupingPrunerConfig(4)))
    # RmbPruner / RmcdbPruner (pruning path)
    rmb = RmbPrunerConfig(32, 72, 0.5, [BlockletType(8, 9)], [2])
    masks["rmb_a"] = recipe.pack_mask_bits(quiet(RmbPruner.prune_tensor_as_rmb, w_a, rmb))
    rmb2 = RmbPrunerConfig(16, 32, 0.5, [BlockletType(4, 4), BlockletType(2, 8)], [1, 1])
    masks["rmb_b"] = recipe.pack_mask_bits(quiet(RmbPruner.prune_tensor_as_rmb, w_b, rmb2))
    rmc = RmcdbPrunerConfig(32, 72, 0.5, [BlockletType(8, 9)], [2], True)
    masks["rmcdb_a"] = recipe.pack_mask_bits(quiet(RmcdbPruner.prune_tensor_as_rmcdb, w_a, rmc))
    rmc2 = RmcdbPrunerConfig(16, 32, 0.0, [BlockletType(4, 4), BlockletType(2, 8)], [1, 2], True)
    masks["rmcdb_b"] = recipe.pack_mask_bits(quiet(RmcdbPruner.prune_tensor_as_rmcdb, w_b, rmc2))
    np.random.seed(79)
    rmc3 = RmcdbPrunerConfig(32, 72, 0.0, [BlockletType(8, 9)], [2], True)
    masks["rmcdb_a_static"] = recipe.pack_mask_bits(quiet(RmcdbPruner.construct_rmcdb_matrix, w_a, rmc3))
    # SRMBRepMasker: every pattern the generator knows, repetitive and not
    for pat in ("RANDOM", "UROW", "RAMANUJAN", "TRANS", "CDIA", "CDIASTRIDE", "COLUMN", "CBAND", "CCDIA",
                "CCOLUMN", "GROUP"):
        for rep in (True, False):
            np.random.seed(80)
            square = pat == "TRANS"
            cfg = SRMBRepMaskerConfig(32, 32 if square else 16, 16, 16 if square else 8, 1 if square else 2, 1, 0.5, "UROW",
                                      0.75, pat, rep, False, 0.5, False)
            m = SRMBRepMasker.construct_mask(np.zeros((64, 32, 3, 3), dtype=np.float32), cfg)
            masks["srmb_%s_%d" % (pat, int(rep))] = recipe.pack_mask_bits(m)
    np.random.seed(81)
    cfg = SRMBRepMaskerConfig(-1, -1, 32, 32, 1, 1, 0.0, "UROW", 0.625, "TRANS", True, True, 0.5, False)
    masks["srmb_trans_dense"] = recipe.pack_mask_bits(
        SRMBRepMasker.construct_mask(np.zeros((32, 32, 1, 1), dtype=np.float32), cfg))
    np.random.seed(82)
    cfg = SRMBRepMaskerConfig(-1, -1, 16, 16, 1, 1, 0.0, "UROW", 0.5, "RAMANUJAN", True, True, 0.5, True)
    masks["srmb_ramanujan_sym"] = recipe.pack_mask_bits(
        SRMBRepMasker.construct_mask(np.zeros((32, 32, 1, 1), dtype=np.float32), cfg))
    # shipped optimal_configs entry run through the real masker (config 4 of BASELINE.json)
    with open(os.path.join(REF, "optimal_configs/drn_d_22/drn_d_22_1024X768_0.00_75.00.json")) as fh:
        oc = json.load(fh)
    entry = oc["configs"][10] if isinstance(oc, dict) and "configs" in oc else None
    if entry is not None:
        np.random.seed(83)
        c = SRMBRepMaskerConfig(entry["obh"], entry["obw"], entry["cbh"], entry["cbw"], entry["ibh"], entry["ibw"],
                                entry["osp"], entry["opat"], entry["isp"], entry["ipat"], entry["is_repetitive"],
                                entry["collapse_tensor"], entry.get("cross_prob", 0.5), entry.get("is_symmetric", False))
        shape = (256, 256, 3, 3)
        masks["srmb_optimal_entry10"] = recipe.pack_mask_bits(
            SRMBRepMasker.construct_mask(np.zeros(shape, dtype=np.float32), c))
        with open(os.path.join(HERE, "srmb_optimal_entry10.json"), "w") as fh:
            json.dump({k: entry[k] for k in ("obh", "obw", "cbh", "cbw", "ibh", "ibw", "osp", "opat", "isp",
                                             "ipat", "is_repetitive", "collapse_tensor")} |
                      {"cross_prob": entry.get("cross_prob", 0.5), "is_symmetric": entry.get("is_symmetric", False),
                       "shape": list(shape)}, fh)
    np.savez_compressed(os.path.join(HERE, "pruner_masks.npz"), **masks)
    print("pruner masks:", len(masks))

    # ------------------------------------------------------------------ 4. BSR exporter
    np.random.seed(5)
    mat = (np.random.randint(1, 100, size=(12, 16)) * (np.random.rand(12, 16) > 0.3)).astype(np.float32)
    mask = BlockPruner.prune_tensor_as_block(mat.reshape(12, 16, 1, 1), 0.5, 4, 4, -1, -1, True).reshape(12, 16)
    bm = BlockPruner.generate_block_matrix(mat * mask, 4, 4)
    with tempfile.NamedTemporaryFile("r", suffix=".txt") as fh:
        BlockPruner.write_block_matrix_to_file(bm, fh.name)
        text = open(fh.name).read()
    np.savez_compressed(os.path.join(HERE, "bsr_case.npz"), mat=mat * mask, text=np.array(text))
    # the one golden file the reference ships (pruners/block_test.txt), kept verbatim as a data fixture
    with open(os.path.join(REF, "pruners/block_test.txt")) as fh:
        open(os.path.join(HERE, "block_test.txt"), "w").write(fh.read())

    # ------------------------------------------------------------------ 5. metrics
    rng = np.random.RandomState(9)
    pred = rng.randint(0, 19, size=4096)
    label = rng.randint(0, 21, size=4096)
    label[label >= 19] = 255
    hist = S.fast_hist(pred.flatten(), label.flatten(), 19)
    ious = S.per_class_iu(hist) * 100
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), pred=pred, label=label, hist=hist, ious=ious,
                        miou=np.float64(round(np.nanmean(ious), 2)),
                        tiny=S.fast_hist(np.array([0, 1, 1, 2]), np.array([0, 1, 2, 255]), 3))
    # upsample kernel
    model = S.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False)
    np.save(os.path.join(HERE, "up_weight_row.npy"), model.up.weight[0, 0].numpy())
    print("done")


if __name__ == "__main__":
    main()
