"""Host-side mirror (video-seg-model-compress_b200/drnb200) against fixtures generated from the real reference.
CPU only: module tree / state_dict keys, every pruner's masks, BSR text export, C-ABI symbol table."""
import collections
import ctypes
import io
import json
import os
import re
import contextlib

import numpy as np
import pytest
import torch

import drnb200
from drnb200 import ffi
from drnb200.pruners import (BlockPruner, BlockPrunerConfig, BlockletType, GroupingPruner,
                             GroupingPrunerConfig, HbPruner, HbPrunerConfig, RmbPruner, RmbPrunerConfig,
                             RmcdbPruner, RmcdbPrunerConfig, SRMBRepMasker, SRMBRepMaskerConfig, make_pruner)
from conftest import ROOT
from helpers import golden, load_keys
from oracle import recipe

MASKS = np.load(golden("pruner_masks.npz"))


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _w(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)).numpy()


W_A, W_B, W_C = _w((64, 32, 3, 3), 11), _w((32, 64, 1, 1), 12), _w((24, 20, 3, 3), 13)


def _same(name, mask):
    ref = recipe.unpack_mask_bits(MASKS[name], mask.shape)
    got = (np.asarray(mask) != 0).astype(np.float32)
    assert np.array_equal(got, ref), name


@pytest.mark.parametrize("arch", ["drn_d_22", "drn_d_38", "drn_d_54", "drn_c_26"])
def test_state_dict_keys_and_shapes(arch):
    ref = load_keys(arch)
    m = drnb200.DRNSeg(arch, 19, pretrained_model=None, pretrained=False)
    mine = collections.OrderedDict((k, tuple(v.shape)) for k, v in m.state_dict().items())
    assert list(mine.items()) == list(ref.items())
    base = drnb200.DRNSeg(arch, 19, pretrained=False, backbone_attr="base")
    assert [k.replace("base.", "layer.", 1) for k in base.state_dict()] == list(ref.keys())


def test_up_weight_is_the_reference_kernel():
    m = drnb200.DRNSeg("drn_d_22", 19, pretrained=False)
    assert np.array_equal(m.up.weight[0, 0].numpy(), np.load(golden("up_weight_row.npy")))
    assert m.up.weight.requires_grad is False and tuple(m.up.weight.shape) == (19, 1, 16, 16)
    assert len(list(m.optim_parameters())) == len(list(m.layer.parameters())) + 2


def test_engine_graph_matches_module_tree():
    for arch, n_convs in (("drn_d_22", 24), ("drn_d_38", 40), ("drn_d_54", 56), ("drn_c_26", 29)):
        m = drnb200.DRNSeg(arch, 19, pretrained=False)
        eng = m.engine()
        convs = [k for k, s in load_keys(arch).items() if len(s) == 4 and s[2] in (1, 3) and k.startswith("layer.")]
        assert sorted(k + ".weight" for op in eng.ops for k in op.keys) == sorted(convs)
        assert sum(len(op.keys) for op in eng.ops) == n_convs
        unfused = drnb200.engine.Engine(m, fuse_downsample=False)
        assert len(unfused.ops) == n_convs and len(eng.ops) <= n_convs
        for i, op in enumerate(eng.ops):
            assert op.input_from is None or op.input_from < i
            assert op.residual_from is None or op.residual_from < i


def test_cuda_path_fails_loudly_without_gpu_input():
    m = drnb200.DRNSeg("drn_d_22", 19, pretrained=False)
    with pytest.raises(ffi.Drnb200Error):
        m(torch.zeros(1, 3, 64, 64))          # CPU tensor: there is no CPU fallback
    # the reference's signature default pretrained=True (semantic_seg.py:127, the training call passes it too): no
    # model-zoo download here, so it warns and keeps the random initialisation instead of failing
    with pytest.warns(UserWarning, match="pretrained=True"):
        m2 = drnb200.DRNSeg("drn_d_22", 19)
    assert list(m2.state_dict().keys()) == list(m.state_dict().keys())
    # inference only: train mode with gradients enabled raises a targeted error instead of returning eval results
    m.train()
    with pytest.raises(ffi.Drnb200Error, match="inference path"):
        m(torch.zeros(1, 3, 64, 64))
    # the fused head hard-codes the analytic bilinear kernel: a checkpoint with a different `up.weight` is refused
    m.eval()
    with torch.no_grad():
        m.up.weight[3, 0, 5, 5] += 0.25
    with pytest.raises(ffi.Drnb200Error, match="fill_up_weights"):
        m.engine(torch.device("cpu"))


def test_module_copies_do_not_share_engines():
    """copy.deepcopy (and pickling) of the mirror must not duplicate raw plan handles; DataParallel-style shallow
    replicas DO share the original's per-device engines"""
    import copy
    m = drnb200.DRNSeg("drn_d_22", 19, pretrained=False)
    eng = m.engine(torch.device("cpu"))
    c = copy.deepcopy(m)
    assert c._engines == {} and c._origin() is c and c.engine(torch.device("cpu")) is not eng
    assert c.engine(torch.device("cpu")).m is c
    replica = m._replicate_for_data_parallel()
    assert replica._engines is m._engines and replica._origin() is m
    assert replica.engine(torch.device("cpu")) is eng and eng.m is m
    assert m.act_dtype_default == "fp16"


BLOCK_CASES = {
    "block_a_uncollapsed": (W_A, (0.75, 16, 8, -1, -1, False)),
    "block_a_collapsed": (W_A, (0.5, 8, 24, -1, -1, True)),
    "block_a_sub": (W_A, (0.5, 8, 4, 32, 16, False)),
    "block_b_1x1conv": (W_B, (0.75, 8, 16, -1, -1, False)),
    "block_c_ragged": (W_C, (0.6, 16, 8, -1, -1, False)),
    "block_a_unstructured": (W_A, (0.9, 1, 1, -1, -1, True)),
    "block_a_fullrow": (W_A, (0.5, -1, 4, -1, -1, False)),
}


@pytest.mark.parametrize("name", sorted(BLOCK_CASES))
def test_block_pruner_masks(name):
    w, args = BLOCK_CASES[name]
    cfg = BlockPrunerConfig(*args)
    m = BlockPruner.generate_mask_by_pruning(w, cfg)
    assert m.shape == w.shape and m.dtype == w.dtype
    _same(name, m)
    np.random.seed(77)
    _same(name + "_static", BlockPruner.generate_mask_by_construction(w, cfg))


def test_hb_grouping_masks():
    hb = HbPrunerConfig([BlockPrunerConfig(0.875, 16, 8, -1, -1, False), BlockPrunerConfig(0.875, 1, 1, -1, -1, True)])
    _same("hb_a", HbPruner.generate_mask(W_A, hb, False))
    np.random.seed(78)
    m = HbPruner.generate_mask(W_A, hb, True)
    _same("hb_a_static", m)
    assert m.max() == MASKS["hb_a_static_max"][0]
    _same("group_a", GroupingPruner.construct_mask(W_A, GroupingPrunerConfig(4)))


def test_rmb_rmcdb_masks():
    _same("rmb_a", _quiet(RmbPruner.prune_tensor_as_rmb, W_A, RmbPrunerConfig(32, 72, 0.5, [BlockletType(8, 9)], [2])))
    _same("rmb_b", _quiet(RmbPruner.prune_tensor_as_rmb, W_B,
                          RmbPrunerConfig(16, 32, 0.5, [BlockletType(4, 4), BlockletType(2, 8)], [1, 1])))
    _same("rmcdb_a", RmcdbPruner.prune_tensor_as_rmcdb(W_A, RmcdbPrunerConfig(32, 72, 0.5, [BlockletType(8, 9)], [2], True)))
    _same("rmcdb_b", RmcdbPruner.prune_tensor_as_rmcdb(
        W_B, RmcdbPrunerConfig(16, 32, 0.0, [BlockletType(4, 4), BlockletType(2, 8)], [1, 2], True)))
    np.random.seed(79)
    _same("rmcdb_a_static", RmcdbPruner.construct_rmcdb_matrix(
        W_A, RmcdbPrunerConfig(32, 72, 0.0, [BlockletType(8, 9)], [2], True)))


@pytest.mark.parametrize("pat", ["RANDOM", "UROW", "RAMANUJAN", "TRANS", "CDIA", "CDIASTRIDE", "COLUMN", "CBAND",
                                 "CCDIA", "CCOLUMN", "GROUP"])
def test_srmbrep_patterns(pat):
    for rep in (True, False):
        np.random.seed(80)
        sq = pat == "TRANS"
        cfg = SRMBRepMaskerConfig(32, 32 if sq else 16, 16, 16 if sq else 8, 1 if sq else 2, 1, 0.5, "UROW",
                                  0.75, pat, rep, False, 0.5, False)
        m = _quiet(SRMBRepMasker.construct_mask, np.zeros((64, 32, 3, 3), dtype=np.float32), cfg)
        _same("srmb_%s_%d" % (pat, int(rep)), m)


def test_srmbrep_special_cases_and_shipped_config():
    np.random.seed(81)
    cfg = SRMBRepMaskerConfig(-1, -1, 32, 32, 1, 1, 0.0, "UROW", 0.625, "TRANS", True, True, 0.5, False)
    _same("srmb_trans_dense", SRMBRepMasker.construct_mask(np.zeros((32, 32, 1, 1), dtype=np.float32), cfg))
    np.random.seed(82)
    cfg = SRMBRepMaskerConfig(-1, -1, 16, 16, 1, 1, 0.0, "UROW", 0.5, "RAMANUJAN", True, True, 0.5, True)
    _same("srmb_ramanujan_sym", SRMBRepMasker.construct_mask(np.zeros((32, 32, 1, 1), dtype=np.float32), cfg))
    with open(golden("srmb_optimal_entry10.json")) as fh:
        e = json.load(fh)
    np.random.seed(83)
    cfg = SRMBRepMaskerConfig(e["obh"], e["obw"], e["cbh"], e["cbw"], e["ibh"], e["ibw"], e["osp"], e["opat"],
                              e["isp"], e["ipat"], e["is_repetitive"], e["collapse_tensor"], e["cross_prob"],
                              e["is_symmetric"])
    m = SRMBRepMasker.construct_mask(np.zeros(tuple(e["shape"]), dtype=np.float32), cfg)
    _same("srmb_optimal_entry10", m)
    assert abs(1 - np.count_nonzero(m) / m.size - 0.75) < 1e-6


def test_pruner_objects_follow_the_reference_contract(tmp_path):
    """config-file schema, mask_dict keys/shape/dtype, apply_masks in place, print_stats, type dispatch"""
    model = drnb200.DRNSeg("drn_d_22", 19, pretrained=False)
    shapes = collections.OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())
    path = tmp_path / "block.json"
    cfg = recipe.block_pruner_config(shapes, 0.75, str(path))
    assert sum(len(c["layer_set"]) for c in cfg["configs"]) == 24
    pruner = make_pruner(str(path), on_gpu=False)
    assert isinstance(pruner, BlockPruner) and sorted(pruner.layer_configs) == sorted(recipe.prunable_keys(shapes))
    pruner.generate_masks(model, is_static=False, verbose=False)
    sd = model.state_dict()
    for k, m in pruner.mask_dict.items():
        assert m.shape == sd[k].shape and m.dtype == sd[k].dtype
    before = sd["layer.6.0.conv2.weight"].clone()
    pruner.apply_masks(model)
    after = model.state_dict()["layer.6.0.conv2.weight"]
    assert torch.equal(after, before * pruner.mask_dict["layer.6.0.conv2.weight"])
    sp = pruner.sparsity()
    assert abs(sp["layer.6.0.conv2.weight"] - 0.75) < 1e-9 and abs(sp["layer.8.0.weight"] - 0.75) < 1e-9
    # fixture produced by the reference's BlockPruner on the same seeded weights
    fx = np.load(golden("fwd_drn_d_22_64x128_block75.npz"))
    sd2 = recipe.make_state_dict(load_keys("drn_d_22"), seed=int(fx["seed"]))
    model.load_state_dict(sd2, strict=False)
    pruner2 = make_pruner(str(path), on_gpu=False)
    pruner2.generate_masks(model)
    for k, m in pruner2.mask_dict.items():
        assert np.array_equal(recipe.pack_mask_bits(m.numpy()), fx["maskbits:" + k]), k
    for ptype, cls in (("hb", HbPruner), ("grouping", GroupingPruner), ("rmb", RmbPruner), ("rmcdb", RmcdbPruner),
                       ("srmbrep", SRMBRepMasker)):
        p = tmp_path / (ptype + ".json")
        p.write_text(json.dumps({"pruner_type": ptype, "configs": []}))
        assert isinstance(make_pruner(str(p), on_gpu=False), cls)
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"pruner_type": "nope", "configs": []}))
    with pytest.raises(ValueError):
        make_pruner(str(bad))


def test_bsr_export_is_byte_identical(tmp_path):
    fx = np.load(golden("bsr_case.npz"))
    bm = BlockPruner.generate_block_matrix(fx["mat"], 4, 4)
    out = tmp_path / "bsr.txt"
    BlockPruner.write_block_matrix_to_file(bm, str(out))
    assert out.read_text() == str(fx["text"])
    back = BlockPruner.read_block_matrix_from_file(str(out))
    assert np.array_equal(BlockPruner.block_matrix_to_dense(back), fx["mat"])
    # the reference's own golden file round-trips through reader + exporter
    ref = BlockPruner.read_block_matrix_from_file(golden("block_test.txt"))
    dense = BlockPruner.block_matrix_to_dense(ref).astype(int)
    out2 = tmp_path / "bt.txt"
    BlockPruner.write_block_matrix_to_file(BlockPruner.generate_block_matrix(dense, 2, 2), str(out2))
    assert out2.read_text() == open(golden("block_test.txt")).read()


def test_shard_frames_partition():
    for n in (0, 1, 7, 8, 25, 64):
        for ws in (1, 2, 3, 8):
            spans = [drnb200.shard_frames(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_c_abi_library_exports_every_declared_symbol():
    """no compute calls: the library must load on a GPU-less host and export what include/drnb200.h declares"""
    header = open(os.path.join(ROOT, "include", "drnb200.h")).read()
    declared = set(re.findall(r"\b(drnb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(ffi.SIGNATURES), declared ^ set(ffi.SIGNATURES)
    assert os.path.exists(ffi.LIB_PATH), "build the library first: make -C video-seg-model-compress_b200/csrc"
    handle = ctypes.CDLL(ffi.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    lib = ffi.lib()
    assert lib.drnb200_version() == 110
    # the timing probes that make results invalid (DRNB200_DBG, `make EXTRA=-DDRNB200_DIAG`) must not be in the
    # library that ships: the built .so travels to the GPU box as it is
    assert b"DRNB200_DBG" not in open(ffi.LIB_PATH, "rb").read(), "libdrnb200.so is a -DDRNB200_DIAG build: make clean && make"
    # argument validation happens before any CUDA call, so it is testable here
    assert lib.drnb200_compact_mask(None, 8, 8, 3, 3, 8, 8, None, None, None, None) == -1
    assert b"null pointer" in lib.drnb200_last_error()


@pytest.mark.parametrize("act", ["fp16", "bf16"])
def test_ingest_table_host_function_is_bit_exact(act):
    """drnb200_ingest_lut is a HOST function of the C ABI (no GPU): it must equal the reference's transform of every
    byte value (fixture from the real data_transforms.py) rounded once to the activation dtype"""
    fx = np.load(golden("frameio.npz"))
    lut = drnb200.ingest_lut(fx["mean"], fx["std"], {"bf16": 0, "fp16": 1}[act])
    ref = torch.from_numpy(fx["lut"]).to(torch.float16 if act == "fp16" else torch.bfloat16).view(torch.int16)
    assert lut.shape == (3, 256) and torch.equal(lut, ref)
    assert drnb200.CITYSCAPE_PALETTE.shape == (20, 3) and np.array_equal(drnb200.CITYSCAPE_PALETTE, fx["palette"])


def test_frameio_fails_loudly_on_cpu_tensors():
    with pytest.raises(ffi.Drnb200Error):
        drnb200.colorize(torch.zeros(4, 4, dtype=torch.uint8))
    model = drnb200.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False).eval()
    with pytest.raises(ffi.Drnb200Error):
        model.predict(torch.zeros(1, 16, 16, 3, dtype=torch.uint8))


def test_checkpoint_ingestion_prefixes_and_prune_buffers(tmp_path):
    """DataParallel `module.` prefix, `base.` flavour of the video scripts, save_checkpoint dicts and
    torch.nn.utils.prune's weight_orig/weight_mask all load into the drop-in module with the masks recovered"""
    import torch.nn.utils.prune as prune
    torch.manual_seed(3)
    src = drnb200.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False)
    ref_sd = collections.OrderedDict((k, v.clone()) for k, v in src.state_dict().items())
    # (a) checkpoint dict + module. prefix + base. backbone name
    ck = {"epoch": 7, "state_dict": collections.OrderedDict(
        ("module." + k.replace("layer.", "base.", 1) if k.startswith("layer.") else "module." + k, v)
        for k, v in ref_sd.items())}
    path = tmp_path / "checkpoint.pth.tar"
    torch.save(ck, str(path))
    dst = drnb200.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False)
    masks = drnb200.load_checkpoint(dst, str(path))
    assert not masks
    for k, v in dst.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    # the `base` flavoured module takes a `layer.` checkpoint
    dst_b = drnb200.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False, backbone_attr="base")
    drnb200.load_checkpoint(dst_b, ref_sd)
    assert torch.equal(dst_b.state_dict()["base.layer3.0.conv1.weight"], ref_sd["layer.layer3.0.conv1.weight"]) \
        if "layer.layer3.0.conv1.weight" in ref_sd else True
    # (b) torch.nn.utils.prune re-parametrisation (semseg_unstructured.py:770-773)
    pr = drnb200.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False)
    pr.load_state_dict(ref_sd)
    convs = [(n, m) for n, m in pr.named_modules() if isinstance(m, torch.nn.Conv2d) and m.kernel_size == (3, 3)][:3]
    for _, m in convs:
        prune.l1_unstructured(m, name="weight", amount=0.9)
    psd = pr.state_dict()
    assert any(k.endswith("weight_orig") for k in psd)
    sd2, masks2 = drnb200.normalize_state_dict(psd)
    assert not any(k.endswith("_orig") or k.endswith("_mask") for k in sd2)
    assert len(masks2) == 3
    for n, m in convs:
        key = n + ".weight"
        assert torch.equal(sd2[key], psd[n + ".weight_orig"] * psd[n + ".weight_mask"])
        assert torch.equal(masks2[key], psd[n + ".weight_mask"])
        assert abs(float(masks2[key].mean()) - 0.1) < 0.01
    dst2 = drnb200.DRNSeg("drn_d_22", 19, pretrained_model=None, pretrained=False)
    got = drnb200.load_checkpoint(dst2, psd)
    assert set(got) == set(masks2) and dst2._mask_dict is got
    # masks implied by the zeros of an already-pruned checkpoint
    z = drnb200.masks_from_zeros(sd2, keys=set(masks2))
    assert all(torch.equal(z[k], masks2[k]) for k in masks2)
    # mismatching checkpoints fail loudly
    with pytest.raises(KeyError):
        drnb200.load_checkpoint(dst2, {"layer.nope.weight": torch.zeros(1)})


def test_multiscale_coefficient_tables_equal_pillows():
    """drnb200.multiscale.bilinear_coeffs (vectorised, product) == oracle/ms_oracle.bilinear_coeffs (scalar
    restatement of Pillow's precompute_coeffs, pinned to the real reference in test_oracle_pinned.py), bit for bit"""
    from drnb200 import multiscale
    from oracle import ms_oracle
    for a, b in [(28, 56), (98, 56), (37, 56), (131, 56), (512, 1024), (1792, 1024), (3584, 2048), (7, 50), (300, 3)]:
        got, ref = multiscale.bilinear_coeffs(a, b), ms_oracle.bilinear_coeffs(a, b)
        assert all(np.array_equal(g, r) and g.dtype == r.dtype for g, r in zip(got, ref)), (a, b)
    assert multiscale.SCALES == [0.5, 0.75, 1.25, 1.5, 1.75]          # semantic_seg.py:578


def test_multiscale_fails_loudly_on_cpu_tensors():
    from drnb200 import multiscale
    with pytest.raises(ffi.Drnb200Error):
        multiscale.resize_accumulate(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 8, 8), first=True)
    with pytest.raises(ffi.Drnb200Error):
        multiscale.argmax_labels(torch.zeros(1, 2, 8, 8))


def test_rmb_rmcdb_text_export_is_byte_identical(tmp_path):
    """dump_fpath of prune_tensor_as_rmb / prune_tensor_as_rmcdb (pruners/RmbPruner.py:247-378,
    pruners/RmcdbPruner.py:320-439) against the files the real reference wrote (gen_golden_export.py)"""
    import zlib
    fx = np.load(golden("blocklet_export.npz"))
    cases = {
        "rmb_a": (RmbPruner.prune_tensor_as_rmb, W_A, RmbPrunerConfig(32, 72, 0.5, [BlockletType(8, 9)], [2])),
        "rmb_b": (RmbPruner.prune_tensor_as_rmb, W_B,
                  RmbPrunerConfig(16, 32, 0.5, [BlockletType(4, 4), BlockletType(2, 8)], [1, 1])),
        "rmcdb_a": (RmcdbPruner.prune_tensor_as_rmcdb, W_A,
                    RmcdbPrunerConfig(32, 72, 0.5, [BlockletType(8, 9)], [2], True)),
        "rmcdb_b": (RmcdbPruner.prune_tensor_as_rmcdb, W_B,
                    RmcdbPrunerConfig(16, 32, 0.0, [BlockletType(4, 4), BlockletType(2, 8)], [1, 2], True)),
    }
    for name, (fn, w, cfg) in cases.items():
        path = tmp_path / (name + ".txt")
        mask = _quiet(fn, w, cfg, str(path))
        _same(name, mask)                                   # exporting does not change the mask
        assert path.read_bytes() == zlib.decompress(fx[name].tobytes()), name


def test_projection_launch_list_is_chosen_by_frame_width():
    """engine.ops_proj: blocks with a stride-1 1x1 shortcut as [conv1] + [conv2 + shortcut in K]; same indices and
    state-dict coverage as the default list; used only when the stage's rows are wider than 128 pixels"""
    from drnb200.engine import Engine, ProjResidualConv, FusedFirstConv
    m = drnb200.DRNSeg("drn_d_22", 19, pretrained=False)
    eng = Engine(m, act_dtype="fp16")
    assert eng.ops_proj is not None and len(eng.ops_proj) == len(eng.ops)
    proj = [o for o in eng.ops_proj if isinstance(o, ProjResidualConv)]
    assert [o.key for o in proj] == ["layer.5.0.conv2", "layer.6.0.conv2"]
    assert sorted(k for o in eng.ops_proj for k in o.keys) == sorted(k for o in eng.ops for k in o.keys)
    assert [type(o) for o in eng.ops if isinstance(o, FusedFirstConv)] == [FusedFirstConv] * 4
    assert eng.ops_for(1024, 2048) is eng.ops_proj and eng.ops_for(64, 1032) is eng.ops_proj
    assert eng.ops_for(512, 1024) is eng.ops and eng.ops_for(1024, 1024) is eng.ops
    for i, (a, b) in enumerate(zip(eng.ops, eng.ops_proj)):
        assert (a is b) or a.key == b.key
    # a bottleneck network has no such block
    assert Engine(drnb200.DRNSeg("drn_d_54", 19, pretrained=False), act_dtype="fp16").ops_proj is None


def test_multiscale_coefficient_properties_random_sizes():
    """properties of Pillow's tables for arbitrary axis sizes: rows normalised, windows inside the source and
    non-decreasing, tap count bounded by ksize — what the kernel's scratch sizing (rows_for) relies on"""
    from drnb200 import multiscale
    rng = np.random.RandomState(5)
    for _ in range(200):
        a, b = int(rng.randint(1, 700)), int(rng.randint(1, 700))
        lo, cnt, kk = multiscale.bilinear_coeffs(a, b)
        scale = a / b
        support = max(scale, 1.0)
        assert kk.shape == (b, int(np.ceil(support)) * 2 + 1)
        assert (cnt >= 1).all() and (cnt <= kk.shape[1]).all() and (lo >= 0).all() and (lo + cnt <= a).all()
        assert np.allclose(kk.sum(1), 1.0, atol=1e-12) and (kk >= 0).all()
        assert (np.diff(lo) >= 0).all() and (np.diff(lo + cnt) >= 0).all()
        # rows a strip of TY output rows touches: the bound drnb200_ms_accumulate sizes its shared-memory scratch with
        for ty in (16, 4, 1):
            need = max(int(lo[min(y0 + ty, b) - 1] + cnt[min(y0 + ty, b) - 1] - lo[y0]) for y0 in range(0, b, ty))
            assert need <= int((ty - 1) * scale + 2.0 * support) + 2, (a, b, ty, need)


def test_ctypes_signatures_match_the_header_prototypes():
    """ABI drift guard: every prototype in include/drnb200.h has as many parameters as its ctypes signature, and
    the ConvDesc mirror has exactly the fields of drnb200_conv_desc, in order"""
    header = open(os.path.join(ROOT, "include", "drnb200.h")).read()
    code = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    for name, (_, argtypes) in ffi.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^)]*)\)\s*;" % name, code)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(argtypes), (name, n, len(argtypes))
    body = re.search(r"typedef struct drnb200_conv_desc \{(.*?)\} drnb200_conv_desc;", code, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            assert decl.startswith("int32_t"), decl
            fields += [f.strip() for f in decl[len("int32_t"):].split(",")]
    assert fields == [f for f, _ in ffi.ConvDesc._fields_]
    assert int(re.search(r"#define DRNB200_KB_PROJ \((\d+) << (\d+)\)", header).group(1)) << 20 == ffi.KB_PROJ
