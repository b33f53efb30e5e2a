"""The oracle (oracle/) against fixtures produced by the real reference (tests/golden/gen_golden.py) and
against the reference's own golden file pruners/block_test.txt.  CPU only."""
import numpy as np
import pytest
import torch

import collections

from helpers import fixture_frames, fixture_state_dict, golden, load_keys
from oracle import compact_oracle, drn_oracle, recipe

FWD = [("fwd_drn_d_22_64x128_dense.npz", "drn_d_22"), ("fwd_drn_d_22_64x128_block75.npz", "drn_d_22"),
       ("fwd_drn_d_38_32x64_block75.npz", "drn_d_38"), ("fwd_drn_d_54_32x64_dense.npz", "drn_d_54"),
       ("fwd_drn_c_26_32x64_dense.npz", "drn_c_26")]


@pytest.mark.parametrize("name,arch", FWD)
def test_forward_matches_reference(name, arch):
    fx = np.load(golden(name))
    sd, _ = fixture_state_dict(arch, fx)
    raw = {}
    logprob, seg = drn_oracle.drnseg_forward(sd, fixture_frames(fx), raw=raw)
    ref_seg = fx["seg"]
    assert seg.shape == ref_seg.shape
    scale = np.abs(ref_seg).max()
    assert np.abs(seg.numpy() - ref_seg).max() <= 2e-5 * scale + 1e-5
    pred = torch.max(logprob, 1)[1].numpy().astype(np.uint8)
    assert (pred == fx["labels"]).mean() >= 0.9999
    sample = logprob[0, :, ::7, ::13].numpy()
    assert np.abs(sample - fx["logprob_sample"]).max() <= 2e-5 * scale + 1e-4
    # every conv of the reference (forward hooks) has the same output statistics
    for cname, (s, a) in zip(fx["tap_names"], fx["tap_stats"]):
        out = raw[str(cname)].double()
        assert abs(float(out.abs().sum()) - a) <= 1e-5 * a + 1e-3, cname
        assert abs(float(out.sum()) - s) <= 1e-5 * a + 1e-3, cname
    assert len(raw) == len(fx["tap_names"])


def test_up_weights_match_fill_up_weights():
    row = np.load(golden("up_weight_row.npy"))
    ax = drn_oracle.up_weights(16).numpy()
    assert np.array_equal(np.outer(ax, ax).astype(np.float32), row)
    assert np.array_equal(ax, 1 - np.abs(2 * np.arange(16) - 15) / 16)   # analytic form used by the head


def test_metrics_match_reference():
    fx = np.load(golden("metrics.npz"))
    hist = drn_oracle.fast_hist(fx["pred"], fx["label"], 19)
    assert np.array_equal(hist, fx["hist"])
    assert np.allclose(drn_oracle.per_class_iu(hist) * 100, fx["ious"], equal_nan=True)
    assert drn_oracle.miou(hist) == float(fx["miou"])
    assert np.array_equal(drn_oracle.fast_hist(np.array([0, 1, 1, 2]), np.array([0, 1, 2, 255]), 3), fx["tiny"])


def _parse_bsr(text):
    lines = text.strip("\n").split("\n")
    rows, cols, bh, bw, nnzb = (int(v) for v in lines[:5])
    vals = np.array(lines[5].split(), dtype=float)
    idx = np.array(lines[6].split(), dtype=int)
    ptr = np.array(lines[7].split(), dtype=int)
    return rows, cols, bh, bw, nnzb, vals, idx, ptr


def test_bsr_golden_file_of_the_reference():
    """pruners/block_test.txt: rebuild the dense matrix from it, re-export, compare byte for byte"""
    text = open(golden("block_test.txt")).read()
    rows, cols, bh, bw, nnzb, vals, idx, ptr = _parse_bsr(text)
    assert (rows, cols, bh, bw, nnzb) == (8, 8, 2, 2, 8)
    assert list(idx) == [1, 3, 0, 3, 0, 2, 0, 3] and list(ptr) == [0, 2, 4, 6, 8]
    dense = np.zeros((rows, cols), dtype=int)
    for rb in range(rows // bh):
        for b in range(ptr[rb], ptr[rb + 1]):
            blk = vals[b * bh * bw:(b + 1) * bh * bw].astype(int).reshape((bh, bw), order="F")
            dense[rb * bh:(rb + 1) * bh, idx[b] * bw:(idx[b] + 1) * bw] = blk
    v2, i2, p2 = compact_oracle.bsr_from_dense(dense, bh, bw)
    assert compact_oracle.bsr_text(rows, cols, bh, bw, v2, i2, p2) == text
    # the same liveness through the tile-list restatement (taps = 1, tile == block)
    rp, kb = compact_oracle.compact_mask(dense.reshape(8, 8, 1, 1), 2, 2)
    assert list(rp) == list(ptr) and list(kb) == list(idx)


def test_bsr_seeded_case():
    fx = np.load(golden("bsr_case.npz"))
    mat = fx["mat"]
    v, i, p = compact_oracle.bsr_from_dense(mat, 4, 4)
    assert compact_oracle.bsr_text(12, 16, 4, 4, v, i, p) == str(fx["text"])


def test_compaction_order_and_expand():
    rng = np.random.RandomState(3)
    mask = np.zeros((32, 64, 3, 3), dtype=np.float32)
    for ot in range(2):
        for cib in range(4):
            if rng.rand() < 0.5:
                mask[ot * 16:(ot + 1) * 16, cib * 16:(cib + 1) * 16] = rng.rand() + 1.5   # values > 1 (Hb)
    mask[3, 5, 1, 2] = 1.0                                                             # a lone element
    rp, kb = compact_oracle.compact_mask(mask, 16, 16)
    exp = compact_oracle.expand_tile_list(rp, kb, 32, 64, 3, 3, 16, 16)
    assert np.all(exp[mask != 0] == 1)
    for ot in range(2):
        seg = kb[rp[ot]:rp[ot + 1]]
        assert np.all(np.diff(seg) > 0)
    assert (0 * 9 + 1 * 3 + 2) in kb[rp[0]:rp[1]]


def test_swizzle_is_a_permutation():
    for pitch in (32, 64, 128):
        offs = {compact_oracle.swizzle_offset(r, c, pitch) for r in range(64) for c in range(pitch // 16)}
        assert offs == set(range(0, 64 * pitch, 16))


# ------------------------------------------------------------------ rows either side of the path
def test_frame_ingest_matches_reference_transforms():
    """oracle ingest == ToTensorVideoImage + Normalize of the reference (data_transforms.py), bit for bit"""
    from oracle import frameio_oracle
    fx = np.load(golden("frameio.npz"))
    x = frameio_oracle.ingest(fx["frame"][None], fx["mean"], fx["std"])[0].numpy()
    assert x.dtype == np.float32 and np.array_equal(x, fx["x"])
    assert np.array_equal(frameio_oracle.ingest_table(fx["mean"], fx["std"]).numpy(), fx["lut"])


def test_palette_matches_reference():
    from oracle import frameio_oracle
    fx = np.load(golden("frameio.npz"))
    assert np.array_equal(frameio_oracle.CITYSCAPE_PALETTE, fx["palette"])
    assert np.array_equal(frameio_oracle.colorize(fx["pred"]), fx["color"])
    # out-of-palette labels (ignore = 255) take the last row, the rule seg_video.py spells out in a comment
    assert np.array_equal(frameio_oracle.colorize(np.array([255, 19, 18])), fx["palette"][[19, 19, 18]])


def test_multiscale_resize_sum_argmax_match_reference():
    """oracle/ms_oracle.py (Pillow BILINEAR restated) == semantic_seg.resize_4d_tensor / test_ms, bit for bit"""
    from helpers import MS_SOURCES, ms_sources
    from oracle import ms_oracle
    fx = np.load(golden("multiscale.npz"))
    H, W = (int(v) for v in fx["target"])
    assert [tuple(s) for s in fx["sources"].tolist()] == MS_SOURCES
    srcs = [t.numpy() for t in ms_sources()]
    for i, src in enumerate(srcs):
        got = ms_oracle.resize_4d_tensor(src, W, H)
        assert got.dtype == np.float32 and np.array_equal(got, fx["dst%d" % i]), MS_SOURCES[i]
    final, pred = ms_oracle.ms_combine(srcs, W, H)
    assert np.array_equal(final, fx["final"]) and np.array_equal(pred, fx["pred"])
    # the coefficient rows are normalised and the upscaling case is the 2-tap triangle
    xmin, cnt, kk = ms_oracle.bilinear_coeffs(28, 56)
    assert kk.shape[1] == 3 and np.allclose(kk.sum(1), 1.0) and cnt.max() <= 3
    xmin, cnt, kk = ms_oracle.bilinear_coeffs(98, 56)
    assert kk.shape[1] == 5 and cnt.max() <= 5


def test_oracle_use_torch_up_matches_the_real_reference():
    """use_torch_up=True (nn.UpsamplingBilinear2d, semantic_seg.py:144-145): oracle vs the real reference's output"""
    fx = np.load(golden("fwd_drn_d_22_40x72_torch_up.npz"))
    shapes = collections.OrderedDict((k, v) for k, v in load_keys("drn_d_22").items() if k != "up.weight")
    sd = recipe.make_state_dict(shapes, seed=int(fx["seed"]))
    x = recipe.make_frames(1, int(fx["hw"][0]), int(fx["hw"][1]), seed=1234 + int(fx["seed"]))
    lp, seg = drn_oracle.drnseg_forward(sd, x, use_torch_up=True)
    assert np.abs(seg.numpy() - fx["seg"]).max() <= 1e-4 * np.abs(fx["seg"]).max()
    assert np.abs(lp.numpy() - fx["logprob"]).max() <= 1e-4 * np.abs(fx["seg"]).max()
    assert (lp.argmax(1).numpy() == fx["labels"]).mean() >= 0.9999


def test_oracle_frame_resize_matches_pil():
    """T.Resize on a PIL frame (seg_video_old.py:125-128) = Pillow's 8-bit BILINEAR resampler: the restatement in
    oracle/frameio_oracle.py against files produced by real torchvision / PIL (tests/golden/gen_golden_resize.py)"""
    from oracle import frameio_oracle
    fx = np.load(golden("frame_resize.npz"))
    for i, (h, w) in enumerate(fx["sizes"]):
        got = frameio_oracle.resize_u8(fx["src"][None], int(h), int(w))[0]
        assert np.array_equal(got, fx["dst%d" % i]), (h, w)
    # the int32 tables the device path uploads are the oracle's
    from drnb200 import frameio as product
    import torch
    lo, cnt, ki, ksize = product._resize_axis_tables(284, 100, torch.device("cpu"))
    olo, ocnt, oki = frameio_oracle.coeffs_8bpc(284, 100)
    assert np.array_equal(lo.numpy(), olo) and np.array_equal(cnt.numpy(), ocnt) and np.array_equal(ki.numpy(), oki)
    assert ksize == oki.shape[1]
