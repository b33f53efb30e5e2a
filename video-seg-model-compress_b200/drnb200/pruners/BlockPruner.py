"""Block masks on the matricised conv weight (reference: pruners/BlockPruner.py:40-432).

A weight [O, I, kh, kw] is viewed as the matrix (O, I*kh*kw) whose column index is ``ci*kh*kw + tap``
(:144).  Blocks are ``block_height x block_width`` rectangles of that matrix; with
``collapse_tensor == False`` the width is given in *input channels* and is multiplied by kh*kw (:157-158),
so a block covers whole channels with all their taps — the shape the B200 tile list can skip.
Magnitude pruning keeps the blocks whose sum |w| is strictly greater than the
``int(sparsity*n_blocks)-1``-th smallest score (:192-200); the static variant draws
``int((1-sparsity)*n_blocks)`` blocks with ``np.random.choice`` (:289-290).  Optionally the matrix is first
cut into ``sub_rows x sub_cols`` sub-matrices that are pruned independently (:210-227).
"""
import collections
import json

import numpy as np

from .Pruner import Pruner


def write_array_to_file(array, fh):
    fh.write("".join(str(e) + " " for e in array) + "\n")


class BlockPrunerConfig:
    def __init__(self, sparsity, block_height, block_width, sub_rows, sub_cols, collapse_tensor):
        self.sparsity = sparsity
        self.block_height = block_height
        self.block_width = block_width
        self.sub_rows = sub_rows
        self.sub_cols = sub_cols
        self.collapse_tensor = collapse_tensor

    def __str__(self):
        return "{} {} {}".format(self.block_height, self.block_width, self.sparsity)


class BlockMatrix():
    """BSR container: ``values`` (column-major inside each block), ``indices`` (block-column of each
    stored block), ``rowBlockPtr`` (first stored block of every block-row)   (:55-74)."""

    def __init__(self, rows, cols, bh, bw, values, indices, rowBlockPtr):
        self.rows, self.cols, self.bh, self.bw = rows, cols, bh, bw
        self.values, self.indices, self.rowBlockPtr = values, indices, rowBlockPtr


def _grid(n, step):
    """[(start, stop)] of consecutive spans of `step` covering range(n); the last may be short"""
    return [(s, min(s + step, n)) for s in range(0, n, step)]


def _resolve(tensor, block_height, block_width, sub_rows, sub_cols, collapse_tensor):
    """shared preamble of both mask generators (:143-163 / :255-275)"""
    mat = tensor.reshape(tensor.shape[0], tensor.size // tensor.shape[0])
    rows, cols = mat.shape
    unit = tensor.size // (tensor.shape[0] * tensor.shape[1])      # kh*kw
    bh = rows if block_height == -1 else block_height
    sr = rows if sub_rows == -1 else sub_rows
    bw = cols if block_width == -1 else (block_width if collapse_tensor else block_width * unit)
    sc = cols if sub_cols == -1 else (sub_cols if collapse_tensor else sub_cols * unit)
    return mat, rows, cols, bh, bw, sr, sc


def _block_scores(mat, bh, bw):
    """sum |w| of every block, accumulated per block in the matrix dtype like the reference (:179-187)"""
    rspans, cspans = _grid(mat.shape[0], bh), _grid(mat.shape[1], bw)
    meta = np.zeros((len(rspans), len(cspans)), dtype=mat.dtype)
    for rb, (r0, r1) in enumerate(rspans):
        for cb, (c0, c1) in enumerate(cspans):
            meta[rb, cb] = np.sum(np.abs(mat[r0:r1, c0:c1]))
    return meta, rspans, cspans


class BlockPruner(Pruner):
    def __init__(self, config_fp, on_gpu=True):
        super(BlockPruner, self).__init__(config_fp, on_gpu)

    def parse_config_file(self, config_fp):
        layer_configs = collections.OrderedDict()
        with open(config_fp) as fh:
            data = json.load(fh)
        for entry in data["configs"]:
            for layer in entry["layer_set"]:
                layer_configs[layer] = BlockPruner.generate_block_pruner_config(entry)
        return layer_configs

    def generate_masks(self, model, is_static=False, verbose=False):
        sd = model.state_dict()
        for layer, cfg in self.layer_configs.items():
            w = sd[layer].cpu().numpy()
            if verbose:
                print("Generating mask for layer {} using {} approach".format(
                    layer, "static" if is_static else "pruning"))
            mask = (BlockPruner.generate_mask_by_construction(w, cfg) if is_static
                    else BlockPruner.generate_mask_by_pruning(w, cfg))
            self._store(layer, mask)

    @staticmethod
    def generate_block_pruner_config(d):
        return BlockPrunerConfig(d["sparsity"], d["block_height"], d["block_width"], d["sub_rows"],
                                 d["sub_cols"], d["collapse_tensor"])

    # ---------------------------------------------------------------- magnitude pruning
    @staticmethod
    def generate_mask_by_pruning(tensor, pconfig, rev_mask=False):
        return BlockPruner.prune_tensor_as_block(tensor, pconfig.sparsity, pconfig.block_height,
                                                 pconfig.block_width, pconfig.sub_rows,
                                                 pconfig.sub_cols, pconfig.collapse_tensor, rev_mask)

    @staticmethod
    def prune_tensor_as_block(tensor, sparsity, block_height, block_width, sub_rows=-1, sub_cols=-1,
                              collapse_tensor=True, rev_mask=False, dump_fpath=None):
        assert 0 <= sparsity <= 1, "Sparsity should be within [0,1]"
        mat, rows, cols, bh, bw, sr, sc = _resolve(tensor, block_height, block_width, sub_rows,
                                                   sub_cols, collapse_tensor)
        mask = np.zeros((rows, cols), dtype=mat.dtype)
        if (rows, cols) == (sr, sc):
            if sparsity > 0:
                if (bh, bw) == (1, 1):
                    score = np.abs(mat)
                    cut = max(0, int(sparsity * score.size) - 1)
                    mask[score > np.sort(score.flatten())[cut]] = 1
                else:
                    meta, rspans, cspans = _block_scores(mat, bh, bw)
                    cut = max(0, int(sparsity * meta.size) - 1)
                    thresh = np.sort(np.abs(meta).flatten())[cut]
                    for rb, cb in zip(*np.nonzero(np.abs(meta) > thresh)):
                        (r0, r1), (c0, c1) = rspans[rb], cspans[cb]
                        mask[r0:r1, c0:c1] = 1
            else:
                mask.fill(1)
        else:
            # independent pruning of every sub-matrix; inside, widths are already in columns (:224)
            for r0, r1 in _grid(rows, sr):
                for c0, c1 in _grid(cols, sc):
                    mask[r0:r1, c0:c1] = BlockPruner.prune_tensor_as_block(
                        mat[r0:r1, c0:c1], sparsity, bh, bw, sr, sc, collapse_tensor=True)
        if rev_mask:
            mask = (mask + 1) % 2
        if dump_fpath is not None:
            BlockPruner.write_block_matrix_to_file(
                BlockPruner.generate_block_matrix(mat * mask, bh, bw), dump_fpath)
        return mask.reshape(tensor.shape)

    # ---------------------------------------------------------------- static (random) construction
    @staticmethod
    def generate_mask_by_construction(tensor, pconfig, rev_mask=False):
        return BlockPruner.construct_tensor_as_block(tensor, pconfig.sparsity, pconfig.block_height,
                                                     pconfig.block_width, pconfig.sub_rows,
                                                     pconfig.sub_cols, pconfig.collapse_tensor, rev_mask)

    @staticmethod
    def construct_tensor_as_block(tensor, sparsity, block_height, block_width, sub_rows=-1,
                                  sub_cols=-1, collapse_tensor=True, rev_mask=False, dump_fpath=None):
        assert 0 <= sparsity <= 1, "Sparsity should be within [0,1]"
        mat, rows, cols, bh, bw, sr, sc = _resolve(tensor, block_height, block_width, sub_rows,
                                                   sub_cols, collapse_tensor)
        mask = np.zeros((rows, cols), dtype=mat.dtype)
        if (rows, cols) == (sr, sc):
            if sparsity > 0:
                rspans, cspans = _grid(rows, bh), _grid(cols, bw)
                n_blocks = len(rspans) * len(cspans)
                keep = np.random.choice(n_blocks, int((1.0 - sparsity) * n_blocks), replace=False)
                for flat in keep:
                    (r0, r1), (c0, c1) = rspans[flat // len(cspans)], cspans[flat % len(cspans)]
                    mask[r0:r1, c0:c1] = 1
            else:
                mask.fill(1)
        else:
            for r0, r1 in _grid(rows, sr):
                for c0, c1 in _grid(cols, sc):
                    mask[r0:r1, c0:c1] = BlockPruner.construct_tensor_as_block(
                        mat[r0:r1, c0:c1], sparsity, bh, bw, sr, sc, collapse_tensor=True)
        if rev_mask:
            mask = (mask + 1) % 2
        if dump_fpath is not None:
            BlockPruner.write_block_matrix_to_file(
                BlockPruner.generate_block_matrix(mat * mask, bh, bw), dump_fpath)
        return mask.reshape(tensor.shape)

    # ---------------------------------------------------------------- BSR export (:344-432)
    @staticmethod
    def generate_block_matrix(mat, block_height, block_width):
        assert len(mat.shape) == 2
        rows, cols = mat.shape
        if block_height == 1 and block_width == 1:
            rr, cc = np.nonzero(mat)
            values = mat[rr, cc].astype(mat.dtype)
            indices = cc.astype(int)
            counts = np.zeros(rows + 1, dtype=int)
            np.add.at(counts, rr, 1)
        else:
            meta, rspans, cspans = _block_scores(mat, block_height, block_width)
            live_r, live_c = np.nonzero(meta)             # row-major order == (rb, cb) loop order
            size = block_height * block_width
            values = np.zeros(len(live_r) * size, dtype=mat.dtype)
            indices = live_c.astype(int)
            counts = np.zeros(len(rspans) + 1, dtype=int)
            for bid, (rb, cb) in enumerate(zip(live_r, live_c)):
                (r0, r1), (c0, c1) = rspans[rb], cspans[cb]
                values[bid * size:(bid + 1) * size] = mat[r0:r1, c0:c1].flatten("F")
                counts[rb] += 1
        ptr = np.zeros_like(counts)
        ptr[1:] = np.cumsum(counts[:-1])
        return BlockMatrix(rows, cols, block_height, block_width, values, indices, ptr)

    @staticmethod
    def write_block_matrix_to_file(block_mat, filepath="block_data.txt"):
        with open(filepath, "w") as fh:
            for v in (block_mat.rows, block_mat.cols, block_mat.bh, block_mat.bw,
                      block_mat.rowBlockPtr[-1]):
                fh.write(str(v) + "\n")
            write_array_to_file(block_mat.values, fh)
            write_array_to_file(block_mat.indices, fh)
            write_array_to_file(block_mat.rowBlockPtr, fh)

    @staticmethod
    def read_block_matrix_from_file(filepath):
        """inverse of write_block_matrix_to_file (the reference has no reader)"""
        with open(filepath) as fh:
            rows, cols, bh, bw, nnzb = (int(fh.readline()) for _ in range(5))
            values = np.array(fh.readline().split(), dtype=float)
            indices = np.array(fh.readline().split(), dtype=int)
            ptr = np.array(fh.readline().split(), dtype=int)
        assert len(indices) == nnzb
        return BlockMatrix(rows, cols, bh, bw, values, indices, ptr)

    @staticmethod
    def block_matrix_to_dense(bm):
        out = np.zeros((bm.rows, bm.cols), dtype=bm.values.dtype)
        size = bm.bh * bm.bw
        for rb in range(len(bm.rowBlockPtr) - 1):
            for bid in range(bm.rowBlockPtr[rb], bm.rowBlockPtr[rb + 1]):
                cb = bm.indices[bid]
                r0, c0 = rb * bm.bh, cb * bm.bw
                r1, c1 = min(r0 + bm.bh, bm.rows), min(c0 + bm.bw, bm.cols)
                out[r0:r1, c0:c1] = bm.values[bid * size:bid * size + (r1 - r0) * (c1 - c0)].reshape(
                    (r1 - r0, c1 - c0), order="F")
        return out
