"""Text exporters of the blocklet formats: the files `prune_tensor_as_rmb(..., dump_fpath=...)` and
`prune_tensor_as_rmcdb(..., dump_fpath=...)` write for the reference's external sparse kernels
(pruners/RmbPruner.py:247-378, pruners/RmcdbPruner.py:320-439).

Layout of both files: a header of scalars, then one line per array (`write_array_to_file`: elements separated and
terminated by a blank).  Blocklets are grouped by their outer block (row-major block id = grb*ncb + gcb, the order
`np.argsort` gives the ids), `indices`/`rowBlockPtr` are the BSR structure over outer blocks, `valPtr`/`indPtr`/
`bletPtr` are exclusive prefix sums per outer block, values are stored column-major per blocklet.
"""
import numpy as np

from .utils import write_array_to_file


class Blocklet(object):
    """one exported blocklet: lives in outer block (grb, gcb) of size rows x cols; `values` is rows x bw;
    `tail` is the per-blocklet-row column index array (RMB) or the diagonal offset (RMCDB)"""
    __slots__ = ("rows", "cols", "grb", "gcb", "bh", "bw", "values", "tail")

    def __init__(self, rows, cols, grb, gcb, bh, bw, values, tail):
        self.rows, self.cols, self.grb, self.gcb = rows, cols, grb, gcb
        self.bh, self.bw, self.values, self.tail = bh, bw, values, tail


def _group(blocklets, n_block_rows, n_block_cols):
    """-> (blocklets in block order, per-block [start, end) pointer, block-column index per live block, rowBlockPtr)"""
    ids = [b.grb * n_block_cols + b.gcb for b in blocklets]
    order = np.argsort(ids)                    # same call as the reference: ties keep numpy's default order
    ordered = [blocklets[i] for i in order]
    live, counts = np.unique(np.sort(ids), return_counts=True)
    ptr = np.zeros(live.size + 1, dtype=int)
    ptr[1:] = np.cumsum(counts)
    row_block_ptr = np.zeros(n_block_rows + 1, dtype=int)
    np.add.at(row_block_ptr, live // n_block_cols + 1, 1)
    return ordered, ptr, (live % n_block_cols).astype(int), np.cumsum(row_block_ptr)


def _exclusive(counts):
    out = np.zeros(len(counts) + 1, dtype=int)
    out[1:] = np.cumsum(counts)
    # the reference stores the counts in an array one longer than needed and shifts them: the last entry is the total
    return out


def _emit(path, header, arrays):
    with open(path, "w") as fh:
        for v in header:
            fh.write(str(v) + "\n")
        for a in arrays:
            write_array_to_file(a, fh)


def write_rmb(path, rows, cols, bh, bw, blocklets):
    """pruners/RmbPruner.py:247-378 (generate_rmb_mat_from_bl_mats + write_rmb_matrix_to_file)"""
    ordered, ptr, indices, row_block_ptr = _group(blocklets, rows // bh, cols // bw)
    n_blocks = len(ptr) - 1
    per_block = [ordered[ptr[i]:ptr[i + 1]] for i in range(n_blocks)]
    val_ptr = _exclusive([sum(b.values.size for b in blk) for blk in per_block])
    ind_ptr = _exclusive([sum(b.tail.size for b in blk) for blk in per_block])
    blet_ptr = _exclusive([len(blk) for blk in per_block])
    values = np.zeros(val_ptr[-1])
    l_indices = np.zeros(ind_ptr[-1], dtype=int)
    row_pat = np.zeros(len(ordered), dtype=int)
    col_pat = np.zeros(len(ordered), dtype=int)
    v = t = 0
    for k, b in enumerate(ordered):
        values[v:v + b.values.size] = b.values.flatten("F")
        l_indices[t:t + b.tail.size] = b.tail.flatten("F")
        v += b.values.size
        t += b.tail.size
        row_pat[k] = int(round(np.log2(b.rows // b.bh)))
        col_pat[k] = int(round(np.log2(b.cols // b.bw)))
    _emit(path, [rows, cols, bh, bw, val_ptr[-1], n_blocks, len(ordered), ind_ptr[-1]],
          [values, indices, row_block_ptr, row_pat, col_pat, l_indices, val_ptr, ind_ptr, blet_ptr])


def write_rmcdb(path, rows, cols, bh, bw, blocklets):
    """pruners/RmcdbPruner.py:320-439 (generate_rmcdb_mat_from_cdbl_mats + write_rmcdb_matrix_to_file)"""
    ordered, ptr, indices, row_block_ptr = _group(blocklets, rows // bh, cols // bw)
    n_blocks = len(ptr) - 1
    per_block = [ordered[ptr[i]:ptr[i + 1]] for i in range(n_blocks)]
    val_ptr = _exclusive([sum(b.values.size for b in blk) for blk in per_block])
    blet_ptr = _exclusive([len(blk) for blk in per_block])
    values = np.zeros(val_ptr[-1])
    v = 0
    for b in ordered:
        values[v:v + b.values.size] = b.values.flatten("F")
        v += b.values.size
    row_pat = np.asarray([b.bh for b in ordered], dtype=int).reshape(-1)
    col_pat = np.asarray([b.bw for b in ordered], dtype=int).reshape(-1)
    offsets = np.asarray([int(b.tail) for b in ordered], dtype=int).reshape(-1)
    _emit(path, [rows, cols, bh, bw, val_ptr[-1], n_blocks, len(ordered)],
          [values, indices, row_block_ptr, row_pat, col_pat, offsets, val_ptr, blet_ptr])
