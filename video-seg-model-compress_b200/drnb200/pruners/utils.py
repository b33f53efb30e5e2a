"""helpers of the reference's pruners/utils.py"""
import numpy as np


def write_array_to_file(array, fh):
    fh.write("".join(str(e) + " " for e in array) + "\n")


def get_meta_matrix(mat, block_height, block_width):
    """sum |w| per block, floor(rows/bh) x floor(cols/bw) blocks, in the matrix dtype (pruners/utils.py:9-31)"""
    assert len(mat.shape) == 2
    if block_height == 1 and block_width == 1:
        return np.copy(mat)
    nrb, ncb = mat.shape[0] // block_height, mat.shape[1] // block_width
    meta = np.zeros((nrb, ncb), dtype=mat.dtype)
    for rb in range(nrb):
        for cb in range(ncb):
            meta[rb, cb] = np.sum(np.abs(mat[rb * block_height:(rb + 1) * block_height,
                                             cb * block_width:(cb + 1) * block_width]))
    return meta
