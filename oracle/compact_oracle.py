"""ORACLE — test infrastructure, not product code.

numpy restatement of subsystem (a): mask -> block-sparse tile list -> packed live weight blocks.

Definitions it follows in the reference:
  * matricisation ``mat = tensor.reshape(O, I*kh*kw)``, column = ci*kh*kw + tap   (pruners/BlockPruner.py:144)
  * "block is live <=> it holds a non-zero"                                   (tools/visualize_layers.py:8-13,
    and ``meta_matrix[rb,cb] != 0`` of BlockPruner.generate_block_matrix, pruners/BlockPruner.py:381-396)
  * BSR arrays ``indices`` (block-column ids, row-major over live blocks) and ``rowBlockPtr``
    (exclusive prefix sum of live blocks per block-row)                        (pruners/BlockPruner.py:344-413)
The CUDA path re-orders the columns of each block-row as K-blocks ``kb = cib*taps + tap`` (all taps of a
channel block adjacent) because the implicit GEMM walks K as (channel block, tap); that permutation is
part of this restatement and is what tests compare bit-for-bit.
"""
import numpy as np


def f32_to_bf16_bits(a):
    """round-to-nearest-even float32 -> bfloat16 bit pattern (uint16)"""
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = (bits + 0x7FFF + ((bits >> 16) & 1)) >> 16
    return rounded.astype(np.uint16)


def f32_to_f16_bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).astype(np.float16).view(np.uint16)


def to_bits(a, act_dtype):
    return f32_to_bf16_bits(a) if act_dtype in (0, "bf16") else f32_to_f16_bits(a)


def block_liveness(mask, tile_o, tile_ci):
    """live[ot, kb] for kb = cib*taps + tap; mask is [O, I, kh, kw] (any dtype; != 0 means kept)"""
    O, I, kh, kw = mask.shape
    taps = kh * kw
    assert O % tile_o == 0 and I % tile_ci == 0
    nz = (np.asarray(mask) != 0).reshape(O // tile_o, tile_o, I // tile_ci, tile_ci, taps)
    live = nz.any(axis=(1, 3))                      # [n_ot, n_cib, taps]
    return live.reshape(O // tile_o, (I // tile_ci) * taps)


def compact_mask(mask, tile_o, tile_ci):
    """-> (row_ptr int32 [n_ot+1], kblk int32 [n_live]) — the rowBlockPtr / indices pair"""
    live = block_liveness(mask, tile_o, tile_ci)
    counts = live.sum(axis=1)
    row_ptr = np.zeros(live.shape[0] + 1, dtype=np.int32)
    row_ptr[1:] = np.cumsum(counts)
    kblk = np.nonzero(live)[1].astype(np.int32)     # row-major: ascending kb inside each ot
    return row_ptr, kblk


def swizzle_offset(row, chunk, pitch):
    """byte offset of 16-byte chunk `chunk` of row `row` in a K-major tile with `pitch`-byte rows under
    the 32/64/128-byte TMA/UMMA swizzle: address bits [4,4+B) ^= bits [7,7+B), B = log2(pitch/16)"""
    off = row * pitch + chunk * 16
    return off ^ (((off >> 7) & ((pitch >> 4) - 1)) << 4)


def pack_weights(w, mask, tile_o, tile_ci, row_ptr, kblk, act_dtype):
    """uint16 image of the packed live tiles, tile j = tile_o rows x tile_ci K-elements (swizzled)"""
    O, I, kh, kw = w.shape
    taps = kh * kw
    wm = np.asarray(w, dtype=np.float32)
    if mask is not None:
        wm = np.where(np.asarray(mask) != 0, wm, np.float32(0))
    bits = to_bits(wm, act_dtype).reshape(O, I, taps)
    pitch = tile_ci * 2
    out = np.zeros(len(kblk) * tile_o * tile_ci, dtype=np.uint16)
    rows = np.arange(tile_o)
    for ot in range(len(row_ptr) - 1):
        for j in range(row_ptr[ot], row_ptr[ot + 1]):
            cib, tap = divmod(int(kblk[j]), taps)
            tile = bits[ot * tile_o:(ot + 1) * tile_o, cib * tile_ci:(cib + 1) * tile_ci, tap]
            base = j * tile_o * tile_ci
            for chunk in range(tile_ci // 8):
                offs = np.array([swizzle_offset(int(r), chunk, pitch) for r in rows]) // 2
                for e in range(8):
                    out[base + offs + e] = tile[:, chunk * 8 + e]
    return out


def expand_tile_list(row_ptr, kblk, O, I, kh, kw, tile_o, tile_ci):
    """dense {0,1} mask covered by the tile list (superset of the element mask)"""
    taps = kh * kw
    m = np.zeros((O, I, taps), dtype=np.float32)
    for ot in range(len(row_ptr) - 1):
        for j in range(row_ptr[ot], row_ptr[ot + 1]):
            cib, tap = divmod(int(kblk[j]), taps)
            m[ot * tile_o:(ot + 1) * tile_o, cib * tile_ci:(cib + 1) * tile_ci, tap] = 1
    return m.reshape(O, I, kh, kw)


# ------------------------------------------------------------------ the reference's own BSR exporter

def bsr_from_dense(mat, bh, bw):
    """BlockPruner.generate_block_matrix for (bh,bw) != (1,1) on an exactly tiled matrix
    (pruners/BlockPruner.py:366-404): values column-major inside a block, row-major over live blocks."""
    rows, cols = mat.shape
    nrb, ncb = rows // bh, cols // bw
    blocks = mat.reshape(nrb, bh, ncb, bw).transpose(0, 2, 1, 3)          # [nrb, ncb, bh, bw]
    live = np.abs(blocks).sum(axis=(2, 3)) != 0
    rb, cb = np.nonzero(live)
    values = np.concatenate([blocks[r, c].flatten("F") for r, c in zip(rb, cb)]) if len(rb) else \
        np.zeros(0, dtype=mat.dtype)
    ptr = np.zeros(nrb + 1, dtype=int)
    ptr[1:] = np.cumsum(live.sum(axis=1))
    return values, cb.astype(int), ptr


def bsr_text(rows, cols, bh, bw, values, indices, ptr):
    """BlockPruner.write_block_matrix_to_file format (pruners/BlockPruner.py:416-432)"""
    def line(a):
        return "".join(str(e) + " " for e in a) + "\n"
    return "%d\n%d\n%d\n%d\n%d\n" % (rows, cols, bh, bw, ptr[-1]) + line(values) + line(indices) + line(ptr)
