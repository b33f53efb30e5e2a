"""Static Kronecker-structured masks, ``mask = OB (x) (CB (x) P) (x) IB``
(reference: pruners/SRMBRepMasker.py:33-383; the only pruner type the shipped optimal_configs/ use).

  OB : outer-block pattern, (rows/obh) x (cols/obw), sparsity ``osp`` / pattern ``opat``
  CB : all-ones (obh/cbh) x (obw/cbw) — repeats the core pattern inside an outer block
  P  : core pattern, (cbh/ibh) x (cbw/ibw), sparsity ``isp`` / pattern ``ipat``
  IB : all-ones inner block ibh x (ibw * kernel_size)
With ``is_repetitive`` one P is shared by all outer blocks, otherwise every live outer block draws its own.
All randomness goes through the global ``np.random`` stream in the same order as the reference, so
``np.random.seed(s)`` reproduces the reference's masks bit for bit.
"""
import collections
import json

import numpy as np

from .Pruner import Pruner


class SRMBRepMaskerConfig():
    def __init__(self, obh, obw, cbh, cbw, ibh, ibw, osp, opat, isp, ipat, is_repetitive,
                 collapse_tensor, cross_prob, is_symmetric):
        self.obh, self.obw = obh, obw
        self.cbh, self.cbw = cbh, cbw
        self.ibh, self.ibw = ibh, ibw
        self.osp, self.opat = osp, opat
        self.isp, self.ipat = isp, ipat
        self.is_repetitive = is_repetitive
        self.collapse_tensor = collapse_tensor
        self.cross_prob = cross_prob
        self.is_symmetric = is_symmetric


_FIELDS = ("obh", "obw", "cbh", "cbw", "ibh", "ibw", "osp", "opat", "isp", "ipat", "is_repetitive",
           "collapse_tensor", "cross_prob", "is_symmetric")


class SRMBRepMasker(Pruner):
    def __init__(self, config_fp, on_gpu=True):
        super(SRMBRepMasker, self).__init__(config_fp, on_gpu)

    def parse_config_file(self, config_fp):
        """extra keys of the shipped configs (make_kwargs, exec_args) are ignored, as in the reference"""
        layer_configs = collections.OrderedDict()
        with open(config_fp) as fh:
            data = json.load(fh)
        for entry in data["configs"]:
            for layer in entry["layer_set"]:
                layer_configs[layer] = SRMBRepMaskerConfig(*[entry[f] for f in _FIELDS])
        return layer_configs

    def generate_masks(self, model, is_static=True, verbose=False):
        sd = model.state_dict()
        for layer, cfg in self.layer_configs.items():
            mask = SRMBRepMasker.construct_mask(sd[layer].cpu().numpy(), cfg)
            if verbose:
                print("Generated mask for layer {}".format(layer))
            self._store(layer, mask)

    # ------------------------------------------------------------------------------ patterns
    @staticmethod
    def get_ramanujan_pattern(rows, cols, d, cross_prob=0.5, is_symmetric=False, debug=False):
        """d-regular bipartite pattern grown by repeated 2-lifts with random edge crossings (:103-168)"""
        assert cols % d == 0
        assert (cols // d) & (cols // d - 1) == 0
        assert rows // (cols // d) > 0
        if is_symmetric:
            assert rows == cols, "When symmetric, #rows = #cols"
        mask = np.zeros((rows, cols), dtype=int)
        cr, cc = rows // (cols // d), d
        mask[:cr, :cc] = 1
        while cc < cols:
            mask[cr:2 * cr, cc:2 * cc] = mask[:cr, :cc]          # clone the current graph
            for l in range(cr):
                for r in range(l if is_symmetric else 0, cc):
                    if mask[l, r] != 1:
                        continue
                    if np.random.binomial(1, cross_prob) != 1:
                        continue
                    mask[l, r] = 0                               # cross the edge pair
                    mask[l + cr, r + cc] = 0
                    mask[l, r + cc] = 1
                    mask[l + cr, r] = 1
                    if is_symmetric:
                        mask[r, l] = 0
                        mask[r + cc, l + cr] = 0
                        mask[r + cc, l] = 1
                        mask[r, l + cr] = 1
            cr, cc = 2 * cr, 2 * cc
        return mask

    @staticmethod
    def _trans_pattern(mask, M, N, nnz_per_row):
        """'TRANS': row- and column-regular random pattern (:195-245)"""
        if nnz_per_row <= int(0.25 * N):
            print("Truly random")
            xs = np.arange(M)
            for _ in range(nnz_per_row):
                while True:
                    ys = np.random.permutation(M)
                    if np.sum(mask[xs, ys]) == 0:
                        mask[xs, ys] = 1
                        break
            return mask
        mask += 1                                                 # start full, remove edges
        degree = np.ones(N, dtype=int) * M
        pool = np.arange(N)
        pool_size = N
        for u in range(M):
            chosen = np.zeros(N)
            for _ in range(N - nnz_per_row):
                cand = pool[:pool_size]
                cand_deg = degree[cand]
                top = np.where(cand_deg == np.max(cand_deg))[0]
                while True:
                    ind = top[np.random.randint(top.size)]
                    v = cand[ind]
                    if chosen[v] != 0:
                        continue
                    mask[u, v] = 0
                    chosen[v] = 1
                    degree[v] -= 1
                    if degree[v] == nnz_per_row:                 # v is done: swap it out of the pool
                        last = pool[pool_size - 1]
                        pool[pool_size - 1] = pool[ind]
                        pool[ind] = last
                        pool_size -= 1
                    break
        return mask

    @staticmethod
    def generate_sparsity_pattern(M, N, sparsity, pattern, cross_prob=0.5, is_symmetric=False):
        nnz = M * int((1.0 - sparsity) * N)
        per_row = nnz // M
        mask = np.zeros((M, N))
        if sparsity == 0:
            mask[:] = 1
            return mask
        rows = range(M)
        if pattern == "RANDOM":
            mask.reshape(M * N)[np.random.choice(M * N, nnz, replace=False)] = 1
        elif pattern == "UROW":
            assert nnz % M == 0
            for i in rows:
                mask[i, np.random.choice(N, per_row, replace=False)] = 1
        elif pattern == "RAMANUJAN":
            mask = SRMBRepMasker.get_ramanujan_pattern(M, N, per_row, cross_prob, is_symmetric)
        elif pattern == "TRANS":
            assert nnz % M == 0
            assert M == N, "Matrix should be square"
            mask = SRMBRepMasker._trans_pattern(mask, M, N, per_row)
        elif pattern in ("CDIA", "CDIASTRIDE", "CBAND", "CCDIA"):
            assert nnz % M == 0
            if pattern == "CDIA":
                base = np.random.choice(N, per_row, replace=False)
            elif pattern == "CDIASTRIDE":
                base = np.arange(0, N, N // per_row)
            elif pattern == "CBAND":
                base = (np.arange(-(per_row // 2), per_row // 2) + N) % N
            else:
                base = np.arange(per_row)
            for i in rows:
                mask[i, (i + base) % N] = 1
        elif pattern == "COLUMN":
            assert nnz % M == 0
            mask[:, np.random.choice(N, per_row, replace=False)] = 1
        elif pattern == "CCOLUMN":
            assert nnz % M == 0
            mask[:, :per_row] = 1
        elif pattern == "GROUP":
            groups = N // per_row
            sh = M // groups
            for g in range(groups):
                mask[g * sh:(g + 1) * sh, g * per_row:(g + 1) * per_row] = 1
        else:
            raise ValueError("Unsupported {}".format(pattern))
        return mask

    # ------------------------------------------------------------------------------ composition
    @staticmethod
    def construct_mask(tensor, config):
        rows, cols = tensor.shape[0], tensor.shape[1]
        ksize = tensor.size // (rows * cols)
        if config.collapse_tensor:
            cols *= ksize
            ksize = 1
        obh = rows if config.obh == -1 else config.obh
        obw = cols if config.obw == -1 else config.obw
        cbh = obh if config.cbh == -1 else config.cbh
        cbw = obw if config.cbw == -1 else config.cbw
        ibh, ibw = config.ibh, config.ibw
        pat = SRMBRepMasker.generate_sparsity_pattern
        outer = pat(rows // obh, cols // obw, config.osp, config.opat, config.cross_prob,
                    config.is_symmetric)
        repeat = np.ones((obh // cbh, obw // cbw), dtype=tensor.dtype)
        inner = np.ones((ibh, ibw * ksize), dtype=tensor.dtype)
        if config.is_repetitive:
            core = pat(cbh // ibh, cbw // ibw, config.isp, config.ipat, config.cross_prob,
                       config.is_symmetric)
            full = np.kron(np.kron(outer, np.kron(repeat, core)), inner)
            return full.reshape(tensor.shape).astype(tensor.dtype)
        grid = np.zeros((rows // ibh, cols // ibw), dtype=tensor.dtype)
        sh, sw = obh // ibh, obw // ibw
        for rb in range(rows // obh):
            for cb in range(cols // obw):
                if outer[rb, cb] == 1:
                    core = pat(cbh // ibh, cbw // ibw, config.isp, config.ipat, config.cross_prob,
                               config.is_symmetric)
                    grid[rb * sh:(rb + 1) * sh, cb * sw:(cb + 1) * sw] += np.kron(repeat, core)
        return np.kron(grid, inner).reshape(tensor.shape)
