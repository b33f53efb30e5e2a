"""Golden fixtures for the blocklet text exporters, produced by the REAL reference:
`RmbPruner.prune_tensor_as_rmb(..., dump_fpath)` (pruners/RmbPruner.py:247-378) and
`RmcdbPruner.prune_tensor_as_rmcdb(..., dump_fpath)` (pruners/RmcdbPruner.py:320-439) on the seeded weights of
gen_golden.py.  The text files are stored zlib-compressed inside blocklet_export.npz.

Run in the build container only:   python tests/golden/gen_golden_export.py
"""
import os
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import import_reference, quiet  # noqa: E402


def seeded_weight(shape, seed):
    import torch
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)).numpy()


def main():
    import_reference()
    from pruners.RmbPruner import RmbPruner, RmbPrunerConfig, BlockletType
    from pruners.RmcdbPruner import RmcdbPruner, RmcdbPrunerConfig
    w_a = seeded_weight((64, 32, 3, 3), 11)
    w_b = seeded_weight((32, 64, 1, 1), 12)
    cases = {
        "rmb_a": (RmbPruner.prune_tensor_as_rmb, w_a, RmbPrunerConfig(32, 72, 0.5, [BlockletType(8, 9)], [2])),
        "rmb_b": (RmbPruner.prune_tensor_as_rmb, w_b,
                  RmbPrunerConfig(16, 32, 0.5, [BlockletType(4, 4), BlockletType(2, 8)], [1, 1])),
        "rmcdb_a": (RmcdbPruner.prune_tensor_as_rmcdb, w_a,
                    RmcdbPrunerConfig(32, 72, 0.5, [BlockletType(8, 9)], [2], True)),
        "rmcdb_b": (RmcdbPruner.prune_tensor_as_rmcdb, w_b,
                    RmcdbPrunerConfig(16, 32, 0.0, [BlockletType(4, 4), BlockletType(2, 8)], [1, 2], True)),
    }
    out = {}
    for name, (fn, w, cfg) in cases.items():
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, name + ".txt")
            quiet(fn, w, cfg, path)
            data = open(path, "rb").read()
        out[name] = np.frombuffer(zlib.compress(data, 9), dtype=np.uint8)
        print(name, len(data), "bytes ->", out[name].size)
    np.savez(os.path.join(HERE, "blocklet_export.npz"), **out)


if __name__ == "__main__":
    main()
