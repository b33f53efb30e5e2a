"""Host-side mirror of the reference's ``pruners`` package (mask sources for the tile-list builder).

Same class names, constructor arguments, config-file schema and ``mask_dict`` contract as
``/pruners`` of the reference, so a script written against ``from pruners.BlockPruner import BlockPruner``
only changes its import.  Masks are produced on the host with numpy exactly as in the reference
(one-time work); what is new is what happens to them afterwards (drnb200.engine).
"""
from .Pruner import Pruner
from .BlockPruner import BlockPruner, BlockPrunerConfig, BlockMatrix
from .HbPruner import HbPruner, HbPrunerConfig
from .GroupingPruner import GroupingPruner, GroupingPrunerConfig
from .RmbPruner import RmbPruner, RmbPrunerConfig, BlockletType
from .RmcdbPruner import RmcdbPruner, RmcdbPrunerConfig
from .SRMBRepMasker import SRMBRepMasker, SRMBRepMaskerConfig

__all__ = ["Pruner", "BlockPruner", "BlockPrunerConfig", "BlockMatrix", "HbPruner", "HbPrunerConfig",
           "GroupingPruner", "GroupingPrunerConfig", "RmbPruner", "RmbPrunerConfig", "BlockletType",
           "RmcdbPruner", "RmcdbPrunerConfig", "SRMBRepMasker", "SRMBRepMaskerConfig", "make_pruner"]


def make_pruner(config_fp, on_gpu=True):
    """pruner_type dispatch of the drivers (semantic_seg.py:830-846)."""
    import json
    with open(config_fp) as fh:
        ptype = json.load(fh)["pruner_type"]
    table = {"block": BlockPruner, "hb": HbPruner, "grouping": GroupingPruner, "rmb": RmbPruner,
             "rmcdb": RmcdbPruner, "srmbrep": SRMBRepMasker}
    if ptype not in table:
        raise ValueError("Invalid type of pruner: %r" % ptype)
    return table[ptype](config_fp, on_gpu)
