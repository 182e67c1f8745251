"""Operator vocabulary of the search space.

`PRIMITIVES` and `Genotype` are the names the reference's scripts import (search/genotypes.py:1-11); the position of a name
in `PRIMITIVES` is the op id stored in the alpha tensors and in the shipped `arch_*.pt` genotypes, so the order is part of
the checkpoint format.  Everything else here is ours: static facts about each primitive that the engine, the schedule
compiler and the tools use instead of re-deriving them from class names.
"""
from collections import namedtuple

Genotype = namedtuple("Genotype", "normal normal_concat reduce reduce_concat")

OpInfo = namedtuple("OpInfo", "name cls convs zoomed identity_at_stride1")
#                      op id -> (name, operator class in operations.py, 3x3 convs, runs at half resolution, param-free when s=1)
OP_TABLE = (
    OpInfo("skip", "FactorizedReduce", 0, False, True),
    OpInfo("conv", "BasicResidual1x", 1, False, False),
    OpInfo("conv_downup", "BasicResidual_downup_1x", 1, True, False),
    OpInfo("conv_2x", "BasicResidual2x", 2, False, False),
    OpInfo("conv_2x_downup", "BasicResidual_downup_2x", 2, True, False),
)
PRIMITIVES = [info.name for info in OP_TABLE]


def op_index(name: str) -> int:
    """op id of a primitive name (raises ValueError for unknown names)"""
    return PRIMITIVES.index(name)


def describe(ops, path, widths=None) -> str:
    """One line per layer of a decoded branch: `layer scale op [width]` -- used by tools/ and error messages."""
    rows = []
    for layer, (op, scale) in enumerate(zip(ops, path)):
        info = OP_TABLE[int(op)]
        width = "" if widths is None or layer >= len(widths) else "  w=%.3f" % float(widths[layer])
        rows.append("%2d  1/%-2d  %-15s%s" % (layer, 8 * 2 ** int(scale), info.name, width))
    return "\n".join(rows)
