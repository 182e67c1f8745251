"""Operator vocabulary of the search space (reference: search/genotypes.py:1-11)."""
from collections import namedtuple

Genotype = namedtuple("Genotype", "normal normal_concat reduce reduce_concat")

# index = op id stored in the alpha tensors / arch_*.pt genotypes
PRIMITIVES = ["skip", "conv", "conv_downup", "conv_2x", "conv_2x_downup"]
