"""pyrodigal_b200 -- B200-native (sm_100a CUDA) implementation of Pyrodigal's find_genes hot path.

Drop-in for `pyrodigal` on that path: GeneFinder / Genes / Gene / Nodes / Node / Sequence / TrainingInfo /
MetagenomicBins keep the reference's names and attributes (src/pyrodigal/__init__.py, lib.pyi).  All
compute runs in libpyrodigal_b200.so on the GPU; there is no CPU fallback.
"""
from . import lib
from .lib import (
    ConnectionScorer,
    Gene,
    GeneFinder,
    Genes,
    Mask,
    Masks,
    MetagenomicBin,
    MetagenomicBins,
    Node,
    Nodes,
    Sequence,
    TrainingInfo,
    MIN_SINGLE_GENOME,
    IDEAL_SINGLE_GENOME,
    TRANSLATION_TABLES,
)

__version__ = "0.1.0"


def __getattr__(name):
    if name == "METAGENOMIC_BINS":
        return lib._LazyBins.get()
    raise AttributeError(name)
