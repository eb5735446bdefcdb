"""CPU check of the bit-parallel node extraction (pyrodigal_b200/csrc/extract_device.cuh) without a GPU:
tests/emu/extract_emu.cu runs the per-word function the kernel k_extract_b is made of over every (strand, frame,
word) in host loops; the node set must equal the oracle's `orc_extract` (itself pinned against the reference).
Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "extract_emu.cu")
LIB = os.path.join(HERE, "emu", "libextract_emu.so")
CSRC = os.path.join(R.ROOT, "pyrodigal_b200", "csrc")


def _emu():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("extract_device.cuh", "codon_masks.hpp", "common.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(p) for p in deps):
        subprocess.check_call(["nvcc", "-x", "cu", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets",
                               "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC])
    lib = C.CDLL(LIB)
    lib.emu_extract.restype = C.c_int
    return lib


def emu_extract(d, tt=11, closed=False, min_gene=90, min_edge_gene=60):
    cap = max(4096, len(d))
    out = np.zeros((cap, 5), dtype=np.int32)
    n = _emu().emu_extract(d.ctypes.data_as(C.c_void_p), len(d), tt, int(closed), min_gene, min_edge_gene,
                           out.ctypes.data_as(C.c_void_p), cap)
    assert n >= 0
    out = out[:n]
    order = np.lexsort((-out[:, 3], out[:, 0]))  # (ndx, forward strand first): compare_nodes, node.c:1578-1587
    return out[order]


def check(seq, tt=11, closed=False, min_gene=90, min_edge_gene=60):
    d, _, _ = orc.encode(seq)
    want = orc.extract(d, tt, orc.make_opts(closed=closed, min_gene=min_gene, min_edge_gene=min_edge_gene))
    got = emu_extract(d, tt, closed, min_gene, min_edge_gene)
    w = np.stack([want["ndx"], want["stop_val"], want["type"], want["strand"], want["edge"]], axis=1).astype(np.int32)
    assert len(got) == len(w), (len(got), len(w))
    assert np.array_equal(got, w), np.nonzero((got != w).any(axis=1))[0][:5]
    return len(w)


@pytest.mark.parametrize("closed", [False, True])
@pytest.mark.parametrize("tt", [11, 4])
@pytest.mark.parametrize("length,gc,seed", [(10000, 0.5, 1234), (3001, 0.35, 7), (25000, 0.66, 11), (120000, 0.5, 3)])
def test_random_contigs(length, gc, seed, tt, closed):
    assert check(R.synth(length, gc, seed), tt=tt, closed=closed) > 0


@pytest.mark.parametrize("closed", [False, True])
@pytest.mark.parametrize("length", list(range(0, 40)) + [89, 90, 91, 92, 93, 94, 95, 96, 97, 98, 99, 127, 128, 129, 191, 192, 193, 1000])
def test_short_sequences(length, closed):
    for seed in range(3):
        check(R.synth(length, 0.5, 100 + seed), closed=closed)


@pytest.mark.parametrize("closed", [False, True])
def test_no_stop_codons_and_unknown_bases(closed):
    # a long ORF without any stop (virtual stop only), Ns sprinkled in, and a sequence of only stops
    check("ATG" + "GCC" * 400 + "ATGGCC" * 50, closed=closed)
    check("CAT" * 300 + "GGC" * 500, closed=closed)
    check(R.synth(20000, 0.5, 5, n_frac=0.01), closed=closed)
    check("TAA" * 200, closed=closed)
    check("N" * 500, closed=closed)
    check("ATG" * 333, closed=closed)


@pytest.mark.parametrize("min_gene,min_edge_gene", [(1, 1), (3, 3), (4, 2), (30, 10), (60, 60), (91, 61), (300, 150), (1000, 90)])
def test_length_thresholds(min_gene, min_edge_gene):
    for closed in (False, True):
        for seed in range(2):
            check(R.synth(30000, 0.6, 40 + seed), closed=closed, min_gene=min_gene, min_edge_gene=min_edge_gene)


def test_high_gc_long_orfs():
    # GC-rich sequence: long ORFs spanning many words, so the look-back crosses several words
    check(R.synth(200000, 0.75, 9))
    check(R.synth(200000, 0.75, 9), tt=4, closed=True)


def test_randomised_parameters():
    """seeded sweep over length, GC content, unknown-base fraction, end mode, translation table and the two length
    thresholds (the word function's masks depend on all of them)"""
    rng = np.random.default_rng(20261017)
    tables = [1, 2, 4, 11, 22, 23, 25, 33]
    for it in range(160):
        length = int(rng.integers(0, 4000)) if it % 4 else int(rng.integers(0, 140))
        seq = R.synth(length, float(rng.uniform(0.2, 0.8)), 50000 + it, n_frac=float(rng.choice([0.0, 0.0, 0.01, 0.1])))
        mg = int(rng.choice([1, 2, 3, 4, 5, 6, 29, 30, 31, 60, 89, 90, 91, 92, 93, 96, 99, 180, 300, 2000]))
        meg = int(rng.choice([1, 2, 3, 4, 30, 59, 60, 61, 62, 63, 90, 120, 600]))
        check(seq, tt=int(rng.choice(tables)), closed=bool(rng.integers(0, 2)), min_gene=mg, min_edge_gene=meg)


@pytest.mark.parametrize("tt", [1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 21, 22, 23, 24, 25, 26, 29, 30, 32, 33])
def test_codon_flag_table_equals_mask_arithmetic(tt):
    """the optional byte table of k_codon_bits (PGPU_CODON_LUT=1) holds exactly what codon_flags computes"""
    assert _emu().emu_lut_mismatches(tt) == 0
