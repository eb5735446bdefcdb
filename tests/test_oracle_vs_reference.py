"""CPU tests, build container only: the C oracle against the LIVE unmodified reference
(oracle/_ref, built by oracle/build_ref.sh).  Skipped where the reference build is absent."""
import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

pytestmark = pytest.mark.skipif(not R.have_reference(), reason="oracle/_ref (reference build) not present")

NODE_INTS = R.INT_FIELDS + ("rbs", "mot_ndx", "mot_len", "mot_spacer", "mot_spacendx")
NODE_FLOATS = R.FLOAT_FIELDS + ("gc_cont",)


def check_meta(seq, closed=False, mask=False, what=""):
    pyrodigal = R.reference()
    g = pyrodigal.GeneFinder(meta=True, closed=closed, mask=mask).find_genes(seq)
    d, gc, unk = orc.encode(seq)
    assert bytes(d) == bytes(memoryview(g.sequence)) and unk == g.sequence.unknown
    masks = orc.find_masks(d, 50) if mask else None
    if mask:
        assert [(m.begin, m.end) for m in g.sequence.masks] == [tuple(x) for x in masks]
    genes, nodes, winner, _ = orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, R.bins_blob(),
                                                  orc.make_opts(closed=closed, masks=masks))
    ref_w = list(pyrodigal.METAGENOMIC_BINS).index(g.metagenomic_bin) if g.metagenomic_bin is not None else -1
    assert winner == ref_w
    rg = [(x["begin"], x["end"], x["start_ndx"], x["stop_ndx"]) for x in g.__getstate__()["genes"]]
    assert rg == [tuple(int(v) for v in x) for x in genes]
    R.assert_nodes_equal(nodes, R.ref_nodes_to_array(g.nodes), ints=NODE_INTS, floats=NODE_FLOATS, what=what)


@pytest.mark.parametrize("length,gc", [(10000, .5), (3000, .4), (1200, .6), (40000, .35), (40000, .65), (200, .5),
                                        (20, .5), (2, .5), (0, .5), (2999, .5), (1500, .5), (60000, .3)])
@pytest.mark.parametrize("closed", [False, True])
def test_meta_synthetic(length, gc, closed):
    check_meta(R.synth(length, gc, seed=1000 + length), closed=closed, what=f"L{length} gc{gc}")


@pytest.mark.parametrize("mask", [False, True])
def test_meta_with_unknown_bases(mask):
    check_meta(R.synth(30000, .45, seed=77, n_frac=0.002), mask=mask, what="N")


@pytest.mark.parametrize("fn", ["KK037166", "SRR492066"])
@pytest.mark.parametrize("mode", ["plain", "mask", "closed"])
def test_meta_real_contigs(fn, mode):
    _, s = R.read_fasta_gz(f"{R.REF_DATA}/{fn}.fna.gz")[0]
    check_meta(s, closed=mode == "closed", mask=mode == "mask", what=fn)


def test_bins_blob_matches_reference():
    pyrodigal = R.reference()
    for i, b in enumerate(pyrodigal.METAGENOMIC_BINS):
        assert bytes(memoryview(b.training_info)) == R.bin_blob(i)


@pytest.mark.parametrize("seed", range(5))
def test_random_short_contigs_all_modes(seed):
    rng = np.random.default_rng(seed)
    for _ in range(8):
        L = int(rng.integers(90, 6000))
        gc = float(rng.uniform(.25, .75))
        check_meta(R.synth(L, gc, seed=int(rng.integers(1 << 30))), closed=bool(rng.integers(2)), what=f"rand{L}")


@pytest.mark.parametrize("seed", range(30))
def test_masks_follow_the_reference_cursor(seed):
    """GeneFinder(mask=True): the reference tests a candidate ORF only against the mask under a per-frame cursor
    (lib.pyx:1959-1966 / 2053-2061), not against every mask (an "intersects any mask" rule fails 5 of these 30)"""
    check_meta(R.trailing_n_case(seed), mask=True, closed=False, what=f"masks{seed}")
