"""CPU tests of the TRAINING oracle (oracle/pyrodigal_oracle.c: orc_train and its pieces) against
(1) tests/golden/train_cases.npz, generated from the unmodified reference (incl. the reference's own
training golden, tests/test_training_info.py:60-66, and the scalars of tests/test_gene_finder.py:329-345)
and (2) the live reference build where present.  Tolerance 0: the struct must be byte-identical."""
import os
import warnings

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAIN = np.load(os.path.join(G, "train_cases.npz"), allow_pickle=True)


def oracle_train(seq, closed=False, mask=False, force_nonsd=False, tt=11, st_wt=4.35):
    d, gc, unk = orc.encode(seq)
    masks = orc.find_masks(d, 50) if mask else None
    return orc.train(d, gc / len(d), translation_table=tt, start_weight=st_wt, force_nonsd=force_nonsd,
                     opts=orc.make_opts(closed=closed, masks=masks))


def diff_fields(a, b):
    x, y = (np.frombuffer(v, dtype=orc.TRAINING_DTYPE)[0] for v in (a, b))
    return [f for f in orc.TRAINING_DTYPE.names if not np.array_equal(x[f], y[f], equal_nan=True)]


@pytest.mark.parametrize("name", list(TRAIN["names"]))
def test_train_golden(name):
    closed, mask, force, tt = (int(v) for v in TRAIN[name + "/opts"])
    blob = oracle_train(TRAIN[name + "/seq"].tobytes(), closed, mask, force, tt, float(TRAIN[name + "/start_weight"]))
    assert diff_fields(blob, TRAIN[name + "/tinf"].tobytes()) == []
    assert blob == TRAIN[name + "/tinf"].tobytes()


@pytest.mark.parametrize("name", list(TRAIN["names"]))
def test_gc_frame_plot_golden(name):
    d, _, _ = orc.encode(TRAIN[name + "/seq"].tobytes())
    assert np.array_equal(orc.gc_frame_plot(d), TRAIN[name + "/gc_frame"])


def test_published_training_scalars():
    """tests/test_gene_finder.py:329-345"""
    t = np.frombuffer(oracle_train(TRAIN["srr_contig/seq"].tobytes()), dtype=orc.TRAINING_DTYPE)[0]
    e = TRAIN["srr_expected"]
    assert t["trans_table"] == 11 and t["st_wt"] == 4.35 and t["uses_sd"] == 1
    assert [t["gc"], *t["bias"], *t["type_wt"]] == list(e)


@pytest.mark.skipif(not R.have_reference(), reason="oracle/_ref (reference build) not present")
@pytest.mark.parametrize("length,gc,kw", [
    (20000, .5, {}), (45000, .3, {}), (45000, .7, dict(closed=True)), (80000, .52, dict(force_nonsd=True)),
    (33333, .41, dict(tt=4)), (50000, .6, dict(st_wt=2.5)), (30001, .5, dict(mask=True, n_frac=0.003)),
    (30002, .5, dict(n_frac=0.003)),
])
def test_train_vs_live_reference(length, gc, kw):
    pyrodigal = R.reference()
    kw = dict(kw)
    seq = R.synth(length, gc, seed=7000 + length, n_frac=kw.pop("n_frac", 0.0))
    gf = pyrodigal.GeneFinder(closed=kw.get("closed", False), mask=kw.get("mask", False))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ti = gf.train(seq, force_nonsd=kw.get("force_nonsd", False), translation_table=kw.get("tt", 11),
                      start_weight=kw.get("st_wt", 4.35))
    mine = oracle_train(seq, **kw)
    assert diff_fields(mine, bytes(memoryview(ti))) == []
    assert mine == bytes(memoryview(ti))


@pytest.mark.skipif(not R.have_reference(), reason="oracle/_ref (reference build) not present")
def test_train_multi_contig_linker():
    """several training sequences are joined by TTAATTAATTAA linkers, one trailing (lib.pyx:5534-5541)"""
    pyrodigal = R.reference()
    parts = [R.synth(15000, .5, seed=s) for s in (1, 2, 3)]
    ti = pyrodigal.GeneFinder().train(*parts)
    joined = b"TTAATTAATTAA".join(parts + [b""])
    assert oracle_train(joined) == bytes(memoryview(ti))
