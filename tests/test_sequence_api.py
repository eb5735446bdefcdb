"""Sequence / Masks mirror (lib.pyx:345-1075): counts, masks, start / stop codon probabilities, pickling (including a
state produced by the reference), checked against the reference's own objects when its build is present."""
import copy
import pickle

import pytest

import refutil as R

L = pytest.importorskip("pyrodigal_b200.lib")
CASES = [(R.synth(5000, 0.4, 3), False), (R.synth(20000, 0.6, 4, n_frac=0.01), True), (b"", False), (b"NNNN", False),
         (b"acgtRYKM" * 10, True)]


@pytest.mark.parametrize("seq,mask", CASES)
def test_sequence_round_trips(seq, mask):
    a = L.Sequence(seq, mask=mask)
    c = pickle.loads(pickle.dumps(a))
    assert bytes(c) == bytes(a) and str(c) == str(a) and len(c) == len(a)
    assert c.gc == a.gc and c.unknown == a.unknown and c.masks == a.masks
    assert pickle.loads(pickle.dumps(a.masks)) == a.masks and copy.copy(a.masks) == a.masks
    assert 0.0 <= a.start_probability() <= 1.0 and 0.0 <= a.stop_probability() <= 1.0
    assert a.__sizeof__() >= len(a)


@pytest.mark.skipif(not R.have_reference(), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("seq,mask", CASES)
def test_sequence_against_reference(seq, mask):
    ref = R.reference()
    a, b = L.Sequence(seq, mask=mask), ref.Sequence(seq.decode(), mask=mask)
    assert bytes(a) == bytes(memoryview(b)) and str(a) == str(b)
    assert (a.gc, a.gc_known, a.unknown) == (b.gc, b.gc_known, b.unknown)
    assert a.start_probability() == b.start_probability() and a.stop_probability() == b.stop_probability()
    assert a.masks.__getstate__() == b.masks.__getstate__()
    st = dict(b.__getstate__())
    st["masks"] = L.Masks(L.Mask(m.begin, m.end) for m in st["masks"])
    d = L.Sequence.__new__(L.Sequence)
    d.__setstate__(st)
    assert bytes(d) == bytes(a) and d.gc == a.gc and d.masks == a.masks
