"""Shared assertions for the operator-level `Nodes` API (reference: src/pyrodigal/tests/test_nodes.py,
test_connection_scorer.py:15-36).  Used twice: on the GPU through the real C ABI (tests/test_gpu_zz_nodes_api.py) and
without a GPU against a stand-in context that answers the two C-ABI operator calls with the oracle
(tests/test_nodes_api_cpu.py), so that the Python marshalling is exercised in the build container as well."""
import copy
import pickle

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

FIELDS_INT = ("ndx", "stop_val", "strand", "type", "edge")
FIELDS_SCORE = ("cscore", "sscore", "rscore", "tscore", "uscore", "gc_cont")


def check_extract(L, seq_bytes, tt=11, closed=False, min_gene=90, min_edge_gene=60, mask=False):
    seq = L.Sequence(seq_bytes, mask=mask)
    nodes = L.Nodes()
    assert len(nodes) == 0
    n = nodes.extract(seq, translation_table=tt, closed=closed, min_gene=min_gene, min_edge_gene=min_edge_gene)
    d, _, _ = orc.encode(seq_bytes)
    masks = orc.find_masks(d, 50) if mask else None
    want = orc.extract(d, tt, orc.make_opts(closed=closed, min_gene=min_gene, min_edge_gene=min_edge_gene, masks=masks))
    assert n == len(want) == len(nodes)
    for f in FIELDS_INT:
        assert np.array_equal(nodes.array[f], want[f]), f
    nodes.sort()  # already in (index, strand) order
    return seq, nodes, d, want


def run_cases(L):
    # --- extract / clear / copy / pickle (test_nodes.py:28-75) ---
    s = R.synth(30000, 0.5, 77)
    seq, nodes, d, want = check_extract(L, s)
    assert len(nodes) > 100
    n0 = len(nodes)
    c1, c2, c3 = nodes.copy(), copy.copy(nodes), pickle.loads(pickle.dumps(nodes))
    for c in (c1, c2, c3):
        assert len(c) == n0 and np.array_equal(c.array, nodes.array) and c.array is not nodes.array
    first = nodes[0]
    assert first.index == int(want["ndx"][0]) and first.strand == int(want["strand"][0]) and first.cscore == 0.0
    nodes.clear()
    assert len(nodes) == 0 and len(c1) == n0
    assert len(L.Nodes().copy()) == 0 and len(pickle.loads(pickle.dumps(L.Nodes()))) == 0
    # other tables, closed ends, thresholds, masks, tiny inputs
    check_extract(L, s, tt=4)
    check_extract(L, s, closed=True)
    check_extract(L, s, min_gene=30, min_edge_gene=20)
    check_extract(L, R.synth(20000, 0.45, 4, n_frac=0.01), mask=True)
    for tiny in (b"ATG", R.synth(89, 0.5, 1), R.synth(200, 0.5, 2)):
        check_extract(L, tiny)
    with pytest.raises(ValueError):
        L.Nodes().extract(L.Sequence(s), translation_table=7)
    # two extractions without clear() append, like the reference; sort() then refuses
    twice = L.Nodes()
    a = twice.extract(L.Sequence(s))
    b = twice.extract(L.Sequence(s))
    assert a == b and len(twice) == 2 * a
    with pytest.raises(NotImplementedError):
        twice.sort()

    # --- score / reset_scores (lib.pyx:2563-2589) ---
    for model, is_meta, closed in ((0, True, False), (11, False, False), (24, True, True)):
        blob = R.bin_blob(model)
        tinf = L.TrainingInfo._from_bytes(blob)
        tt = tinf.translation_table
        seq, nodes, d, ref = check_extract(L, s, tt=tt, closed=closed)
        for rep in range(2):  # a second score() on the same nodes sees the converted edge flags
            orc.reset_scores(ref)
            orc.score(d, ref, blob, closed=closed, is_meta=is_meta)
            nodes.reset_scores()
            assert not nodes.array["cscore"].any() and (nodes.array["traceb"] == -1).all()
            nodes.score(seq, tinf, closed=closed, is_meta=is_meta)
            for f in FIELDS_SCORE:
                assert np.array_equal(nodes.array[f], ref[f]), (model, rep, f)
            assert np.array_equal(nodes.array["edge"], ref["edge"]) and np.array_equal(nodes.array["rbs"], ref["rbs"])
            assert not nodes.array["star_ptr"].any()  # Nodes.score does not record overlapping starts
    other = L.Sequence(R.synth(30000, 0.5, 78))
    with pytest.raises(ValueError):
        nodes.score(other, tinf)


def run_sequence_operator_cases(L):
    """Sequence.shine_dalgarno / max_gc_frame_plot (reference: tests/test_sequence.py:52-75 and the oracle)"""
    tinf = L.TrainingInfo._from_bytes(R.bin_blob(0))
    seq = L.Sequence("AGGAGGTTAGCAAATATG")
    for i in range(10):
        assert seq.shine_dalgarno(i, 15, tinf) == (24 if i == 0 else 13 if i == 3 else 0), i
        assert seq.shine_dalgarno(i, 15, tinf, exact=False) == 0, i
    seq = L.Sequence("AGGTGGTTAGCAAATATG")
    for i in range(10):
        assert seq.shine_dalgarno(i, 15, tinf) == (6 if i == 0 else 0), i
        assert seq.shine_dalgarno(i, 15, tinf, exact=False) == (19 if i == 0 else 0), i
    for bad in (dict(pos=-1, start=5), dict(pos=1, start=-5), dict(pos=1, start=9, strand=0)):
        with pytest.raises(ValueError):
            seq.shine_dalgarno(training_info=tinf, **bad)
    # seeded sweep against the oracle: both strands, both modes, windows anywhere around the start
    rng = np.random.default_rng(11)
    s = R.synth(600, 0.45, 90) + b"AGGAGGTAGGAGGAAGGAGNNGGAGG" * 4 + R.synth(300, 0.6, 91)
    d, _, _ = orc.encode(s)
    seq = L.Sequence(s)
    for model in (0, 24):
        blob = R.bin_blob(model)
        tinf = L.TrainingInfo._from_bytes(blob)
        rbs_wt = np.ascontiguousarray(tinf.rbs_weights, dtype=np.float64)
        for _ in range(120):
            start = int(rng.integers(0, len(s)))
            pos = max(0, start - int(rng.integers(0, 30)))
            strand = int(rng.choice([1, -1]))
            for exact in (True, False):
                want = orc.shine_dalgarno(d, pos, start, rbs_wt, strand=strand, exact=exact)
                assert seq.shine_dalgarno(pos, start, tinf, strand=strand, exact=exact) == want, (model, pos, start, strand, exact)
    # GC frame plot
    for text in (R.synth(5000, 0.5, 3), R.synth(361, 0.3, 4), R.synth(1000, 0.7, 5, n_frac=0.02), b"ACGTAC", b"GC"):
        d, _, _ = orc.encode(text)
        plot = L.Sequence(text).max_gc_frame_plot()
        assert plot.typecode == "i" and list(plot) == orc.gc_frame_plot(d).astype(int).tolist(), len(text)
    with pytest.raises(ValueError):
        L.Sequence(b"ACGT").max_gc_frame_plot(window_size=-1)
