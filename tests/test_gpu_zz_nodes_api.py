"""Operator-level `Nodes` API on the GPU (Nodes.extract / sort / clear / copy / reset_scores / score through
pgpu_extract_nodes and pgpu_score_nodes), against the oracle.  Named to run after the other GPU suites."""
import pytest

import nodes_api_cases as cases

pytestmark = pytest.mark.gpu


def test_nodes_api_gpu():
    import pyrodigal_b200.lib as L
    cases.run_cases(L)


def test_sequence_operators_gpu():
    import pyrodigal_b200.lib as L
    cases.run_sequence_operator_cases(L)


@pytest.mark.parametrize("tt", [11, 4, 25])
def test_codon_table_variant_of_k_codon_bits(monkeypatch, tt):
    """PGPU_CODON_LUT=1 (off by default): the byte-table variant of k_codon_bits must extract the same nodes"""
    import numpy as np
    import refutil as R
    from oracle import oracle as orc
    from pyrodigal_b200 import _capi
    monkeypatch.setenv("PGPU_CODON_LUT", "1")
    ctx = _capi.Context(0)
    ctx.set_models(R.bins_blob(), 50)
    try:
        for length, gc, seed, nfrac, closed in ((40000, .5, 1, 0.0, False), (3001, .62, 3, 0.0, True), (20000, .45, 4, .002, False)):
            seq = R.synth(length, gc, seed, n_frac=nfrac)
            d, _, _ = orc.encode(seq)
            want = orc.extract(d, tt, orc.make_opts(closed=closed))
            got = ctx.extract_nodes(np.frombuffer(seq, np.uint8), tt, _capi.make_opts(closed=closed))
            for f in ("ndx", "stop_val", "strand", "type", "edge"):
                assert np.array_equal(got[f], want[f]), (tt, length, f)
    finally:
        ctx.close()
