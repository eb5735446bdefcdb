"""Operator-level `Nodes` API on the GPU (Nodes.extract / sort / clear / copy / reset_scores / score through
pgpu_extract_nodes and pgpu_score_nodes), against the oracle.  Named to run after the other GPU suites."""
import pytest

import nodes_api_cases as cases

pytestmark = pytest.mark.gpu


def test_nodes_api_gpu():
    import pyrodigal_b200.lib as L
    cases.run_cases(L)
