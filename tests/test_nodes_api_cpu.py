"""CPU check of the Python marshalling of the operator-level `Nodes` API: the two C-ABI operator calls it makes
(`Context.extract_nodes`, `Context.score_nodes`) are answered by the oracle through a stand-in context, everything
else is the product's `pyrodigal_b200.lib`.  Test infrastructure only; the same assertions run against the real
library in tests/test_gpu_zz_nodes_api.py."""
import threading

import numpy as np
import pytest

import nodes_api_cases as cases
from oracle import oracle as orc


class OracleContext:
    def __init__(self, blob=None):
        self.blob, self.lock = blob, threading.Lock()
        self.first_pass_seen = []

    def _extract(self, seq, tt, opts):
        d, _, _ = orc.encode(bytes(seq))
        masks = orc.find_masks(d, opts.min_mask) if opts.mask else None
        o = orc.make_opts(closed=bool(opts.closed), min_gene=opts.min_gene, min_edge_gene=opts.min_edge_gene, masks=masks)
        return d, orc.extract(d, tt, o)

    def extract_nodes(self, seq, translation_table, opts):
        _, nodes = self._extract(seq, translation_table, opts)
        return {k: nodes[k].astype(t) for k, t in (("ndx", np.int32), ("stop_val", np.int32), ("strand", np.int8),
                                                   ("type", np.uint8), ("edge", np.uint8))}

    def score_nodes(self, seq, model, opts, is_meta=False, first_pass=True):
        from pyrodigal_b200 import _capi
        assert model == 0
        self.first_pass_seen.append(first_pass)
        tt = int(np.frombuffer(self.blob, np.int32, count=1, offset=8)[0])
        d, nodes = self._extract(seq, tt, opts)
        for _ in range(1 if first_pass else 2):   # a later pass sees the edge flags converted by the first one
            orc.reset_scores(nodes)
            orc.score(d, nodes, self.blob, closed=bool(opts.closed), is_meta=is_meta)
        orc.record_overlapping_starts(nodes, self.blob, flag=1, max_overlap=opts.max_overlap)
        out = np.zeros(len(nodes), dtype=_capi.NODE_DTYPE)
        for f in ("ndx", "stop_val", "strand", "type", "edge", "rbs", "mot_ndx", "mot_len", "mot_spacer", "mot_spacendx",
                  "mot_score", "cscore", "uscore", "tscore", "rscore", "sscore", "gc_cont", "star_ptr"):
            out[f] = nodes[f]
        return out


def test_nodes_api_marshalling(monkeypatch):
    L = pytest.importorskip("pyrodigal_b200.lib")
    made = []

    def context_for(blob, n_models, device=0):
        if n_models != 1:          # the built-in model set: used for extraction only
            return OracleContext()
        made.append(OracleContext(bytes(blob)))
        return made[-1]

    monkeypatch.setattr(L, "_context_for", context_for)
    cases.run_cases(L)
    # the first score() of freshly extracted nodes is a first pass, the second one is not
    assert [c.first_pass_seen for c in made[:2]] == [[True], [False]]
