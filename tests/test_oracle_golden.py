"""CPU tests: the C oracle (oracle/) against the golden fixtures generated from the unmodified
reference (tests/golden/make_golden.py).  Bit-exact on integers, tolerance 0 on doubles (the
oracle is built without FMA contraction, like the reference)."""
import os
import re

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NODE_INTS = R.INT_FIELDS + ("rbs", "mot_ndx", "mot_len", "mot_spacer", "mot_spacendx")
NODE_FLOATS = R.FLOAT_FIELDS + ("gc_cont",)


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=True)


META = _load("meta_cases.npz")
SINGLE = _load("single_cases.npz")
DP = _load("dp_cases.npz")
MISC = _load("misc.npz")


def run_oracle_meta(seq, closed, mask):
    d, gc, unk = orc.encode(seq)
    masks = orc.find_masks(d, 50) if mask else None
    o = orc.make_opts(closed=closed, masks=masks)
    return orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, R.bins_blob(), o)


@pytest.mark.parametrize("name", list(META["names"]))
def test_meta_golden(name):
    seq = META[name + "/seq"].tobytes()
    closed, mask = (int(v) for v in META[name + "/opts"])
    genes, nodes, winner, pairs = run_oracle_meta(seq, closed, mask)
    assert winner == int(META[name + "/winner"])
    assert np.array_equal(genes, META[name + "/genes"])
    R.assert_nodes_equal(nodes, META[name + "/nodes"], ints=NODE_INTS, floats=NODE_FLOATS, what=name)


@pytest.mark.parametrize("name", [n for n in META["names"] if len(META[n + "/prodigal"])])
def test_meta_vs_prodigal_cli_headers(name):
    """gene boundaries vs the Prodigal CLI outputs the reference's own tests pin
    (tests/test_gene_finder.py:101-179)"""
    seq = META[name + "/seq"].tobytes()
    genes, nodes, winner, _ = run_oracle_meta(seq, 0, 0)
    hdr = list(META[name + "/prodigal"])
    assert len(hdr) == len(genes)
    for h, g in zip(hdr, genes):
        f = [x.strip() for x in h.split("#")]
        assert (int(f[1]), int(f[2])) == (int(g["begin"]), int(g["end"]))
        assert int(f[3]) == int(nodes[g["start_ndx"]]["strand"])


@pytest.mark.parametrize("name", list(SINGLE["names"]))
def test_single_golden(name):
    seq = SINGLE[name + "/seq"].tobytes()
    closed, mask = (int(v) for v in SINGLE[name + "/opts"])
    d, gc, unk = orc.encode(seq)
    genes, nodes, ipath = orc.find_genes_single(d, SINGLE[name + "/tinf"].tobytes(), orc.make_opts(closed=closed))
    assert np.array_equal(genes, SINGLE[name + "/genes"])
    R.assert_nodes_equal(nodes, SINGLE[name + "/nodes"],
                         ints=NODE_INTS + ("traceb", "tracef", "ov_mark", "star_ptr", "elim"),
                         floats=NODE_FLOATS + ("score",), what=name)
    hdr = list(SINGLE[name + "/prodigal"])
    if hdr:
        assert len(hdr) == len(genes)
        for h, g in zip(hdr, genes):
            f = [x.strip() for x in h.split("#")]
            assert (int(f[1]), int(f[2])) == (int(g["begin"]), int(g["end"]))


@pytest.mark.parametrize("name", list(DP["names"]))
@pytest.mark.parametrize("final", [True, False])
def test_dp_golden(name, final):
    arr = DP[name + "/in"].copy()
    b = int(DP[name + "/bin"])
    orc.score_connections(arr, R.bin_blob(b), final=final)
    tag = "final" if final else "train"
    assert np.array_equal(arr["traceb"], DP[f"{name}/{tag}/traceb"])
    assert np.array_equal(arr["ov_mark"], DP[f"{name}/{tag}/ov_mark"])
    assert np.array_equal(arr["score"], DP[f"{name}/{tag}/score"])


def test_node_counts_per_translation_table():
    """tests/test_nodes.py:28-39 (2970 for tt=4, 2293 for tt=11 on SRR492066) and every other table"""
    d, _, _ = orc.encode(MISC["srr_seq"].tobytes())
    counts = dict((int(a), int(b)) for a, b in MISC["srr_node_counts"])
    assert counts[4] == 2970 and counts[11] == 2293
    for tt, n in counts.items():
        assert len(orc.extract(d, tt)) == n, tt


def test_shine_dalgarno_known_answers():
    d, _, _ = orc.encode(MISC["srr_seq"].tobytes())
    w = np.frombuffer(R.bin_blob(20), dtype=np.float64, count=28, offset=80)
    for pos, start, strand, exact, want in MISC["srr_sd"]:
        # reverse strand coordinates are strand-relative in the reference API
        got = orc.shine_dalgarno(d, int(pos), int(start), w, int(strand), bool(exact))
        assert got == want, (pos, start, strand, exact)


def test_skip_filter_matches_six_clause_predicate():
    """impl/generic.h:29-36 evaluated literally vs the oracle's class table, all 4x4x3x3 classes"""
    arr = np.zeros(2, dtype=orc.NODE_DTYPE)
    for t1 in range(4):
        for s1 in (1, -1):
            for f1 in range(3):
                for t2 in range(4):
                    for s2 in (1, -1):
                        for f2 in range(3):
                            arr[0]["type"], arr[0]["strand"], arr[0]["ndx"] = t1, s1, 3 + f1
                            arr[1]["type"], arr[1]["strand"], arr[1]["ndx"] = t2, s2, 30 + f2
                            st1, st2 = t1 == 3, t2 == 3
                            want = ((not st1 and not st2 and s1 == s2) or (s1 == 1 and not st1 and s2 != 1)
                                    or (s1 != 1 and st1 and s2 == 1) or (s1 != 1 and not st1 and s2 == 1 and st2)
                                    or (s1 == s2 == 1 and not st1 and st2 and f1 != f2)
                                    or (s1 == s2 and s1 != 1 and st1 and not st2 and f1 != f2))
                            assert orc.skippable(arr, 0, 1) == int(want)
