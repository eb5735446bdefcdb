"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
(a) the golden fixtures generated from the unmodified reference and (b) the CPU oracle on seeded inputs.
Integers / indices bit-exact; doubles compared with tolerance 0 first and 1e-5 (the north star's bound)
as the hard limit -- see TOL below."""
import os

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5      # BASELINE.json north star: cscore/sscore within 1e-5
EXACT = 0.0     # what we actually expect: identical doubles (no FMA, same summation order)
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = np.load(os.path.join(G, "meta_cases.npz"), allow_pickle=True)
SINGLE = np.load(os.path.join(G, "single_cases.npz"), allow_pickle=True)
DP = np.load(os.path.join(G, "dp_cases.npz"), allow_pickle=True)


@pytest.fixture(scope="module")
def capi():
    from pyrodigal_b200 import _capi
    return _capi


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    c.set_models(R.bins_blob(), 50)
    yield c
    c.close()


def cmp_float(a, b, what, tol=EXACT):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.size == 0:
        return
    d = np.abs(a - b)
    if d.max() > tol:
        k = int(np.argmax(d))
        raise AssertionError(f"{what}: max |diff| {d.max():.3e} at {k}: gpu {a.flat[k]!r} ref {b.flat[k]!r} "
                             f"({int((d > tol).sum())} of {a.size} differ)")


def cmp_int(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)[:8]
        raise AssertionError(f"{what}: {len(np.argwhere(a != b))} mismatches, first at {bad.ravel().tolist()}: "
                             f"gpu {a[tuple(bad[0])]} ref {b[tuple(bad[0])]}")


def cmp_nodes(gpu, ref, what, dp=False, star=False):
    assert len(gpu) == len(ref), (what, len(gpu), len(ref))
    for f in ("ndx", "stop_val", "strand", "type", "edge", "rbs"):
        cmp_int(gpu[f], ref[f], f"{what}.{f}")
    for f, g in (("mot_ndx", "mot_ndx"), ("mot_len", "mot_len"), ("mot_spacer", "mot_spacer"), ("mot_spacendx", "mot_spacendx")):
        cmp_int(gpu[f], ref[g], f"{what}.{f}")
    for f in ("cscore", "uscore", "tscore", "rscore", "sscore", "gc_cont"):
        cmp_float(gpu[f], ref[f], f"{what}.{f}")
    if star or dp:
        cmp_int(gpu["star_ptr"], ref["star_ptr"], f"{what}.star_ptr")
    if dp:
        for f in ("traceb", "tracef", "ov_mark", "elim"):
            cmp_int(gpu[f], ref[f], f"{what}.{f}")
        cmp_float(gpu["score"], ref["score"], f"{what}.score")


# ---------------------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("length,gc,seed,nfrac", [(10000, .5, 1, 0), (50000, .35, 2, 0), (3001, .62, 3, 0),
                                                   (20000, .45, 4, .002), (100, .5, 5, 0), (5, .5, 6, 0), (3, .5, 7, 0)])
@pytest.mark.parametrize("tt", [11, 4, 1, 2, 22, 23, 25, 33])
@pytest.mark.parametrize("closed", [False, True])
def test_extract_nodes(ctx, capi, length, gc, seed, nfrac, tt, closed):
    seq = R.synth(length, gc, seed, n_frac=nfrac)
    d, _, _ = orc.encode(seq)
    want = orc.extract(d, tt, orc.make_opts(closed=closed))
    got = ctx.extract_nodes(np.frombuffer(seq, np.uint8), tt, capi.make_opts(closed=closed))
    for f in ("ndx", "stop_val", "strand", "type", "edge"):
        cmp_int(got[f], want[f], f"extract.{f}")


@pytest.mark.parametrize("mask_size", [50, 10])
def test_extract_nodes_masked(ctx, capi, mask_size):
    seq = R.synth(40000, .5, 11, n_frac=0.003)
    d, _, _ = orc.encode(seq)
    masks = orc.find_masks(d, mask_size)
    assert len(masks) > 0
    want = orc.extract(d, 11, orc.make_opts(masks=masks))
    got = ctx.extract_nodes(np.frombuffer(seq, np.uint8), 11, capi.make_opts(mask=True, min_mask=mask_size))
    for f in ("ndx", "stop_val", "strand", "type", "edge"):
        cmp_int(got[f], want[f], f"extract_masked.{f}")


@pytest.mark.parametrize("length,gc,seed", [(10000, .5, 21), (2500, .4, 22), (60000, .6, 23), (900, .55, 24)])
@pytest.mark.parametrize("model", [0, 2, 8, 11, 12, 24, 33, 49])
@pytest.mark.parametrize("is_meta", [True, False])
def test_score_nodes(ctx, capi, length, gc, seed, model, is_meta):
    seq = R.synth(length, gc, seed)
    d, _, _ = orc.encode(seq)
    blob = R.bin_blob(model)
    tt = int(np.frombuffer(blob, np.int32, count=1, offset=8)[0])
    ref = orc.extract(d, tt)
    for first_pass in (True, False):
        # first_pass=False == a second model scored on the same node array (edge flags carried)
        orc.reset_scores(ref)
        orc.score(d, ref, blob, closed=False, is_meta=is_meta)
        orc.record_overlapping_starts(ref, blob, flag=1, max_overlap=60)
        got = ctx.score_nodes(np.frombuffer(seq, np.uint8), model, capi.make_opts(), is_meta=is_meta, first_pass=first_pass)
        cmp_nodes(got, ref, f"score_nodes[m{model} L{length} fp{first_pass}]", star=True)


@pytest.mark.parametrize("name", list(DP["names"]))
@pytest.mark.parametrize("final", [True, False])
def test_score_connections_golden(ctx, name, final):
    """the reference's tests/test_connection_scorer.py protocol: same node arrays in, score/traceb/ov_mark out"""
    a = DP[name + "/in"]
    b = int(DP[name + "/bin"])
    score, traceb, ov, pairs, ms = ctx.score_connections(
        a["ndx"], a["stop_val"], a["strand"], a["type"], a["cscore"], a["sscore"], a["rscore"], a["uscore"],
        a["gc_score"], a["star_ptr"], b, final)
    tag = "final" if final else "train"
    cmp_int(traceb, DP[f"{name}/{tag}/traceb"], f"{name}.traceb")
    cmp_int(ov, DP[f"{name}/{tag}/ov_mark"], f"{name}.ov_mark")
    cmp_float(score, DP[f"{name}/{tag}/score"], f"{name}.score")
    ref = a.copy()
    assert pairs == orc.score_connections(ref, R.bin_blob(b), final=final)


@pytest.mark.parametrize("length,gc,tt,seed", [(30000, .35, 11, 71), (45000, .5, 11, 72), (60000, .65, 11, 73), (40000, .3, 4, 74),
                                               (25000, .72, 11, 75)])
@pytest.mark.parametrize("scored", [False, True])
def test_training_dp_vs_oracle(ctx, length, gc, tt, seed, scored):
    """ConnectionScorer.score_connections(final=False) -- the training DP, k_dp_dq<.., 0> -- on node arrays the oracle
    extracted from synthetic sequences: with the state the training pass really has (scores reset, first start of every
    frame recorded) and with scored nodes / best recorded starts (the comparison of the bias sum with the final DP's value
    in the triple-overlap search, _connection.h:320-324, then sees non-zero values)"""
    rng = np.random.default_rng(seed)
    d, _, _ = orc.encode(R.synth(length, gc, seed))
    model = 0 if tt == 4 else 11
    blob = R.bin_blob(model)
    nodes = orc.extract(d, tt)
    if scored:
        orc.score(d, nodes, blob)
    else:
        orc.reset_scores(nodes)
    orc.record_overlapping_starts(nodes, blob, flag=1 if scored else 0)
    nodes["gc_score"] = rng.integers(0, 60, size=(len(nodes), 3)).astype(np.float64) * rng.choice([1.0, 0.5, 0.0], size=(len(nodes), 1))
    ref = nodes.copy()
    ref["ov_mark"] = -1   # (the DP only writes the marker of nodes something leads into)
    pairs = orc.score_connections(ref, blob, final=False)
    score, traceb, ov, got_pairs, ms = ctx.score_connections(
        nodes["ndx"], nodes["stop_val"], nodes["strand"], nodes["type"], nodes["cscore"], nodes["sscore"], nodes["rscore"],
        nodes["uscore"], nodes["gc_score"], nodes["star_ptr"], model, False)
    cmp_int(traceb, ref["traceb"], "train_dp.traceb")
    cmp_int(ov, ref["ov_mark"], "train_dp.ov_mark")
    cmp_float(score, ref["score"], "train_dp.score")
    assert got_pairs == pairs


@pytest.mark.parametrize("name", list(DP["names"]))
def test_compute_skippable(ctx, name):
    """the skip filter as an operator (ConnectionScorer.compute_skippable, lib.pyx:1321-1334): every fixture, windows at
    both ends and in the middle of the node array, the reference's 500-node window and shorter / longer ones"""
    a = DP[name + "/in"]
    n = len(a)
    checked = 0
    for i in sorted({0, 1, 2, 17, min(n - 1, 499), min(n - 1, 500), min(n - 1, 501), n // 3, n // 2, n - 2, n - 1}):
        if i < 0 or i >= n:
            continue
        for width in (1, 7, 500, 1000, 2500):
            mn = max(0, i - width)
            skip = ctx.compute_skippable(a["strand"], a["type"], a["ndx"], mn, i)
            want = np.array([orc.skippable(a, j, i) for j in range(mn, i)], dtype=np.uint8)
            cmp_int(skip[mn:i], want, f"skippable[{name} i={i} min={mn}]")
            checked += i - mn
    assert checked > 0


def test_skippable_plugin_signature(ctx, capi):
    """pgpu_skippable(strands, types, frames, min, i, skip) -- the reference's skippable_t (lib.pxd:120) -- gives what the
    operator gives, writes [min, i) only"""
    a = DP[list(DP["names"])[0] + "/in"]
    n = len(a)
    strands = a["strand"].astype(np.int8).view(np.uint8)     # 1 / 255, as BaseConnectionScorer._index stores them
    frames = (a["ndx"] % 3).astype(np.uint8)
    for i, mn in ((n - 1, max(0, n - 1 - 500)), (n // 2, max(0, n // 2 - 1000)), (5, 0), (1, 0)):
        got = capi.skippable_plugin(strands, a["type"], frames, mn, i)
        want = ctx.compute_skippable(a["strand"], a["type"], a["ndx"], mn, i)
        cmp_int(got[mn:i], want[mn:i], f"skippable_plugin[i={i}]")
        assert (got[:mn] == 0xAA).all() and (got[i:] == 0xAA).all()


def test_result_nodes_in_reference_struct_layout(ctx, capi):
    """pgpu_result_nodes_struct: the final nodes as 128-byte `struct _node` records (node.h:41-76) == the pgpu_node records"""
    seqs = [R.synth(9000, .45, 31), R.synth(4000, .62, 32)]
    res = run_meta(ctx, capi, seqs)
    for k in range(len(seqs)):
        a, b = res.nodes(k), res.nodes_struct(k)
        assert len(a) == len(b) > 0 and b.dtype.itemsize == 128
        for f in ("ndx", "stop_val", "traceb", "tracef", "star_ptr", "strand", "type", "edge", "elim", "ov_mark", "rbs",
                  "gc_cont", "mot_score", "cscore", "uscore", "tscore", "rscore", "sscore", "score"):
            assert np.array_equal(a[f], b[f]), f
        bits = b["mot_bits"]
        assert np.array_equal(bits & 0xfff, a["mot_ndx"]) and np.array_equal((bits >> 12) & 0xf, a["mot_spacer"])
        assert np.array_equal((bits >> 16) & 0x7, a["mot_len"]) and np.array_equal((bits >> 19) & 0x3, a["mot_spacendx"])
        assert not b["gc_score"].any() and not b["gc_bias"].any()
    res.free()


# ---------------------------------------------------------------------------------------------------
# find_genes
# ---------------------------------------------------------------------------------------------------
def run_meta(ctx, capi, seqs, closed=False, mask=False, want_nodes=True):
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    flat = np.ascontiguousarray(np.concatenate(arrs)) if arrs else np.zeros(0, np.uint8)
    return ctx.find_genes_batch(flat, off, capi.make_opts(meta=True, closed=closed, mask=mask, want_nodes=want_nodes))


@pytest.mark.parametrize("name", list(META["names"]))
def test_find_genes_meta_golden(ctx, capi, name):
    seq = META[name + "/seq"].tobytes()
    closed, mask = (int(v) for v in META[name + "/opts"])
    res = run_meta(ctx, capi, [seq], closed=closed, mask=mask)
    s = res.summary[0]
    assert int(s["winner"]) == int(META[name + "/winner"]), name
    cmp_int(res.genes, META[name + "/genes"], f"{name}.genes")
    cmp_nodes(res.nodes(0), META[name + "/nodes"], name)
    # start/stop node records copied with the genes == the node array entries
    nodes = res.nodes(0)
    for k, g in enumerate(res.genes):
        assert res.gene_nodes[k, 0].tobytes() == nodes[g["start_ndx"]].tobytes()
        assert res.gene_nodes[k, 1].tobytes() == nodes[g["stop_ndx"]].tobytes()


@pytest.mark.parametrize("name", list(SINGLE["names"]))
def test_find_genes_single_golden(capi, name):
    seq = SINGLE[name + "/seq"]
    closed, mask = (int(v) for v in SINGLE[name + "/opts"])
    c = capi.Context(0)
    c.set_models(SINGLE[name + "/tinf"].tobytes(), 1)
    res = c.find_genes_batch(np.ascontiguousarray(seq), np.array([0, len(seq)], np.int64),
                             capi.make_opts(meta=False, single_model=0, closed=closed, want_nodes=True))
    cmp_int(res.genes, SINGLE[name + "/genes"], f"{name}.genes")
    cmp_nodes(res.nodes(0), SINGLE[name + "/nodes"], name, dp=True)
    c.close()


def test_find_genes_batch_vs_oracle(ctx, capi):
    """one batched call over ragged contigs (empty, tiny, short-meta, long) == oracle contig by contig"""
    rng = np.random.default_rng(99)
    seqs = [b"", b"ACG", R.synth(89, .5, 1), R.synth(120000, .52, 2)]
    for k in range(60):
        seqs.append(R.synth(int(rng.integers(100, 30000)), float(rng.uniform(.28, .72)), 1000 + k,
                            n_frac=0.001 if k % 7 == 0 else 0.0))
    res = run_meta(ctx, capi, seqs)
    total_pairs = 0
    for k, s in enumerate(seqs):
        d, gc, unk = orc.encode(s)
        genes, nodes, winner, pairs = orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, R.bins_blob())
        total_pairs += pairs
        assert int(res.summary["winner"][k]) == winner, k
        assert int(res.summary["gc_count"][k]) == gc and int(res.summary["unknown"][k]) == unk
        a, b = res.gene_off[k], res.gene_off[k + 1]
        cmp_int(res.genes[a:b], genes, f"contig{k}.genes")
        if winner >= 0:
            cmp_nodes(res.nodes(k), nodes, f"contig{k}")
    assert res.stats["pairs"] == total_pairs
    assert res.stats["kernel_launches"] > 0


@pytest.mark.parametrize("closed,mask", [(True, False), (False, True)])
def test_find_genes_options_vs_oracle(ctx, capi, closed, mask):
    seqs = [R.synth(int(L), gc, 500 + i, n_frac=0.002 if mask else 0.0)
            for i, (L, gc) in enumerate([(8000, .4), (2500, .55), (40000, .65), (1400, .5)])]
    res = run_meta(ctx, capi, seqs, closed=closed, mask=mask)
    for k, s in enumerate(seqs):
        d, gc, unk = orc.encode(s)
        masks = orc.find_masks(d, 50) if mask else None
        genes, nodes, winner, _ = orc.find_genes_meta(d, gc / len(d), R.bins_blob(), orc.make_opts(closed=closed, masks=masks))
        assert int(res.summary["winner"][k]) == winner
        a, b = res.gene_off[k], res.gene_off[k + 1]
        cmp_int(res.genes[a:b], genes, f"contig{k}.genes")
        if winner >= 0:
            cmp_nodes(res.nodes(k), nodes, f"contig{k}")


@pytest.mark.parametrize("algo", [3, 6])
def test_dp_kernel_variants_single_golden(capi, algo, monkeypatch):
    """every DP field of the single-mode goldens, through the per-chain deque kernel (3) and through the
    model-lane kernel forced onto a one-chain group (6)"""
    monkeypatch.setenv("PGPU_DP_ALGO", str(algo))
    for name in list(SINGLE["names"]):
        seq = SINGLE[name + "/seq"]
        closed, mask = (int(v) for v in SINGLE[name + "/opts"])
        c = capi.Context(0)
        c.set_models(SINGLE[name + "/tinf"].tobytes(), 1)
        res = c.find_genes_batch(np.ascontiguousarray(seq), np.array([0, len(seq)], np.int64),
                                 capi.make_opts(meta=False, single_model=0, closed=closed, want_nodes=True))
        cmp_int(res.genes, SINGLE[name + "/genes"], f"{name}.genes")
        cmp_nodes(res.nodes(0), SINGLE[name + "/nodes"], name, dp=True)
        c.close()


def test_dp_model_lane_kernel_equals_per_chain_kernel(capi, monkeypatch):
    """PGPU_DP_VERIFY: in one meta batch, every score / traceback / overlap frame of every chain from the
    model-lane kernel (one warp per contig x table, one lane per model) equals the per-chain kernel's"""
    monkeypatch.setenv("PGPU_DP_VERIFY", "1")
    c = capi.Context(0)
    c.set_models(R.bins_blob(), 50)
    rng = np.random.default_rng(7)
    seqs = [b"", R.synth(89, .5, 1), R.synth(150000, .55, 3), R.synth(90000, .33, 4)]
    for k in range(80):
        seqs.append(R.synth(int(rng.integers(100, 40000)), float(rng.uniform(.26, .74)), 3000 + k,
                            n_frac=0.001 if k % 9 == 0 else 0.0))
    res = run_meta(c, capi, seqs)
    assert res.stats["n_chains"] > 5 * len(seqs)
    # same genes as the default context's path
    monkeypatch.delenv("PGPU_DP_VERIFY")
    monkeypatch.setenv("PGPU_DP_ALGO", "3")
    c3 = capi.Context(0)
    c3.set_models(R.bins_blob(), 50)
    res3 = run_meta(c3, capi, seqs)
    cmp_int(res.genes, res3.genes, "ml-vs-dq.genes")
    assert res.gene_nodes.tobytes() == res3.gene_nodes.tobytes()
    c.close(); c3.close()


def test_resident_batch_matches_host_batch(ctx, capi):
    seqs = [R.synth(5000 + 700 * k, .4 + 0.02 * k, 700 + k) for k in range(12)]
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    flat = np.ascontiguousarray(np.concatenate(arrs))
    o = capi.make_opts(meta=True)
    r1 = ctx.find_genes_batch(flat, off, o)
    b = ctx.upload(flat, off)
    r2 = b.run(o)
    r3 = b.run(o)
    for r in (r2, r3):
        cmp_int(r.genes, r1.genes, "resident.genes")
        assert r.gene_nodes.tobytes() == r1.gene_nodes.tobytes()
    b.free()


def test_python_api_drop_in():
    """GeneFinder / Genes / Gene surface vs the attribute table recorded from the reference"""
    import pyrodigal_b200
    for name in ("cfg1_10k", "KK037166", "s20k_N_mask"):
        seq = META[name + "/seq"].tobytes()
        closed, mask = (bool(v) for v in META[name + "/opts"])
        gf = pyrodigal_b200.GeneFinder(meta=True, closed=closed, mask=mask)
        genes = gf.find_genes(seq)
        table = META[name + "/gene_table"]
        assert len(genes) == len(table)
        for g, row in zip(genes, table):
            assert (g.begin, g.end, g.strand, int(g.partial_begin), int(g.partial_end), g.start_type) == tuple(row[:6])
            assert str(g.rbs_motif) == row[6] and str(g.rbs_spacer) == row[7]
            for got, want in zip((g.gc_cont, g.cscore, g.rscore, g.sscore, g.tscore, g.uscore, g.score, g.confidence()), row[8:]):
                assert abs(got - float(want)) <= TOL
        w = int(META[name + "/winner"])
        assert (genes.metagenomic_bin is pyrodigal_b200.METAGENOMIC_BINS[w]) if w >= 0 else genes.metagenomic_bin is None
        assert len(genes.nodes) == len(META[name + "/nodes"])
        hdr = list(META[name + "/prodigal"])
        if hdr and not (closed or mask):
            for g, h in zip(genes, hdr):
                f = [x.strip() for x in h.split("#")]
                assert (int(f[1]), int(f[2]), int(f[3])) == (g.begin, g.end, g.strand)
                assert g._gene_data(f[4].split(";")[0].split("=")[1].rsplit("_", 1)[0]) == f[4]
    # single mode through the Python surface
    name = "srr_trained"
    ti = pyrodigal_b200.TrainingInfo._from_bytes(SINGLE[name + "/tinf"].tobytes())
    genes = pyrodigal_b200.GeneFinder(ti).find_genes(SINGLE[name + "/seq"].tobytes())
    table = SINGLE[name + "/gene_table"]
    assert [(g.begin, g.end, g.strand) for g in genes] == [tuple(r[:3]) for r in table]
    with pytest.raises(RuntimeError):
        pyrodigal_b200.GeneFinder().find_genes(b"ACGT")
    with pytest.raises(ValueError):
        pyrodigal_b200.GeneFinder(ti, meta=True)


def test_long_contig_meta_and_single_vs_oracle(ctx, capi):
    """one long chain (1.2 Mbp): windows slide over > 1000 nodes, many ORFs, both modes; plus a tt=4 model on a
    GC-rich contig, which produces the giant-ORF windows of SURVEY T2"""
    seq = R.synth(1_200_000, 0.5, 4242)
    d, gc, unk = orc.encode(seq)
    res = run_meta(ctx, capi, [seq])
    genes, nodes, winner, pairs = orc.find_genes_meta(d, gc / len(d), R.bins_blob())
    assert int(res.summary["winner"][0]) == winner
    cmp_int(res.genes, genes, "long.genes")
    cmp_nodes(res.nodes(0), nodes, "long")
    assert res.stats["pairs"] == pairs
    for b in (0, 30):  # bin 0: translation table 4 (no TGA stop) -> ORFs of 10^4..10^5 nodes at GC 0.5
        c = capi.Context(0)
        c.set_models(R.bin_blob(b), 1)
        sub = np.frombuffer(seq, np.uint8)[:600_000]
        r = c.find_genes_batch(np.ascontiguousarray(sub), np.array([0, len(sub)], np.int64),
                               capi.make_opts(meta=False, single_model=0, want_nodes=True))
        g2, n2, ipath = orc.find_genes_single(d[:600_000], R.bin_blob(b))
        cmp_int(r.genes, g2, f"single{b}.genes")
        cmp_nodes(r.nodes(0), n2, f"single{b}", dp=True)
        assert int(r.summary["ipath"][0]) == ipath
        c.close()


def test_sub_batching_matches_single_batch(capi):
    """a tiny workspace limit forces run_all to split the contigs into several sub-batches; genes, their node
    records and the per-contig node arrays must be identical to the single-batch run"""
    seqs = [R.synth(60_000 + 1000 * (k % 7), 0.35 + 0.005 * k, 9000 + k) for k in range(64)]
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    flat = np.ascontiguousarray(np.concatenate(arrs))
    o = capi.make_opts(meta=True, want_nodes=True)
    c1 = capi.Context(0)
    c1.set_models(R.bins_blob(), 50)
    r1 = c1.find_genes_batch(flat, off, o)
    c2 = capi.Context(0)
    c2.set_models(R.bins_blob(), 50)
    capi.check(capi.lib.pgpu_set_workspace_limit(c2.handle, 1), c2.handle)  # -> 1 Mbp per sub-batch
    r2 = c2.find_genes_batch(flat, off, o)
    assert r2.stats["kernel_launches"] > 2 * r1.stats["kernel_launches"]
    assert r1.summary.tobytes() == r2.summary.tobytes()
    cmp_int(r2.genes, r1.genes, "subbatch.genes")
    assert r2.gene_nodes.tobytes() == r1.gene_nodes.tobytes()
    for k in (0, 17, 63):
        assert r1.nodes(k).tobytes() == r2.nodes(k).tobytes()
    for k in range(len(seqs)):
        a, b = r2.gene_off[k], r2.gene_off[k + 1]
        g = np.zeros(b - a, dtype=capi.GENE_DTYPE)
        capi.check(capi.lib.pgpu_result_genes(r2.handle, k, capi.ptr(g)))
        cmp_int(g, r1.genes[a:b], f"subbatch.contig{k}")
    r1.free(); r2.free(); c1.close(); c2.close()


def test_single_chain_16mbp_vs_oracle(capi):
    """cfg5-style stress (one long chromosome, single mode, one DP chain of ~0.9 M nodes): gene boundaries and the
    whole DP state must match the oracle; also exercises int32/int64 offsets well beyond the batch tests"""
    seq = R.synth(16_000_000, 0.5, 55)
    d, gc, unk = orc.encode(seq)
    blob = R.bin_blob(20)
    g2, n2, ipath = orc.find_genes_single(d, blob)
    c = capi.Context(0)
    c.set_models(blob, 1)
    a = np.frombuffer(seq, np.uint8)
    r = c.find_genes_batch(np.ascontiguousarray(a), np.array([0, len(a)], np.int64),
                           capi.make_opts(meta=False, single_model=0, want_nodes=True))
    cmp_int(r.genes, g2, "chr.genes")
    cmp_nodes(r.nodes(0), n2, "chr", dp=True)
    assert int(r.summary["ipath"][0]) == ipath and len(n2) > 500_000
    r.free(); c.close()


@pytest.mark.parametrize("closed", [False, True])
def test_gene_only_final_pass_equals_full_final_pass(ctx, capi, closed):
    """meta mode without node arrays re-scores only the genes' ORFs: the start / stop node records of every gene must
    be byte-identical to the ones taken from the full final pass (want_nodes=True), which is checked against the
    oracle node by node elsewhere"""
    rng = np.random.default_rng(5)
    seqs = [b"", R.synth(95, .5, 3), R.synth(150000, .45, 4)]
    for k in range(80):
        seqs.append(R.synth(int(rng.integers(200, 40000)), float(rng.uniform(.28, .72)), 7000 + k,
                            n_frac=0.002 if k % 5 == 0 else 0.0))
    full = run_meta(ctx, capi, seqs, closed=closed, want_nodes=True)
    lean = run_meta(ctx, capi, seqs, closed=closed, want_nodes=False)
    assert np.array_equal(full.gene_off, lean.gene_off)
    assert len(lean.genes) > 500
    assert lean.genes.tobytes() == full.genes.tobytes()
    assert lean.gene_nodes.tobytes() == full.gene_nodes.tobytes()
    # and against the full node arrays directly
    gn = lean.gene_nodes.reshape(-1, 2)
    for k in range(len(seqs)):
        a, b = full.gene_off[k], full.gene_off[k + 1]
        if b > a:
            nodes = full.nodes(k)
            # field-wise: fancy indexing leaves the padding bytes of a structured copy uninitialised
            assert np.array_equal(nodes[full.genes["start_ndx"][a:b]], gn[a:b, 0])
            assert np.array_equal(nodes[full.genes["stop_ndx"][a:b]], gn[a:b, 1])


@pytest.mark.parametrize("closed", [False, True])
def test_gene_only_final_pass_vs_oracle(ctx, capi, closed):
    """the path bench.py times (meta mode, want_nodes=False): genes, winner and the start / stop node record of every
    gene -- all five scores, rbs, motif, gc_cont -- directly against the oracle's node arrays"""
    rng = np.random.default_rng(15)
    seqs = [b"", R.synth(95, .5, 3), R.synth(90000, .47, 14)]
    for k in range(40):
        seqs.append(R.synth(int(rng.integers(200, 30000)), float(rng.uniform(.28, .72)), 17000 + k,
                            n_frac=0.002 if k % 5 == 0 else 0.0))
    lean = run_meta(ctx, capi, seqs, closed=closed, want_nodes=False)
    gn = lean.gene_nodes.reshape(-1, 2)
    checked = 0
    for k, s in enumerate(seqs):
        d, gc, unk = orc.encode(s)
        genes, nodes, winner, _ = orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, R.bins_blob(),
                                                      orc.make_opts(closed=closed))
        assert int(lean.summary["winner"][k]) == winner, k
        a, b = lean.gene_off[k], lean.gene_off[k + 1]
        cmp_int(lean.genes[a:b], genes, f"contig{k}.genes")
        if b > a:
            cmp_nodes(gn[a:b, 0], nodes[genes["start_ndx"]], f"contig{k}.start_nodes")
            cmp_nodes(gn[a:b, 1], nodes[genes["stop_ndx"]], f"contig{k}.stop_nodes")
            checked += b - a
    assert checked > 300


@pytest.mark.parametrize("many_parts", [False, True])
def test_two_lane_host_batches_match_single_stream(capi, monkeypatch, many_parts):
    """host-input batches above PGPU_LANE_MIN_BP run as sub-batches on two worker threads / streams; the stitched
    result (summaries, genes, gene node records, node arrays, counters) must be identical to the single-stream run"""
    seqs = [R.synth(20_000 + 3000 * (k % 11), 0.32 + 0.004 * k, 12000 + k) for k in range(90)] + [b"", R.synth(50, .5, 1)]
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    flat = np.ascontiguousarray(np.concatenate(arrs))
    o = capi.make_opts(meta=True, want_nodes=True)
    monkeypatch.setenv("PGPU_LANES", "1")
    c1 = capi.Context(0)
    c1.set_models(R.bins_blob(), 50)
    r1 = c1.find_genes_batch(flat, off, o)
    monkeypatch.setenv("PGPU_LANES", "2")
    monkeypatch.setenv("PGPU_LANE_MIN_BP", "0")
    c2 = capi.Context(0)
    c2.set_models(R.bins_blob(), 50)
    if many_parts:
        capi.check(capi.lib.pgpu_set_workspace_limit(c2.handle, 1), c2.handle)  # -> 1 Mbp per sub-batch, several per lane
    for rep in range(2):  # the second call reuses the lane streams and the recycled buffers
        r2 = c2.find_genes_batch(flat, off, o)
        assert r1.summary.tobytes() == r2.summary.tobytes()
        assert np.array_equal(r1.gene_off, r2.gene_off)
        cmp_int(r2.genes, r1.genes, "lanes.genes")
        assert r2.gene_nodes.tobytes() == r1.gene_nodes.tobytes()
        for k in (0, 1, 44, 45, 46, 89):
            assert r1.nodes(k).tobytes() == r2.nodes(k).tobytes()
        for key in ("n_contigs", "total_bp", "total_nodes", "total_chain_nodes", "n_chains", "total_genes", "pairs", "dp_steps"):
            assert r1.stats[key] == r2.stats[key], key
        for k in range(len(seqs)):
            a, b = r2.gene_off[k], r2.gene_off[k + 1]
            g = np.zeros(b - a, dtype=capi.GENE_DTYPE)
            capi.check(capi.lib.pgpu_result_genes(r2.handle, k, capi.ptr(g)))
            cmp_int(g, r1.genes[a:b], f"lanes.contig{k}")
        r2.free()
    # the genes-only path (no node arrays) through the lanes as well
    o2 = capi.make_opts(meta=True, want_nodes=False)
    ra, rb = c1.find_genes_batch(flat, off, o2), c2.find_genes_batch(flat, off, o2)
    assert ra.gene_nodes.tobytes() == rb.gene_nodes.tobytes() and ra.genes.tobytes() == rb.genes.tobytes()
    ra.free(); rb.free(); r1.free(); c1.close(); c2.close()


def test_coding_score_lane_groups(capi):
    """k_coding_orf / k_overlap_lanes with 4 / 8 / 16 / 32 lanes per ORF (one lane per model of the extraction), on contigs
    whose GC content sweeps the whole range, so that extractions with 1 ... 27 models (both translation tables) occur"""
    c = capi.Context(0)
    c.set_models(R.bins_blob(), 50)
    seqs = [R.synth(2500 + 137 * k, 0.24 + 0.02 * k, 4100 + k) for k in range(27)]
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    flat = np.ascontiguousarray(np.concatenate(arrs))
    res = c.find_genes_batch(flat, off, capi.make_opts(meta=True, want_nodes=True))
    for k, s in enumerate(seqs):
        d, gc, unk = orc.encode(s)
        genes, nodes, winner, pairs = orc.find_genes_meta(d, gc / len(d), R.bins_blob())
        a, b = res.gene_off[k], res.gene_off[k + 1]
        assert int(res.summary["winner"][k]) == winner and b - a == len(genes), k
        if winner >= 0:
            n = res.nodes(k)
            for f in ("ndx", "cscore", "sscore", "score", "traceb"):
                assert np.array_equal(n[f], nodes[f]), (k, f)
    assert res.stats["n_chains"] > 27 * 5
    c.close()


def test_coding_score_shared_memory_tables(capi, monkeypatch):
    """k_coding_flat (dicodon tables of four neighbouring models in shared memory, ORF slots drawn from a counter) against
    k_coding_orf (PGPU_CODING_VERIFY: every raw coding score of every chain, bit for bit) on a GC sweep -- extractions with
    1 ... 27 models, i.e. every table set and every group width -- then against the oracle; PGPU_CODING_SMEM=0 gives the
    same genes"""
    seqs = [R.synth(2000 + 613 * (k % 5), 0.24 + 0.02 * (k % 27), 6300 + k) for k in range(54)]
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    flat = np.ascontiguousarray(np.concatenate(arrs))
    out = {}
    for mode, verify in (("1", "0"), ("1", "1"), ("0", "1")):
        monkeypatch.setenv("PGPU_CODING_SMEM", mode)
        monkeypatch.setenv("PGPU_CODING_VERIFY", verify)
        c = capi.Context(0)
        c.set_models(R.bins_blob(), 50)
        res = c.find_genes_batch(flat, off, capi.make_opts(meta=True, want_nodes=False))
        out[mode, verify] = (res.genes.tobytes(), res.gene_nodes.tobytes(), res.summary.tobytes(), res.stats["kernel_launches"])
        if (mode, verify) == ("1", "1"):
            for k in (0, 5, 11, 20, 26, 33, 47):
                d, gc, unk = orc.encode(seqs[k])
                genes, nodes, winner, pairs = orc.find_genes_meta(d, gc / len(d), R.bins_blob())
                a, b = res.gene_off[k], res.gene_off[k + 1]
                assert int(res.summary["winner"][k]) == winner and b - a == len(genes), k
                cmp_int(res.genes[a:b]["begin"], genes["begin"], f"smem.contig{k}.begin")
                cmp_int(res.genes[a:b]["end"], genes["end"], f"smem.contig{k}.end")
        res.free(); c.close()
    assert out["1", "1"][3] == out["1", "0"][3] + 2, "the self-check did not run: k_coding_flat was not selected"
    assert out["0", "1"][3] == out["1", "0"][3] - 3, "PGPU_CODING_SMEM=0: one coding kernel instead of four, no self-check"
    for key in (("1", "1"), ("0", "1")):
        assert out[key][:3] == out["1", "0"][:3], key


def test_coding_score_many_plan_entries(capi, monkeypatch):
    """k_cq_plan lays out the ORF slots of more than 1024 plan entries (several rounds of its block-wide scan) and of
    classes that need padding: 700 short contigs over the whole GC range, raw coding scores checked against k_coding_orf
    (PGPU_CODING_VERIFY), genes of a few contigs against the oracle"""
    monkeypatch.setenv("PGPU_CODING_SMEM", "1")
    monkeypatch.setenv("PGPU_CODING_VERIFY", "1")
    seqs = [R.synth(400 + 37 * (k % 23), 0.26 + 0.0006 * k, 9100 + k) for k in range(700)]
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    flat = np.ascontiguousarray(np.concatenate(arrs))
    c = capi.Context(0)
    c.set_models(R.bins_blob(), 50)
    res = c.find_genes_batch(flat, off, capi.make_opts(meta=True, want_nodes=False))
    assert res.stats["n_chains"] > 4 * 1024
    for k in (0, 99, 350, 699):
        d, gc, unk = orc.encode(seqs[k])
        genes, nodes, winner, pairs = orc.find_genes_meta(d, gc / len(d), R.bins_blob())
        a, b = res.gene_off[k], res.gene_off[k + 1]
        assert int(res.summary["winner"][k]) == winner and b - a == len(genes), k
        cmp_int(res.genes[a:b]["begin"], genes["begin"], f"plan.contig{k}.begin")
    res.free(); c.close()


def test_dp_model_lane_kernel_gc_sweep(capi, monkeypatch):
    """k_dp_ml on extractions with 1 ... 27 chains (lanes), both register budgets: checked by the library's own
    PGPU_DP_VERIFY comparison with k_dp_dq on a GC sweep (1 ... 27 models per extraction) and against the oracle"""
    monkeypatch.setenv("PGPU_DP_ML_MINB", "8")   # the 64-register instantiation; the default one runs everywhere else
    monkeypatch.setenv("PGPU_DP_VERIFY", "1")
    c = capi.Context(0)
    c.set_models(R.bins_blob(), 50)
    seqs = [R.synth(3000 + 911 * (k % 7), 0.24 + 0.02 * (k % 27), 5200 + k) for k in range(54)]
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    flat = np.ascontiguousarray(np.concatenate(arrs))
    res = c.find_genes_batch(flat, off, capi.make_opts(meta=True, want_nodes=True))
    for k in (0, 3, 13, 26, 27, 40, 53):
        d, gc, unk = orc.encode(seqs[k])
        genes, nodes, winner, pairs = orc.find_genes_meta(d, gc / len(d), R.bins_blob())
        a, b = res.gene_off[k], res.gene_off[k + 1]
        assert int(res.summary["winner"][k]) == winner and b - a == len(genes), k
        if winner >= 0:
            n = res.nodes(k)
            for f in ("score", "traceb", "ov_mark", "cscore"):
                assert np.array_equal(n[f], nodes[f]), (k, f)
    c.close()
