"""GPU parity tests of the training path (run with -m gpu on a B200): GeneFinder.train through the C ABI
(pgpu_train) against (a) tests/golden/train_cases.npz, generated from the unmodified reference -- including the
reference's own training golden (tests/test_training_info.py:60-66) and the published scalars of
tests/test_gene_finder.py:329-345 -- and (b) the CPU oracle on seeded inputs.  The training struct must be
byte-identical (tolerance 0)."""
import os
import warnings

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRAIN = np.load(os.path.join(G, "train_cases.npz"), allow_pickle=True)


def diff_fields(a, b):
    x, y = (np.frombuffer(v, dtype=orc.TRAINING_DTYPE)[0] for v in (a, b))
    out = []
    for f in orc.TRAINING_DTYPE.names:
        if not np.array_equal(x[f], y[f], equal_nan=True):
            d = np.abs(np.asarray(x[f], np.float64) - np.asarray(y[f], np.float64))
            out.append((f, int((d > 0).sum()), float(np.nanmax(d))))
    return out


def diagnose(prefix, seq, closed, mask, force, tt, st_wt):
    """which intermediate array of the training pass differs from the oracle first (PGPU_TRAIN_DUMP files)"""
    d, gc, unk = orc.encode(seq)
    masks = orc.find_masks(d, 50) if mask else None
    blob, nodes = orc.train(d, gc / len(d), translation_table=tt, start_weight=st_wt, force_nonsd=force,
                            opts=orc.make_opts(closed=closed, masks=masks), return_nodes=True)
    rep = []

    def load(name, dt):
        p = f"{prefix}.{name}.bin"
        return np.fromfile(p, dtype=dt) if os.path.exists(p) else None

    def cmp(name, got, want):
        if got is None:
            rep.append(f"{name}: not dumped")
            return
        want = np.asarray(want).reshape(-1)
        if got.shape != want.shape:
            rep.append(f"{name}: shape {got.shape} vs {want.shape}")
            return
        bad = np.flatnonzero(~((got == want) | (np.isnan(got.astype(np.float64)) & np.isnan(want.astype(np.float64)))))
        rep.append(f"{name}: {len(bad)} of {len(want)} differ" + (f", first at {bad[0]}: gpu {got[bad[0]]} oracle {want[bad[0]]}" if len(bad) else ""))

    ndx = load("ndx", np.int32)
    cmp("ndx", ndx, nodes["ndx"])
    cmp("gp", load("gp", np.int8), orc.gc_frame_plot(d))
    starts = nodes["type"] != 3
    gs = load("gc_score", np.float64)
    if gs is not None and len(gs) == 3 * len(nodes):
        cmp("gc_score(starts)", gs.reshape(-1, 3)[starts].reshape(-1), nodes["gc_score"][starts])
    gb = load("gc_bias", np.int8)
    if gb is not None and len(gb) == len(nodes):
        cmp("gc_bias(starts)", gb[starts], nodes["gc_bias"][starts].astype(np.int8))
    cmp("star_ptr", load("star_ptr", np.int32), nodes["star_ptr"])
    cmp("dp score", load("score", np.float64), nodes["score"])
    cmp("traceb (untangled)", load("traceb", np.int32), nodes["traceb"])
    cs = load("cscore", np.float64)
    if cs is not None and len(cs) == len(nodes):
        cmp("cscore(starts)", cs[starts], nodes["cscore"][starts])
    rb = load("rbs", np.uint8)
    if rb is not None and len(rb) == 2 * len(nodes):
        ne = starts & (nodes["edge"] == 0)
        cmp("rbs(non-edge starts)", rb.reshape(-1, 2)[ne].reshape(-1), nodes["rbs"][ne].astype(np.uint8))
    for n in ("ipath", "n_intervals"):
        v = load(n, np.int32)
        rep.append(f"{n}: {v}")
    return "\n".join(rep)


def gpu_train(seq, closed=False, mask=False, force=False, tt=11, st_wt=4.35, tmp=None, expect=None):
    import pyrodigal_b200
    gf = pyrodigal_b200.GeneFinder(closed=closed, mask=mask)
    prefix = None
    if tmp is not None:
        prefix = str(tmp / "dump")
        os.environ["PGPU_TRAIN_DUMP"] = prefix
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ti = gf.train(seq, force_nonsd=force, start_weight=st_wt, translation_table=tt)
    finally:
        os.environ.pop("PGPU_TRAIN_DUMP", None)
    blob = bytes(ti)
    if expect is not None and blob != expect:
        msg = f"training struct differs in {diff_fields(blob, expect)}"
        if prefix:
            msg += "\n" + diagnose(prefix, seq, closed, mask, force, tt, st_wt)
        raise AssertionError(msg)
    return gf, ti


@pytest.mark.parametrize("name", list(TRAIN["names"]))
def test_train_golden(name, tmp_path):
    closed, mask, force, tt = (int(v) for v in TRAIN[name + "/opts"])
    gpu_train(TRAIN[name + "/seq"].tobytes(), closed, mask, force, tt, float(TRAIN[name + "/start_weight"]),
              tmp=tmp_path, expect=TRAIN[name + "/tinf"].tobytes())


def test_published_training_scalars():
    """tests/test_gene_finder.py:329-345"""
    gf, info = gpu_train(TRAIN["srr_contig/seq"].tobytes())
    e = TRAIN["srr_expected"]
    assert info.translation_table == 11 and info.start_weight == 4.35 and info.uses_sd
    assert [info.gc, *info.bias, *info.type_weights] == list(e)
    assert gf.training_info is info


def oracle_train(seq, closed=False, mask=False, force=False, tt=11, st_wt=4.35):
    d, gc, unk = orc.encode(seq)
    masks = orc.find_masks(d, 50) if mask else None
    return orc.train(d, gc / len(d), translation_table=tt, start_weight=st_wt, force_nonsd=force,
                     opts=orc.make_opts(closed=closed, masks=masks))


@pytest.mark.parametrize("length,gc,kw", [
    (20000, .5, {}), (45001, .3, {}), (45002, .7, dict(closed=True)), (300000, .52, dict(force=True)),
    (33333, .41, dict(tt=4)), (50000, .6, dict(st_wt=2.5)), (30001, .5, dict(mask=True, n_frac=0.003)),
    (1200000, .48, {}), (800000, .66, dict(tt=4, force=True)),
], ids=lambda v: str(v).replace(" ", ""))
def test_train_vs_oracle(length, gc, kw, tmp_path):
    kw = dict(kw)
    seq = R.synth(length, gc, seed=9000 + length, n_frac=kw.pop("n_frac", 0.0))
    gpu_train(seq, tmp=tmp_path, expect=oracle_train(seq, **kw), **kw)


def test_train_multi_contig_linker():
    """lib.pyx:5534-5541: contigs are joined by TTAATTAATTAA, with one trailing linker"""
    parts = [R.synth(15000, .5, seed=s) for s in (1, 2, 3)]
    import pyrodigal_b200
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ti = pyrodigal_b200.GeneFinder().train(*parts)
        ts = pyrodigal_b200.GeneFinder().train(*(p.decode() for p in parts))
    want = oracle_train(b"TTAATTAATTAA".join(parts + [b""]))
    assert bytes(ti) == want and bytes(ts) == want


def test_train_then_find_genes_single_mode():
    """BASELINE config 2 in miniature: train on a genome, then find_genes with the trained model"""
    seq = R.synth(600000, .5, seed=4242)
    gf, ti = gpu_train(seq)
    g = gf.find_genes(seq)
    d, gc, unk = orc.encode(seq)
    genes, nodes, ipath = orc.find_genes_single(d, oracle_train(seq))
    assert [(x.begin, x.end, x.strand) for x in g] == \
        [(int(a["begin"]), int(a["end"]), int(nodes[a["start_ndx"]]["strand"])) for a in genes]
    assert np.array_equal(g.nodes.array["cscore"], nodes["cscore"]) and np.array_equal(g.nodes.array["sscore"], nodes["sscore"])


def test_train_errors():
    import pyrodigal_b200
    from pyrodigal_b200 import _capi
    c = _capi.Context(0)
    seq = np.frombuffer(R.synth(30000, .5, seed=1), dtype=np.uint8)
    with pytest.raises(ValueError, match="at least 20000"):
        c.train(seq[:1000], _capi.make_opts())
    with pytest.raises(ValueError, match="translation table"):
        c.train(seq, _capi.make_opts(), translation_table=7)
    with pytest.raises(RuntimeError, match="metagenomic"):
        c.train(seq, _capi.make_opts(meta=True))
    c.close()
