"""Helpers shared by parity tests: access to the unmodified reference (oracle/_ref, only in
the build container), synthetic sequences, node/gene comparison."""
import gzip
import lzma
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_DATA = os.path.join(REF_DIR, "pyrodigal", "tests", "data")
TRAINING_SIZE = 558392

TYPE_NAMES = {"ATG": 0, "GTG": 1, "TTG": 2, "Edge": 0, "STOP": 3}


def have_reference():
    return os.path.exists(os.path.join(REF_DIR, "pyrodigal", "__init__.py"))


def reference():
    """import the unmodified reference build (oracle/_ref)"""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import pyrodigal

    return pyrodigal


_bins = None


def bins_blob():
    global _bins
    if _bins is None:
        with lzma.open(os.path.join(ROOT, "pyrodigal_b200", "data", "metagenomic_bins.bin.xz")) as f:
            _bins = f.read()
        assert len(_bins) == 50 * TRAINING_SIZE
    return _bins


def bin_blob(i):
    return bins_blob()[i * TRAINING_SIZE:(i + 1) * TRAINING_SIZE]


def synth(length, gc=0.5, seed=0, n_frac=0.0):
    """iid nucleotides, P(A)=P(T)=(1-gc)/2, P(C)=P(G)=gc/2"""
    rng = np.random.default_rng(seed)
    p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
    a = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=length, p=p)
    if n_frac > 0:
        # sprinkle single Ns and a few long runs
        idx = rng.random(length) < n_frac
        a = a.copy()
        a[idx] = ord("N")
        for _ in range(max(1, length // 20000)):
            s = int(rng.integers(0, max(1, length - 200)))
            a[s:s + int(rng.integers(30, 150))] = ord("N")
    return a.tobytes()


def trailing_n_case(seed):
    """the advisor's scenario (round 1): a contig that ends in 3-5 Ns -- always a mask, lib.pyx:711-712 -- with an N run of
    >= 50 inside the last open reading frame: the forward cursor rests on the trailing mask (begin == last), so the inner
    run is never tested and the reference keeps the starts that span it"""
    rng = np.random.default_rng(seed)
    L = int(rng.integers(2000, 5000))
    a = np.frombuffer(synth(L, .7, 3000 + seed), np.uint8).copy()
    tail, inner = int(rng.integers(3, 6)), int(rng.integers(50, 90))
    pos = L - int(rng.integers(150, 500))
    a[pos:pos + inner] = ord("N")
    a[L - tail:] = ord("N")
    return a.tobytes()


def read_fasta_gz(path):
    seqs, name, buf = [], None, []
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    seqs.append((name, "".join(buf)))
                name, buf = line[1:].split()[0], []
            else:
                buf.append(line.strip())
    if name is not None:
        seqs.append((name, "".join(buf)))
    return seqs


def ref_nodes_to_array(nodes):
    """reference Nodes -> oracle NODE_DTYPE array via Nodes.__getstate__ (lib.pyx:1708-1752)"""
    from oracle import oracle as orc

    st = nodes.__getstate__()
    out = np.zeros(len(st), dtype=orc.NODE_DTYPE)
    for k, s in enumerate(st):
        o = out[k]
        o["ndx"] = s["ndx"]; o["stop_val"] = s["stop_val"]; o["strand"] = s["strand"]; o["type"] = s["type"]
        o["edge"] = s["edge"]; o["elim"] = s["elim"]; o["gc_bias"] = s["gc_bias"]
        o["star_ptr"] = s["star_ptr"]; o["traceb"] = s["traceb"]; o["tracef"] = s["tracef"]
        o["ov_mark"] = s["ov_mark"]; o["rbs"] = s["rbs"]
        m = s["motif"]
        o["mot_ndx"] = m["ndx"]; o["mot_len"] = m["len"]; o["mot_spacer"] = m["spacer"]
        o["mot_spacendx"] = m["spacendx"]
        o["gc_score"] = s["gc_score"]
        for f in ("cscore", "uscore", "tscore", "rscore", "sscore", "score", "gc_cont"):
            o[f] = s[f]
    return out


INT_FIELDS = ("ndx", "stop_val", "strand", "type", "edge")
FLOAT_FIELDS = ("cscore", "uscore", "tscore", "rscore", "sscore")


def assert_nodes_equal(a, b, ints=INT_FIELDS, floats=FLOAT_FIELDS, tol=0.0, what=""):
    assert len(a) == len(b), f"{what}: node count {len(a)} != {len(b)}"
    for f in ints:
        if not np.array_equal(a[f], b[f]):
            bad = np.argwhere(a[f] != b[f])[:5].ravel()
            raise AssertionError(f"{what}: int field {f} differs at {bad}: {a[f][bad]} vs {b[f][bad]}")
    for f in floats:
        x, y = np.asarray(a[f], dtype=np.float64), np.asarray(b[f], dtype=np.float64)
        d = np.abs(x - y)
        d[np.isnan(d)] = np.inf
        d[(np.isnan(x) & np.isnan(y))] = 0
        if d.size and d.max() > tol:
            k = int(np.argmax(d))
            raise AssertionError(f"{what}: float field {f} differs at {k}: {x[k]!r} vs {y[k]!r} (tol {tol})")


def genes_from_oracle(seq, meta=True, tinf_blob=None, closed=False, mask=False, num_seq=1):
    """A pyrodigal_b200.Genes built from ORACLE results, for CPU tests of the host-side writers (the product fills
    the same records from the GPU; this helper exists so that formatting can be tested where there is no GPU)."""
    import pyrodigal_b200 as P
    from oracle import oracle as orc
    from pyrodigal_b200 import _capi, lib as L
    if isinstance(seq, str):
        seq = seq.encode()
    d, gc, unk = orc.encode(seq)
    masks = orc.find_masks(d, 50) if mask else None
    opts = orc.make_opts(closed=closed, masks=masks)
    if meta:
        genes, nodes, winner, _ = orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, bins_blob(), opts)
        mbin = P.METAGENOMIC_BINS[winner] if winner >= 0 else None
        ti, ipath = (mbin.training_info if mbin else None), -1
    else:
        genes, nodes, ipath = orc.find_genes_single(d, tinf_blob, opts)
        mbin, ti = None, L.TrainingInfo._from_bytes(tinf_blob)
    a = np.zeros(len(nodes), dtype=_capi.NODE_DTYPE)
    for f in _capi.NODE_DTYPE.names:
        a[f] = nodes[f]
    g = np.zeros(len(genes), dtype=_capi.GENE_DTYPE)
    for f in ("begin", "end", "start_ndx", "stop_ndx"):
        g[f] = genes[f]
    gn = np.zeros((len(genes), 2), dtype=_capi.NODE_DTYPE)
    if len(genes):
        gn[:, 0], gn[:, 1] = a[g["start_ndx"]], a[g["stop_ndx"]]
    return L.Genes(g, gn, sequence=L.Sequence(seq, mask=mask), training_info=ti, metagenomic_bin=mbin, meta=meta,
                   nodes=L.Nodes(a), ipath=ipath, num_seq=num_seq)
