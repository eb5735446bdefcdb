"""CPU tests of the host-side halves of the "next" rows: FASTA ingest, the output writers and
Gene.translate.  Expected texts come from tests/golden/writer_cases.npz, produced by the reference's own command
line (pyrodigal.cli.main); the gene / node records are filled from the oracle here (no GPU) and from the GPU in
tests/test_gpu_cli.py.  Everything is compared byte for byte."""
import io
import os
import warnings

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
W = np.load(os.path.join(G, "writer_cases.npz"), allow_pickle=True)


def run_writers(name):
    from pyrodigal_b200 import fasta
    argv = list(W[name + "/argv"])
    batch = fasta.read_batch(W[name + "/fasta"].tobytes())
    meta = argv[argv.index("-p") + 1] == "meta"
    closed, mask, gbk = "-c" in argv, "-m" in argv, "gbk" in argv
    tinf = None
    if not meta:
        tt = int(argv[argv.index("-g") + 1]) if "-g" in argv else 11
        joined = b"TTAATTAATTAA".join([batch.sequence(k).tobytes() for k in range(len(batch))] + [b""]) if len(batch) > 1 \
            else batch.sequence(0).tobytes()
        d, gc, unk = orc.encode(joined)
        tinf = orc.train(d, gc / len(d), translation_table=tt, force_nonsd="-n" in argv,
                         opts=orc.make_opts(closed=closed, masks=orc.find_masks(d, 50) if mask else None))
    o, a, dd, s = io.StringIO(), io.StringIO(), io.StringIO(), io.StringIO()
    for k in range(len(batch)):
        genes = R.genes_from_oracle(batch.sequence(k).tobytes(), meta=meta, tinf_blob=tinf, closed=closed, mask=mask, num_seq=k + 1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            (genes.write_genbank if gbk else genes.write_gff)(o, batch.ids[k])
            genes.write_translations(a, batch.ids[k], include_stop="--no-stop-codon" not in argv)
            genes.write_genes(dd, batch.ids[k])
            genes.write_scores(s, batch.ids[k])
    return {"o": o.getvalue(), "a": a.getvalue(), "d": dd.getvalue(), "s": s.getvalue()}, tinf


def first_difference(mine, want):
    for i, (x, y) in enumerate(zip(mine.splitlines(), want.splitlines())):
        if x != y:
            return f"line {i}: {x[:160]!r} != {y[:160]!r}"
    return f"lengths {len(mine)} vs {len(want)}"


@pytest.mark.parametrize("name", list(W["names"]))
def test_writers_match_reference_cli(name):
    import datetime
    got, tinf = run_writers(name)
    for k in ("o", "a", "d", "s"):
        want = W[f"{name}/{k}"].tobytes().decode()
        if k == "o" and "gbk" in list(W[name + "/argv"]):
            # the LOCUS line carries today's date
            today = datetime.date.today().strftime("%d-%b-%y").upper()
            fix = lambda t: "\n".join(l[:l.rfind(" ") + 1] + today if l.startswith("LOCUS") else l for l in t.split("\n"))
            got[k], want = fix(got[k]), fix(want)
        assert got[k] == want, f"{name}/{k}: {first_difference(got[k], want)}"
    if name + "/t" in W.files:
        assert tinf == W[name + "/t"].tobytes()


def test_fasta_batch_layout_and_edge_cases():
    from pyrodigal_b200 import fasta
    txt = b">a desc one\r\nACGT\r\nAC\r\n\r\n>b\nGGG\n>c only header\n>d\nTT"
    b = fasta.read_batch(txt)
    assert b.ids == ["a", "b", "c", "d"] and b.descriptions == ["desc one", "", "only header", ""]
    assert b.flat.tobytes() == b"ACGTACGGGTT" and list(b.offsets) == [0, 6, 9, 9, 11]
    assert [r.seq for r in fasta.parse(io.BytesIO(txt))] == ["ACGTAC", "GGG", "", "TT"]
    assert len(fasta.read_batch(b"")) == 0
    with pytest.raises(ValueError, match="not in FASTA format"):
        fasta.read_batch(b"ACGT\nACGT\n")
    import gzip
    assert fasta.read_batch(gzip.compress(txt)).flat.tobytes() == b"ACGTACGGGTT"
    big = W["multi_meta_gbk/fasta"].tobytes()
    bb = fasta.read_batch(big)
    assert len(bb) == 4 and [len(bb.sequence(k)) for k in range(4)] == [9000, 700, 15000, 90]


def test_translation_tables_and_options():
    """Gene.translate: alternative tables, non-strict handling of unknown bases, stop trimming (lib.pyx:2926-3047)"""
    from pyrodigal_b200 import lib as L
    assert L._translation_table(11).tobytes() == b"KKNNRRSSTTTTIMIIEEDDGGGGAAAAVVVVQQHHRRRRPPPPLLLL**YY*WCCSSSSLLFF"
    assert L._translation_table(4)[(3 << 4) + (1 << 2) + 0] == ord("W")          # TGA in the Mycoplasma code
    seq = b"ATGGCNCCNAAATGANNNTAA" + b"ACGT" * 10
    g = R.genes_from_oracle(seq, meta=True)
    g._genes = np.array([(1, 21, 0, 0)], dtype=g._genes.dtype)
    nd = np.zeros((1, 2), dtype=g._gene_nodes.dtype)
    nd["strand"] = 1
    g._gene_nodes = nd
    g.training_info = L._LazyBins.get()[20].training_info
    gene = g[0]
    assert gene.translate() == "MXXK*X*"
    assert gene.translate(strict=False) == "MAPK*X*"
    assert gene.translate(strict=False, include_stop=False, unknown_residue="?") == "MAPK*?"
    with pytest.warns(UserWarning, match="different STOP codons"):
        assert gene.translate(4) == "MXXKWX*"
    with pytest.raises(ValueError):
        gene.translate(7)
