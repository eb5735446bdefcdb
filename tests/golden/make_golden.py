"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/build_ref.sh).  Run in the build container only:

    python tests/golden/make_golden.py

Fixtures (all small):
  meta_cases.npz    find_genes(meta=True) on seeded synthetic contigs + two real contigs from the
                    reference's own test data: winner bin, genes (begin,end,start_ndx,stop_ndx),
                    the final node array (Nodes.__getstate__), Prodigal-CLI gene headers.
  single_cases.npz  single mode (trained TrainingInfo blob, genes, nodes incl. DP state).
  dp_cases.npz      operator level: node arrays in, ConnectionScorer.score_connections out,
                    final=True and final=False (the reference's tests/test_connection_scorer.py
                    protocol, with real scores / star_ptr / gc_score injected via __setstate__).
  train_cases.npz   GeneFinder.train(): sequence + options in, raw TrainingInfo struct out, incl. the
                    reference's own golden (tests/test_training_info.py:60-66, the 100 kb slice trained with
                    closed=True must equal GCF_..._100kb.tinf_closed.bin.gz) and the contig of
                    tests/test_gene_finder.py:329-345; plus Sequence.max_gc_frame_plot() of every case.
  writer_cases.npz  the reference's command line (pyrodigal.cli.main) run on small FASTA inputs: GFF / GenBank, protein
                    and nucleotide FASTA, score tables, and the training file it writes in single mode -- byte for byte
                    what the drop-in CLI and writers must produce.
  misc.npz          node counts per translation table (tests/test_nodes.py:28-39), Shine-Dalgarno
                    known answers (tests/test_sequence.py:52-75).
"""
import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as R  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (only for the NODE_DTYPE container + scoring of dp inputs)

pyrodigal = R.reference()
BINS = list(pyrodigal.METAGENOMIC_BINS)


def genes_array(g):
    st = g.__getstate__()
    a = np.zeros(len(st["genes"]), dtype=orc.GENE_DTYPE)
    for k, x in enumerate(st["genes"]):
        a[k] = (x["begin"], x["end"], x["start_ndx"], x["stop_ndx"])
    return a


def gene_table(g):
    """per-gene public attributes the drop-in API must reproduce"""
    rows = []
    for x in g:
        rows.append((x.begin, x.end, x.strand, int(x.partial_begin), int(x.partial_end), x.start_type,
                     str(x.rbs_motif), str(x.rbs_spacer), x.gc_cont, x.cscore, x.rscore, x.sscore, x.tscore,
                     x.uscore, x.score, x.confidence()))
    return rows


def prodigal_headers(name, mode):
    p = os.path.join(R.REF_DATA, f"{name}.{mode}.fna.gz")
    if not os.path.exists(p):
        return []
    out = []
    with gzip.open(p, "rt") as f:
        for line in f:
            if line.startswith(">"):
                out.append(line[1:].strip())
    return out


def meta_cases():
    cases = {}
    specs = [
        ("cfg1_10k", dict(length=10000, gc=0.5, seed=1234), {}),
        ("s3000", dict(length=3000, gc=0.4, seed=101), {}),
        ("s1200", dict(length=1200, gc=0.6, seed=102), {}),
        ("s2999", dict(length=2999, gc=0.5, seed=109), {}),
        ("s50k_lowgc", dict(length=50000, gc=0.35, seed=103), {}),
        ("s30k_closed", dict(length=30000, gc=0.55, seed=104), dict(closed=True)),
        ("s20k_N_mask", dict(length=20000, gc=0.45, seed=105, n_frac=0.002), dict(mask=True)),
        ("s20k_N_nomask", dict(length=20000, gc=0.45, seed=105, n_frac=0.002), {}),
        ("s200", dict(length=200, gc=0.5, seed=106), {}),
        ("s20", dict(length=20, gc=0.5, seed=107), {}),
        ("s0", dict(length=0, gc=0.5, seed=108), {}),
    ]
    for name, sk, gk in specs:
        seq = R.synth(**sk)
        cases[name] = (seq, gk, None)
    # mask=True with a trailing N run and an N run inside the last open reading frame: the reference's per-frame
    # mask cursor (lib.pyx:1959-1966) keeps starts that an "intersects any mask" test would drop
    for seed in (0, 2, 7):
        cases[f"trailing_N_mask_{seed}"] = (R.trailing_n_case(seed), dict(mask=True), None)
    for fn in ("KK037166", "SRR492066"):
        _, s = R.read_fasta_gz(os.path.join(R.REF_DATA, fn + ".fna.gz"))[0]
        cases[fn] = (s.encode(), {}, fn)
    out = {"names": np.array(list(cases))}
    for name, (seq, gk, real) in cases.items():
        gf = pyrodigal.GeneFinder(meta=True, **gk)
        g = gf.find_genes(seq)
        out[name + "/seq"] = np.frombuffer(seq, dtype=np.uint8)
        out[name + "/opts"] = np.array([int(gk.get("closed", False)), int(gk.get("mask", False))])
        out[name + "/winner"] = np.array(BINS.index(g.metagenomic_bin) if g.metagenomic_bin is not None else -1)
        out[name + "/genes"] = genes_array(g)
        out[name + "/nodes"] = R.ref_nodes_to_array(g.nodes)
        out[name + "/gene_table"] = np.array(gene_table(g), dtype=object)
        out[name + "/prodigal"] = np.array(prodigal_headers(real, "meta") if real else [], dtype=object)
    np.savez_compressed(os.path.join(HERE, "meta_cases.npz"), **out)
    print("meta_cases", len(cases))


def single_cases():
    out = {}
    names = []
    # (a) single mode with a built-in model as training info, (b) a genuinely trained model
    _, kk = R.read_fasta_gz(os.path.join(R.REF_DATA, "KK037166.fna.gz"))[0]
    _, srr = R.read_fasta_gz(os.path.join(R.REF_DATA, "SRR492066.fna.gz"))[0]
    specs = [
        ("kk_bin20", kk.encode(), BINS[20].training_info, {}, None),
        ("s40k_bin0_tt4", R.synth(40000, 0.4, seed=301), BINS[0].training_info, {}, None),
        ("s40k_bin33_closed", R.synth(40000, 0.6, seed=302), BINS[33].training_info, dict(closed=True), None),
    ]
    gf = pyrodigal.GeneFinder()
    tinf = gf.train(srr)
    specs.append(("srr_trained", srr.encode(), tinf, {}, "SRR492066"))
    gf2 = pyrodigal.GeneFinder()
    tinf2 = gf2.train(kk)
    specs.append(("kk_trained", kk.encode(), tinf2, {}, "KK037166"))
    for name, seq, ti, gk, real in specs:
        g = pyrodigal.GeneFinder(ti, **gk).find_genes(seq)
        names.append(name)
        out[name + "/seq"] = np.frombuffer(seq, dtype=np.uint8)
        out[name + "/opts"] = np.array([int(gk.get("closed", False)), int(gk.get("mask", False))])
        out[name + "/tinf"] = np.frombuffer(bytes(memoryview(ti)), dtype=np.uint8)
        out[name + "/genes"] = genes_array(g)
        out[name + "/nodes"] = R.ref_nodes_to_array(g.nodes)
        out[name + "/gene_table"] = np.array(gene_table(g), dtype=object)
        out[name + "/prodigal"] = np.array(prodigal_headers(real, "single") if real else [], dtype=object)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "single_cases.npz"), **out)
    print("single_cases", len(names))


def nodes_from_array(arr):
    """oracle-layout array -> reference Nodes via __setstate__ (lib.pyx:1754-1794)"""
    st = []
    for o in arr:
        st.append({
            "type": int(o["type"]), "edge": bool(o["edge"]), "ndx": int(o["ndx"]), "strand": int(o["strand"]),
            "stop_val": int(o["stop_val"]), "star_ptr": [int(v) for v in o["star_ptr"]],
            "gc_bias": int(o["gc_bias"]), "gc_score": [float(v) for v in o["gc_score"]],
            "cscore": float(o["cscore"]), "gc_cont": float(o["gc_cont"]), "rbs": [int(v) for v in o["rbs"]],
            "motif": {"ndx": int(o["mot_ndx"]), "len": int(o["mot_len"]), "spacer": int(o["mot_spacer"]),
                      "spacendx": int(o["mot_spacendx"]), "score": float(o["mot_score"])},
            "uscore": float(o["uscore"]), "tscore": float(o["tscore"]), "rscore": float(o["rscore"]),
            "sscore": float(o["sscore"]), "traceb": int(o["traceb"]), "tracef": int(o["tracef"]),
            "ov_mark": int(o["ov_mark"]), "score": float(o["score"]), "elim": bool(o["elim"]),
        })
    n = pyrodigal.Nodes()
    n.__setstate__(st)
    return n


def dp_inputs(seq, bin_index, seed, closed=False):
    """Scored nodes + star_ptr + synthetic gc_score, produced by the REFERENCE where it has a public
    entry point (extract/sort/reset/score) and by the oracle for record_overlapping_starts (no
    public entry point in the reference; its result is validated through find_genes parity)."""
    ti = BINS[bin_index].training_info
    s = pyrodigal.Sequence(seq)
    nodes = pyrodigal.Nodes()
    nodes.extract(s, translation_table=ti.translation_table, closed=closed)
    nodes.sort()
    nodes.reset_scores()
    nodes.score(s, ti, closed=closed, is_meta=True)
    arr = R.ref_nodes_to_array(nodes)
    orc.record_overlapping_starts(arr, R.bin_blob(bin_index), flag=1, max_overlap=60)
    rng = np.random.default_rng(seed)
    arr["gc_score"] = rng.uniform(-0.5, 1.5, size=(len(arr), 3)).round(3)
    arr["gc_bias"] = rng.integers(0, 3, size=len(arr))
    return arr


def dp_cases():
    out = {}
    names = []
    _, kk = R.read_fasta_gz(os.path.join(R.REF_DATA, "KK037166.fna.gz"))[0]
    specs = [
        ("kk_bin38", kk.encode(), 38),
        ("s30k_bin0_tt4", R.synth(30000, 0.5, seed=401), 0),   # tt=4 at GC .5: giant-ORF windows (T2)
        ("s60k_bin25", R.synth(60000, 0.6, seed=402), 25),
        ("s5k_bin10", R.synth(5000, 0.35, seed=403), 10),
    ]
    for name, seq, b in specs:
        arr = dp_inputs(seq, b, seed=len(seq) + b)
        ti = BINS[b].training_info
        names.append(name)
        out[name + "/bin"] = np.array(b)
        out[name + "/in"] = arr
        for final in (True, False):
            n = nodes_from_array(arr)
            sc = pyrodigal.lib.ConnectionScorer(backend="generic")
            sc.index(n)
            sc.score_connections(n, ti, final=final)
            res = R.ref_nodes_to_array(n)
            # also the unfiltered Prodigal score_connection (backend=None) must agree
            n0 = nodes_from_array(arr)
            sc0 = pyrodigal.lib.ConnectionScorer(backend=None)
            sc0.index(n0)
            sc0.score_connections(n0, ti, final=final)
            res0 = R.ref_nodes_to_array(n0)
            assert np.array_equal(res["traceb"], res0["traceb"]) and np.array_equal(res["score"], res0["score"])
            tag = "final" if final else "train"
            out[f"{name}/{tag}/score"] = res["score"]
            out[f"{name}/{tag}/traceb"] = res["traceb"]
            out[f"{name}/{tag}/ov_mark"] = res["ov_mark"]
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "dp_cases.npz"), **out)
    print("dp_cases", len(names))


def misc():
    out = {}
    # node counts per translation table: tests/test_nodes.py:28-39
    _, srr = R.read_fasta_gz(os.path.join(R.REF_DATA, "SRR492066.fna.gz"))[0]
    s = pyrodigal.Sequence(srr)
    counts = []
    for tt in sorted(pyrodigal.TRANSLATION_TABLES):
        n = pyrodigal.Nodes()
        counts.append((tt, n.extract(s, translation_table=tt)))
    out["srr_node_counts"] = np.array(counts)
    # Shine-Dalgarno known answers on a real contig, both strands, exact & mismatch
    ti = BINS[20].training_info
    rows = []
    rng = np.random.default_rng(7)
    for _ in range(4000):
        start = int(rng.integers(0, len(srr)))
        pos = start - int(rng.integers(5, 21))
        if pos < 0:
            continue
        for strand in (1, -1):
            for exact in (True, False):
                rows.append((pos, start, strand, int(exact), s.shine_dalgarno(pos, start, ti, strand=strand, exact=exact)))
    out["srr_sd"] = np.array(rows)
    out["srr_seq"] = np.frombuffer(srr.encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "misc.npz"), **out)
    print("misc ok")


def train_cases():
    import warnings
    out = {}
    names = []
    _, srr = R.read_fasta_gz(os.path.join(R.REF_DATA, "SRR492066.fna.gz"))[0]
    recs = R.read_fasta_gz(os.path.join(R.REF_DATA, "GCF_001457455.1_NCTC11397_genomic_100kb.fna.gz"))
    g100 = "TTAATTAATTAA".join([s for _, s in recs] + [""]) if len(recs) > 1 else recs[0][1]
    with gzip.open(os.path.join(R.REF_DATA, "GCF_001457455.1_NCTC11397_genomic_100kb.tinf_closed.bin.gz"), "rb") as f:
        expected_100k = bytes(memoryview(pyrodigal.TrainingInfo.load(f)))
    # name, sequence, GeneFinder kwargs, train kwargs
    specs = [
        ("ref100k_closed", g100.encode(), dict(closed=True), {}),
        ("ref100k_open", g100.encode(), {}, {}),
        ("srr_contig", srr.encode(), {}, {}),
        ("s60k_nonsd", R.synth(60000, 0.45, seed=501), {}, {}),
        ("s30k_tt4_forced", R.synth(30000, 0.38, seed=502), {}, dict(force_nonsd=True, translation_table=4)),
        ("s40k_N_mask", R.synth(40000, 0.55, seed=503, n_frac=0.002), dict(mask=True), dict(start_weight=3.9)),
        ("s25k_closed_hi", R.synth(25000, 0.68, seed=504), dict(closed=True), {}),
        ("s20000_min", R.synth(20000, 0.5, seed=505), {}, {}),
    ]
    for name, seq, gk, tk in specs:
        gf = pyrodigal.GeneFinder(**gk)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ti = gf.train(seq, **tk)
        blob = bytes(memoryview(ti))
        if name == "ref100k_closed":
            assert blob == expected_100k, "reference build does not reproduce its own training golden"
        names.append(name)
        out[name + "/seq"] = np.frombuffer(seq, dtype=np.uint8)
        out[name + "/opts"] = np.array([int(gk.get("closed", False)), int(gk.get("mask", False)),
                                        int(tk.get("force_nonsd", False)), int(tk.get("translation_table", 11))])
        out[name + "/start_weight"] = np.array(float(tk.get("start_weight", 4.35)))
        out[name + "/tinf"] = np.frombuffer(blob, dtype=np.uint8)
        out[name + "/gc_frame"] = np.array(pyrodigal.Sequence(seq).max_gc_frame_plot(), dtype=np.int8)
    # the published scalars of tests/test_gene_finder.py:329-345
    out["srr_expected"] = np.array([0.3010045159434068, 2.6770525781861187, 0.17260535063729165, 0.1503420711765898,
                                    0.71796361273324, -1.3722361344058844, -2.136731395763296])
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "train_cases.npz"), **out)
    print("train_cases", len(names))


def writer_cases():
    import io
    import tempfile
    from pyrodigal import cli
    out = {}
    names = []
    _, srr = R.read_fasta_gz(os.path.join(R.REF_DATA, "SRR492066.fna.gz"))[0]

    def fasta_text(records, width=70):
        lines = []
        for name, seq in records:
            lines.append(">" + name)
            lines.extend(seq[i:i + width] for i in range(0, len(seq), width))
        return "\n".join(lines) + "\n"

    specs = [
        ("srr_meta_gff", fasta_text([("SRR492066 test contig", srr)]), ["-p", "meta"]),
        ("multi_meta_gbk", fasta_text([(f"ctg{k} synthetic", R.synth(L, gc, seed=600 + k, n_frac=nf).decode())
                                       for k, (L, gc, nf) in enumerate([(9000, .4, 0), (700, .5, 0), (15000, .62, .002), (90, .5, 0)])]),
         ["-p", "meta", "-f", "gbk", "-m", "--no-stop-codon"]),
        ("multi_single_train", fasta_text([(f"chr{k}", R.synth(L, .5, seed=610 + k).decode()) for k, L in enumerate([30000, 12000, 8000])]),
         ["-p", "single", "-c"]),
        ("single_nonsd_tt4", fasta_text([("g", R.synth(40000, .38, seed=620).decode())]), ["-p", "single", "-n", "-g", "4"]),
    ]
    for name, text, argv in specs:
        with tempfile.TemporaryDirectory() as tmp:
            fa = os.path.join(tmp, "in.fna")
            with open(fa, "w") as f:
                f.write(text)
            paths = {k: os.path.join(tmp, k) for k in ("o", "a", "d", "s", "t")}
            full = ["-i", fa, "-o", paths["o"], "-a", paths["a"], "-d", paths["d"], "-s", paths["s"]] + argv
            if "single" in argv:
                full += ["-t", paths["t"]]
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                rc = cli.main(full, stdout=io.StringIO(), stderr=io.StringIO())
            assert rc == 0
            names.append(name)
            out[name + "/fasta"] = np.frombuffer(text.encode(), dtype=np.uint8)
            out[name + "/argv"] = np.array(argv, dtype=object)
            for k in ("o", "a", "d", "s"):
                out[f"{name}/{k}"] = np.frombuffer(open(paths[k], "rb").read(), dtype=np.uint8)
            if os.path.exists(paths["t"]):
                out[name + "/t"] = np.frombuffer(open(paths["t"], "rb").read(), dtype=np.uint8)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "writer_cases.npz"), **out)
    print("writer_cases", len(names))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "writers":
        writer_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "meta":
        meta_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_cases()
        sys.exit(0)
    meta_cases()
    single_cases()
    dp_cases()
    misc()
    train_cases()
    writer_cases()
