"""GPU test of the command line and the writers end to end (run with -m gpu on a B200): pyrodigal_b200.cli.main on
the FASTA inputs of tests/golden/writer_cases.npz must reproduce, byte for byte, what the reference's command line
(pyrodigal.cli.main) wrote for the same arguments -- GFF / GenBank, protein and nucleotide FASTA, score tables and,
in single mode, the training file produced by `GeneFinder.train` on the GPU."""
import datetime
import io
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
W = np.load(os.path.join(G, "writer_cases.npz"), allow_pickle=True)


def first_difference(mine, want):
    for i, (x, y) in enumerate(zip(mine.splitlines(), want.splitlines())):
        if x != y:
            return f"line {i}: {x[:160]!r} != {y[:160]!r}"
    return f"lengths {len(mine)} vs {len(want)}"


@pytest.mark.parametrize("name", list(W["names"]))
def test_cli_matches_reference_cli(name, tmp_path):
    from pyrodigal_b200 import cli
    argv = list(W[name + "/argv"])
    fa = tmp_path / "in.fna"
    fa.write_bytes(W[name + "/fasta"].tobytes())
    paths = {k: str(tmp_path / k) for k in ("o", "a", "d", "s", "t")}
    full = ["-i", str(fa), "-o", paths["o"], "-a", paths["a"], "-d", paths["d"], "-s", paths["s"]] + argv
    if "single" in argv:
        full += ["-t", paths["t"]]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert cli.main(full, stdout=io.StringIO(), stderr=io.StringIO()) == 0
    for k in ("o", "a", "d", "s"):
        got, want = open(paths[k]).read(), W[f"{name}/{k}"].tobytes().decode()
        if k == "o" and "gbk" in argv:
            today = datetime.date.today().strftime("%d-%b-%y").upper()
            fix = lambda t: "\n".join(l[:l.rfind(" ") + 1] + today if l.startswith("LOCUS") else l for l in t.split("\n"))
            got, want = fix(got), fix(want)
        assert got == want, f"{name}/{k}: {first_difference(got, want)}"
    if name + "/t" in W.files:
        assert open(paths["t"], "rb").read() == W[name + "/t"].tobytes()
        # second run: the training file now exists and is read back instead of training again (cli.py:238-243)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert cli.main(full, stdout=io.StringIO(), stderr=io.StringIO()) == 0
        assert open(paths["o"]).read() == W[f"{name}/o"].tobytes().decode()


def test_cli_rejects_training_file_in_meta_mode(tmp_path):
    from pyrodigal_b200 import cli
    fa = tmp_path / "in.fna"
    fa.write_text(">a\nACGT\n")
    err = io.StringIO()
    assert cli.main(["-i", str(fa), "-p", "meta", "-t", str(tmp_path / "t.bin")], stdout=io.StringIO(), stderr=err) == 1
    assert "cannot specify metagenomic sequence with a training file" in err.getvalue()
