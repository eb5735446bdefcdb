"""CPU tests of the host logic: the C-ABI library loads and exports every symbol include/pyrodigal_b200.h
declares (no compute without a GPU), the Python mirror fails loudly without a device, and the N > 1
sharding / reduction plumbing of bench.py works with world_size 2 over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import refutil as R

ROOT = R.ROOT


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pyrodigal_b200.h")).read()
    names = sorted(set(re.findall(r"\b(pgpu_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "pyrodigal_b200", "libpyrodigal_b200.so"))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import pyrodigal_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pyrodigal_b200.GeneFinder(meta=True).find_genes(b"ACGT" * 100)


def test_python_surface_argument_errors():
    import pyrodigal_b200 as p
    ti = p.METAGENOMIC_BINS[0].training_info
    with pytest.raises(ValueError):
        p.GeneFinder(ti, meta=True)
    with pytest.raises(ValueError):
        p.GeneFinder(meta=True, min_gene=0)
    with pytest.raises(ValueError):
        p.GeneFinder(meta=True, max_overlap=100, min_gene=90)
    with pytest.raises(RuntimeError):
        p.GeneFinder(meta=True).train(b"ACGT")
    assert len(p.METAGENOMIC_BINS) == 50 and p.METAGENOMIC_BINS[0].training_info.translation_table == 4
    t2 = p.TrainingInfo._from_bytes(bytes(ti))
    assert t2 == ti and t2.gc == ti.gc and t2.uses_sd == ti.uses_sd
    s = p.Sequence("ACGTNNNNacgt", mask=True, mask_size=3)
    assert len(s) == 12 and s.unknown == 4 and abs(s.gc - 4 / 12) < 1e-12 and [(m.begin, m.end) for m in s.masks] == [(4, 8)]


def test_bench_shards_are_disjoint_slices_of_the_config():
    import bench
    a, oa = bench.make_contigs(*bench.shard_range(0, 6))
    b, ob = bench.make_contigs(*bench.shard_range(1, 6))
    ab, oab = bench.make_contigs(0, 12)
    assert np.array_equal(np.concatenate([a, b]), ab)
    assert np.array_equal(np.concatenate([oa, ob[1:] + oa[-1]]), oab)
    assert set(np.unique(ab)) <= set(b"ACGT")


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np
import bench, refutil as R
from oracle import oracle as orc
D = bench.Dist("gloo", "cpu")
first, count = bench.shard_range(D.rank, 3)
flat, off = bench.make_contigs(first, count)
genes = 0
for k in range(count):   # the oracle stands in for the GPU path: this test is about sharding + reductions
    s = flat[off[k]:off[k + 1]][:4000].tobytes()
    d, gc, unk = orc.encode(s)
    genes += len(orc.find_genes_meta(d, gc / len(d), R.bins_blob())[0])
D.barrier()
tot_bp, tot_genes = D.reduce([int(off[-1]), genes], "sum")
mx, = D.reduce([float(D.rank + 1)], "max")
if D.rank == 0:
    print("RESULT", int(tot_bp), int(tot_genes), mx, D.world)
D.close()
'''


def test_two_rank_gloo_sharding_and_reduction(tmp_path):
    import bench
    from oracle import oracle as orc
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571", str(w), ROOT],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    flat, off = bench.make_contigs(0, 6)
    genes = 0
    for k in range(6):
        s = flat[off[k]:off[k + 1]][:4000].tobytes()
        d, gc, unk = orc.encode(s)
        genes += len(orc.find_genes_meta(d, gc / len(d), R.bins_blob())[0])
    assert int(line[1]) == int(off[-1]) and int(line[2]) == genes and float(line[3]) == 2.0 and int(line[4]) == 2


def test_train_argument_errors_mirror_the_reference():
    """lib.pyx:5522-5555: checked on the host, before any device work"""
    import warnings
    import pyrodigal_b200 as p
    with pytest.raises(RuntimeError, match="metagenomic"):
        p.GeneFinder(meta=True).train("A" * 30000)
    with pytest.raises(ValueError, match="not a valid translation table"):
        p.GeneFinder().train("A" * 30000, translation_table=7)
    with pytest.raises(ValueError, match="at least 20000"):
        p.GeneFinder().train("ACGT" * 100)
    with pytest.raises(ValueError, match="at least 20000"):
        p.GeneFinder().train(b"ACGT" * 100, b"ACGT" * 100)
    with pytest.raises(NotImplementedError):
        p.GeneFinder().train(p.Sequence("ACGT" * 6000), "ACGT")
    import torch
    if not torch.cuda.is_available():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with pytest.raises(RuntimeError, match="no CPU fallback"):   # enough sequence: reaches pgpu_create
                p.GeneFinder().train("ACGT" * 6000)
