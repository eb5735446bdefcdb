"""pyrodigal_b200.distributed on the CPU: world_size 2 over gloo.  The engine on every rank is the CUDA-on-CPU emulation
of the product's kernels (tests/emu, test infrastructure); the gathered genes are checked against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
root, mode, out = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests")); sys.path.insert(0, os.path.join(root, "tests", "emu"))
import numpy as np
import torch.distributed as dist
import refutil as R
import emu_capi
from pyrodigal_b200 import distributed as PD
dist.init_process_group("gloo")
rank = dist.get_rank()
capi = emu_capi.load()
ctx = capi.Context(0)
ctx.set_models(R.bins_blob(), 50)
seqs = [R.synth(1500 + 700 * k, 0.3 + 0.03 * k, 40 + k, n_frac=0.002 if k == 5 else 0.0) for k in range(11)] + [b"", b"ACGT"]
arrs = [np.frombuffer(s, np.uint8) for s in seqs]
off = np.zeros(len(arrs) + 1, np.int64)
np.cumsum([len(a) for a in arrs], out=off[1:])
flat = np.ascontiguousarray(np.concatenate(arrs))
ran = []
def runner(shard, shard_off, shard_dev):
    ran.append(len(shard_off) - 1)
    return ctx.find_genes_batch(np.ascontiguousarray(shard), shard_off, capi.make_opts(meta=True))
if mode == "root" and rank != 0:
    res = PD.find_genes_sharded(None, None, input="root", runner=runner)
else:
    res = PD.find_genes_sharded(flat, off, input=mode, runner=runner)
assert (res is None) == (rank != 0)
assert 0 < ran[0] < len(seqs)          # both ranks got a share
if rank == 0:
    np.savez(out, summary=res.summary, genes=res.genes, gene_nodes=res.gene_nodes, gene_off=res.gene_off, share=ran[0])
dist.destroy_process_group()
'''


@pytest.mark.parametrize("mode", ["replicated", "root"])
def test_find_genes_sharded_two_ranks_gloo(tmp_path, mode):
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    out = tmp_path / "res.npz"
    port = 29581 if mode == "replicated" else 29583
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(w), ROOT, mode, str(out)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert p.returncode == 0, p.stderr[-3000:]
    z = np.load(out)
    seqs = [R.synth(1500 + 700 * k, 0.3 + 0.03 * k, 40 + k, n_frac=0.002 if k == 5 else 0.0) for k in range(11)] + [b"", b"ACGT"]
    assert len(z["summary"]) == len(seqs)
    total = 0
    for k, s in enumerate(seqs):
        d, gc, unk = orc.encode(s)
        genes, nodes, winner, _ = orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, R.bins_blob())
        assert int(z["summary"]["winner"][k]) == winner, k
        a, b = int(z["gene_off"][k]), int(z["gene_off"][k + 1])
        assert b - a == len(genes), k
        for f in ("begin", "end", "start_ndx", "stop_ndx"):
            assert np.array_equal(z["genes"][f][a:b], genes[f]), (k, f)
        if b > a:
            gn = z["gene_nodes"][a:b]
            for f in ("ndx", "strand", "cscore", "sscore", "rscore", "tscore", "uscore"):
                assert np.array_equal(gn[:, 0][f], nodes[genes["start_ndx"]][f]), (k, f)
                assert np.array_equal(gn[:, 1][f], nodes[genes["stop_ndx"]][f]), (k, f)
        total += b - a
    assert total > 10


def test_lpt_partition_balances_and_is_deterministic():
    from pyrodigal_b200 import distributed as PD
    rng = np.random.default_rng(1)
    lengths = rng.integers(1000, 100001, size=4000)
    gc = rng.uniform(.3, .7, size=4000)
    cost = PD.contig_cost(lengths, gc, model_gc=np.linspace(.25, .75, 50))
    for world in (1, 2, 4, 8):
        owner = PD.lpt_partition(cost, world)
        assert np.array_equal(owner, PD.lpt_partition(cost, world))
        loads = np.bincount(owner, weights=cost, minlength=world)
        assert loads.max() / loads.mean() < 1.01
    assert PD.lpt_partition(np.zeros(0), 4).shape == (0,)
    # a contiguous run of ids is passed on without copying
    flat = np.arange(100, dtype=np.uint8)
    off = np.array([0, 10, 30, 60, 100], np.int64)
    sh, so = PD._shard_view(flat, off, np.array([1, 2]))
    assert sh.base is flat and so.tolist() == [0, 20, 50]
    sh, so = PD._shard_view(flat, off, np.array([0, 3]))
    assert sh.tolist() == list(range(10)) + list(range(60, 100)) and so.tolist() == [0, 10, 50]
