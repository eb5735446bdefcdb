"""The product's CUDA kernels, run on the CPU.

tests/emu/build_emu.py compiles the unmodified sources of pyrodigal_b200/csrc/*.cu with g++ against a host stand-in of
the CUDA runtime and the SIMT execution model (tests/emu/cuda_emu/: device threads are fibers that switch at
__syncthreads / warp intrinsics), and the GPU parity tests of tests/test_gpu_parity.py are collected here a second time
with their `capi` / `ctx` fixtures bound to that library.  This is TEST INFRASTRUCTURE: it lets the build container
(no GPU) check the kernel sources themselves against the oracle and the goldens; the product package only ever loads
libpyrodigal_b200.so and still fails without a CUDA device.  What the emulation cannot show: anything that depends on
real concurrency between blocks or on the hardware's memory model (the GPU run remains the parity gate)."""
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

import refutil as R  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import test_gpu_parity as G  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import emu_capi
    return emu_capi.load()


@pytest.fixture(scope="module")
def ctx(capi):
    c = capi.Context(0)
    c.set_models(R.bins_blob(), 50)
    yield c
    c.close()


FULL = os.environ.get("PGPU_EMU_FULL") == "1"   # also the four heavy cases (about 8 minutes more)
_heavy_mark = pytest.mark.skipif(not FULL, reason="heavy under emulation: set PGPU_EMU_FULL=1")


def _clone(fn):
    """an independent copy of a test function: pytest marks are stored ON the function object, so marking the
    function imported from test_gpu_parity would also skip it in the GPU run (that happened in round 1)"""
    g = types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, fn.__closure__)
    g.__dict__.update(fn.__dict__)
    g.__kwdefaults__ = fn.__kwdefaults__
    g.pytestmark = list(getattr(fn, "pytestmark", []))
    return g


def heavy(fn):
    return _heavy_mark(_clone(fn))

# the operator-level and end-to-end parity tests of the GPU suite, unchanged (the Mbp-sized inputs stay GPU-only)
test_extract_nodes = G.test_extract_nodes
test_extract_nodes_masked = G.test_extract_nodes_masked
test_score_nodes = G.test_score_nodes
test_score_connections_golden = G.test_score_connections_golden
test_training_dp_vs_oracle = G.test_training_dp_vs_oracle
test_compute_skippable = G.test_compute_skippable
test_skippable_plugin_signature = G.test_skippable_plugin_signature
test_result_nodes_in_reference_struct_layout = G.test_result_nodes_in_reference_struct_layout
test_find_genes_meta_golden = G.test_find_genes_meta_golden
test_find_genes_single_golden = G.test_find_genes_single_golden
test_find_genes_batch_vs_oracle = G.test_find_genes_batch_vs_oracle
test_find_genes_options_vs_oracle = G.test_find_genes_options_vs_oracle
test_dp_kernel_variants_single_golden = G.test_dp_kernel_variants_single_golden
test_resident_batch_matches_host_batch = G.test_resident_batch_matches_host_batch
test_two_lane_host_batches_match_single_stream = heavy(G.test_two_lane_host_batches_match_single_stream)
test_dp_model_lane_kernel_equals_per_chain_kernel = heavy(G.test_dp_model_lane_kernel_equals_per_chain_kernel)
test_sub_batching_matches_single_batch = heavy(G.test_sub_batching_matches_single_batch)
test_gene_only_final_pass_equals_full_final_pass = heavy(G.test_gene_only_final_pass_equals_full_final_pass)


def _batch(seqs):
    arrs = [np.frombuffer(s, np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    return np.ascontiguousarray(np.concatenate(arrs)), off


@pytest.mark.parametrize("closed", [False, True])
def test_gene_only_final_pass_small(ctx, capi, closed):
    """the genes-only final scoring pass (meta mode without node arrays) against the full pass, on a small batch"""
    seqs = [b"", R.synth(95, .5, 3)] + [R.synth(2000 + 900 * k, .3 + .03 * k, 7000 + k, n_frac=0.002 if k % 4 == 0 else 0.0)
                                         for k in range(14)]
    flat, off = _batch(seqs)
    full = ctx.find_genes_batch(flat, off, capi.make_opts(meta=True, closed=closed, want_nodes=True))
    lean = ctx.find_genes_batch(flat, off, capi.make_opts(meta=True, closed=closed, want_nodes=False))
    assert len(lean.genes) > 40
    assert lean.genes.tobytes() == full.genes.tobytes() and lean.gene_nodes.tobytes() == full.gene_nodes.tobytes()


def test_dp_model_lane_kernel_small(capi, monkeypatch):
    """k_dp_ml (one lane per model) against k_dp_dq (one warp per chain) through the library's own self-check"""
    monkeypatch.setenv("PGPU_DP_VERIFY", "1")
    c = capi.Context(0)
    c.set_models(R.bins_blob(), 50)
    flat, off = _batch([R.synth(6000 + 2500 * k, .3 + .05 * k, 300 + k) for k in range(8)])
    res = c.find_genes_batch(flat, off, capi.make_opts(meta=True))   # PGPU_DP_VERIFY fails the call on any difference
    assert int(res.summary["n_genes"].sum()) > 20
    c.close()


def test_two_lanes_small(capi, monkeypatch):
    """PGPU_LANES=2: two worker threads / sub-batches, results stitched in order (the emulation runs one kernel at a time)"""
    seqs = [R.synth(3000 + 1700 * k, .32 + .03 * k, 12000 + k) for k in range(11)] + [b"", R.synth(50, .5, 1)]
    flat, off = _batch(seqs)
    o = capi.make_opts(meta=True, want_nodes=True)
    monkeypatch.setenv("PGPU_LANES", "1")
    c1 = capi.Context(0)
    c1.set_models(R.bins_blob(), 50)
    r1 = c1.find_genes_batch(flat, off, o)
    monkeypatch.setenv("PGPU_LANES", "2")
    monkeypatch.setenv("PGPU_LANE_MIN_BP", "0")
    c2 = capi.Context(0)
    c2.set_models(R.bins_blob(), 50)
    for rep in range(2):
        r2 = c2.find_genes_batch(flat, off, o)
        assert r2.stats["kernel_launches"] > 1.5 * r1.stats["kernel_launches"]   # really ran as two sub-batches
        assert r1.summary.tobytes() == r2.summary.tobytes() and np.array_equal(r1.gene_off, r2.gene_off)
        assert r1.genes.tobytes() == r2.genes.tobytes() and r1.gene_nodes.tobytes() == r2.gene_nodes.tobytes()
        for k in (0, 5, 6, 10):
            assert r1.nodes(k).tobytes() == r2.nodes(k).tobytes()
        for key in ("n_contigs", "total_bp", "total_nodes", "total_chain_nodes", "n_chains", "total_genes", "pairs", "dp_steps"):
            assert r1.stats[key] == r2.stats[key], key
    c1.close(); c2.close()


# ---- the Python mirror on top of the emulated library (pyrodigal_b200.lib with its ctypes binding swapped) ----
@pytest.fixture()
def emulated_lib(capi, monkeypatch):
    import pyrodigal_b200.lib as L
    monkeypatch.setattr(L, "_capi", capi)
    saved = dict(L._ctx_cache)
    L._ctx_cache.clear()
    yield L
    for c in L._ctx_cache.values():
        c.close()
    L._ctx_cache.clear()
    L._ctx_cache.update(saved)


def test_nodes_operator_api(emulated_lib):
    import nodes_api_cases
    nodes_api_cases.run_cases(emulated_lib)


def test_sequence_operators(emulated_lib):
    import nodes_api_cases
    nodes_api_cases.run_sequence_operator_cases(emulated_lib)


@pytest.mark.parametrize("name", ["s20000_min", "s30k_tt4_forced", "s40k_N_mask"] +
                         (["s25k_closed_hi", "s60k_nonsd", "srr_contig", "ref100k_open", "ref100k_closed"] if FULL else []))
def test_train_golden(emulated_lib, name, tmp_path):
    import test_gpu_train
    test_gpu_train.test_train_golden(name, tmp_path)


def test_train_multi_contig_linker(emulated_lib):
    import test_gpu_train
    test_gpu_train.test_train_multi_contig_linker()


def test_python_api_drop_in(emulated_lib):
    G.test_python_api_drop_in()


def test_command_line_end_to_end(emulated_lib, tmp_path):
    """FASTA ingest -> emulated kernels -> writers, against what the reference's command line wrote (meta and single
    mode, GFF / GenBank / FASTA / score tables, training file)"""
    import test_gpu_cli
    for k, name in enumerate(test_gpu_cli.W["names"]):
        d = tmp_path / str(k)
        d.mkdir()
        test_gpu_cli.test_cli_matches_reference_cli(name, d)
    test_gpu_cli.test_cli_rejects_training_file_in_meta_mode(tmp_path)


def test_randomised_batches_against_the_oracle(monkeypatch):
    """a short seeded run of tests/emu/fuzz.py and fuzz_train.py (odd contigs: Ns, repeats, tiny / extreme-GC contigs,
    IUPAC letters; random end mode, masks, translation tables, start weights)"""
    import fuzz
    import fuzz_train
    monkeypatch.setattr(sys, "argv", ["fuzz", "12", "2026"])
    assert fuzz.main() == 0
    monkeypatch.setattr(sys, "argv", ["fuzz_train", "10", "2026"])
    assert fuzz_train.main() == 0


test_coding_score_lane_groups = G.test_coding_score_lane_groups
test_coding_score_shared_memory_tables = G.test_coding_score_shared_memory_tables
test_coding_score_many_plan_entries = G.test_coding_score_many_plan_entries
test_dp_model_lane_kernel_gc_sweep = G.test_dp_model_lane_kernel_gc_sweep
