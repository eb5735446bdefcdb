"""Secondary measurement (BASELINE.json configs[1]/[4] style): ONE long contig, single mode, one DP chain.
The chain is serial in the node index, so this is a latency test of the DP step, not a throughput test.
Times the GPU path (C ABI, host input) and, beside it, the reference on one host core (a single contig
cannot use more than one core in the reference).   python tools/bench_single_chain.py [Mbp ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as R  # noqa: E402
from pyrodigal_b200 import _capi  # noqa: E402

sizes = [float(x) for x in sys.argv[1:]] or [4.64, 16.0]
blob = R.bin_blob(20)
ctx = _capi.Context(0)
ctx.set_models(blob, 1)
opts = _capi.make_opts(meta=False, single_model=0)
try:
    ref = R.reference()
    ti = list(ref.METAGENOMIC_BINS)[20].training_info
    gf = ref.GeneFinder(ti)
except Exception:
    gf = None
for mbp in sizes:
    seq = R.synth(int(mbp * 1e6), 0.508, 2)
    a = np.ascontiguousarray(np.frombuffer(seq, np.uint8))
    off = np.array([0, len(a)], np.int64)
    for _ in range(2):
        r = ctx.find_genes_batch(a, off, opts)
    t0 = time.perf_counter()
    r = ctx.find_genes_batch(a, off, opts)
    dt = time.perf_counter() - t0
    st = r.stats
    line = {"Mbp": mbp, "gpu_s": round(dt, 4), "gpu_Mbp_s": round(mbp / dt, 1), "nodes": int(st["total_nodes"]),
            "genes": int(r.summary["n_genes"][0]), "dp_ms": round(st["ms_dp"], 2),
            "us_per_dp_step": round(st["ms_dp"] * 1e3 / max(1, st["dp_steps"]), 3)}
    if gf is not None and mbp <= 20:
        t0 = time.perf_counter()
        g = gf.find_genes(seq)
        line["reference_1core_s"] = round(time.perf_counter() - t0, 3)
        line["reference_genes"] = len(g)
    print(line)
