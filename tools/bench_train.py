"""Secondary measurement, BASELINE.json configs[1] ("E. coli K-12 MG1655 (4.6 Mbp), single mode (train+find_genes),
1xB200"): the real genome is not available offline, so the stand-in of SURVEY.md 8(d) is used -- an iid
4 641 652-bp sequence with gc = 0.508, seed 2.  Times GeneFinder.train + find_genes on the GPU (C ABI, host
input) and, beside it, the unmodified reference on one host core (one sequence cannot use more than one core in
the reference), and checks that both produce the same training struct and genes.
    python tools/bench_train.py [Mbp ...]     ->  one JSON line per size"""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refutil as R  # noqa: E402
import pyrodigal_b200  # noqa: E402

warnings.simplefilter("ignore")
sizes = [float(x) for x in sys.argv[1:]] or [4.641652]
try:
    ref = R.reference()
except Exception:
    ref = None
for mbp in sizes:
    seq = R.synth(int(round(mbp * 1e6)), 0.508, 2)
    line = {"workload": f"cfg2 stand-in: iid {len(seq)} bp gc .508 seed 2, single mode", "Mbp": mbp}
    for forced in (False, True):
        tag = "nonsd" if forced else "auto"
        gf = pyrodigal_b200.GeneFinder()
        gf.train(seq, force_nonsd=forced)          # warm-up (context creation, pool growth)
        t0 = time.perf_counter()
        ti = gf.train(seq, force_nonsd=forced)
        t_train = time.perf_counter() - t0
        st = gf.last_train_stats
        gf.find_genes_many([seq])
        t0 = time.perf_counter()
        g = gf.find_genes_many([seq])[0]
        t_find = time.perf_counter() - t0
        line[tag] = {"uses_sd": bool(ti.uses_sd), "gpu_train_s": round(t_train, 4), "gpu_find_genes_s": round(t_find, 4),
                     "gpu_Mbp_s_train_plus_find": round(mbp / (t_train + t_find), 2), "genes": len(g),
                     "train_first_gene_set_ms": round(st["ms_dp"], 2), "train_start_rounds_ms": round(st["ms_score"], 2),
                     "train_kernel_launches": int(st["kernel_launches"]), "nodes": int(st["total_nodes"])}
        if ref is not None and mbp <= 20:
            rg = ref.GeneFinder()
            t0 = time.perf_counter()
            rti = rg.train(seq, force_nonsd=forced)
            r_train = time.perf_counter() - t0
            t0 = time.perf_counter()
            rgenes = rg.find_genes(seq)
            r_find = time.perf_counter() - t0
            line[tag].update({"reference_1core_train_s": round(r_train, 3), "reference_1core_find_genes_s": round(r_find, 3),
                              "reference_Mbp_s_train_plus_find": round(mbp / (r_train + r_find), 2),
                              "training_struct_identical": bytes(memoryview(rti)) == bytes(ti),
                              "genes_identical": [(x.begin, x.end, x.strand) for x in rgenes] == [(x.begin, x.end, x.strand) for x in g]})
        if ti.uses_sd is False and not forced:
            line["nonsd"] = line[tag]   # the heuristic already chose the non-SD path: nothing new to time
            break
    print(json.dumps(line))
