#!/usr/bin/env python
"""bench.py -- Mbp/s of find_genes on synthetic contig batches, 1..8 B200.

    python bench.py --gpus N --steps K --warmup W [--config NAME]     # our CUDA path
    python bench.py --impl reference --gpus N --steps K ... [--config NAME]   # the reference's own CPU path

Configurations (BASELINE.json `configs`, generators of SURVEY.md 8d):
    cfg4 (default)  the 100 000-contig synthetic metagenome (lengths uniform 1-100 kbp, GC uniform 0.30-0.70, iid
                    bases, seeds 1 000 000 + k), meta mode, sharded over the GPUs of one box: every rank owns 12 500
                    contigs (~630 Mbp), so per-GPU work is fixed as N grows ("weak" scaling) and N = 8 is exactly the
                    100k-contig / ~5 Gbp batch
    cfg4-full       all 100 000 contigs (5.05 Gbp) on every GPU through the library's sub-batching (the north-star target
                    configuration on ONE GPU; N > 1 = replicas)
    cfg3            1 000 contigs, same distributions (generator seed 3, contig seeds 10 000 + k), meta mode, one batched call
    cfg2            E. coli stand-in (iid 4 641 652 bp, gc 0.508, seed 2), single mode: a step = GeneFinder.train + find_genes
    cfg5 / cfg5-tt4 one 50 Mbp chromosome (gc 0.50, seed 5), single mode with built-in bin 20 / bin 0 (translation table 4:
                    giant-ORF windows): one DP chain, a latency test of the DP step
A "step" is one pass of the whole hot path (encode -> add_nodes -> score_nodes -> overlapping starts -> connection DP for
every (contig, model) chain -> winner / traceback / genes -> final re-score) over the rank's input.

Printed JSON line (rank 0): see the task contract; additionally `roofline` (the DP kernel of the configuration),
`cpu_baseline`, `e2e` (C ABI, host buffers; plus the Python surface and a pageable-input figure), `parity_checked`,
`clocks`, `gpu_launches`, `phases`.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONTIGS_PER_GPU = 12500
DP_BYTES_FINAL = 72   # algorithmic HBM bytes of one DP step, final = 1
DP_BYTES_TRAIN = 64   # ... of one step of the training DP (final = 0)


def synth_into(out, gc, seed):
    """iid bases, P(A)=P(T)=(1-gc)/2, P(C)=P(G)=gc/2, through a single uniform draw per base (in chunks)"""
    g = np.random.default_rng(seed)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    a = (1 - gc) / 2
    for s in range(0, len(out), 1 << 24):
        u = g.random(min(1 << 24, len(out) - s), dtype=np.float32)
        code = (u >= a).astype(np.uint8) + (u >= a + gc / 2) + (u >= a + gc)
        out[s:s + len(u)] = lut[code]


def config_table(seed=4, universe=100_000):
    """(lengths, gcs) of every contig of a configuration, drawn once so that shards are disjoint subsets of it"""
    rng = np.random.default_rng(seed)
    lengths = rng.integers(1_000, 100_001, size=universe)
    gcs = rng.uniform(0.30, 0.70, size=universe)
    return lengths, gcs


def make_contigs(first, count, seed=4, universe=100_000, contig_seed0=1_000_000, ids=None):
    """contigs [first, first+count) (or the contigs `ids`) of a (lengths uniform 1-100 kbp, GC uniform .30-.70)
    configuration: (flat uint8 ASCII array, int64 offsets).  Defaults = cfg4; cfg3 = (seed 3, universe 1000, contig
    seeds 10 000 + k)"""
    lengths, gcs = config_table(seed, universe)
    idx = (np.arange(first, first + count) if ids is None else np.asarray(ids)) % universe
    count = len(idx)
    lens = lengths[idx]
    offsets = np.zeros(count + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    flat = np.empty(int(offsets[-1]), dtype=np.uint8)
    for k, i in enumerate(idx):
        synth_into(flat[offsets[k]:offsets[k + 1]], gcs[i], contig_seed0 + int(i))
    return flat, offsets


def shard_range(rank, contigs_per_gpu):
    """contigs owned by `rank` under the by-index partition: a disjoint slice of the 100k-contig configuration"""
    return rank * contigs_per_gpu, contigs_per_gpu


def shard_ids(rank, world, contigs_per_gpu):
    """contigs owned by `rank` when `world` GPUs share the first world x contigs_per_gpu contigs of cfg4 (weak scaling:
    the job grows with N, N = 8 is the whole 100k-contig batch): greedy longest-processing-time partition over the
    estimated cost per contig (pyrodigal_b200.distributed).  One rank = the by-index slice."""
    if world == 1:
        return np.arange(contigs_per_gpu)
    from pyrodigal_b200 import distributed as PD
    import pyrodigal_b200
    n = min(world * contigs_per_gpu, 100_000)
    lengths, gcs = config_table()
    model_gc = [b.training_info.gc for b in pyrodigal_b200.METAGENOMIC_BINS]
    owner = PD.lpt_partition(PD.contig_cost(lengths[:n], gcs[:n], model_gc), world)
    return np.flatnonzero(owner == rank)


class Dist:
    """torch.distributed plumbing of the bench (barrier, max over ranks, sums); nccl on GPUs, gloo in CPU tests"""

    def __init__(self, backend=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.device = device if device is not None else "cpu"
        if self.world > 1 and not dist.is_initialized():
            kw = {}
            if backend == "nccl":
                # NCCL kernels on high-priority streams: the gather of one step's gene records runs while the next
                # step's kernels fill the GPU, and must not queue behind their pending CTAs
                os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
                kw["device_id"] = torch.device(self.device)
            dist.init_process_group(backend, **kw)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        if str(self.device).startswith("cuda"):
            self.torch.cuda.synchronize()

    def reduce(self, values, op="sum"):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def close(self):
        if self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled every 250 ms DURING the timed region (B200_PROFILING.md's clocks line).
    Read through NVML in this process (the library nvidia-smi itself queries): starting an `nvidia-smi` process per
    sample initialises the driver API every time and was seen to delay the launches of the step being timed by tens
    of ms.  Falls back to ONE `nvidia-smi -lms 200` process started before the region when pynvml is missing."""

    REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml, self.handle, self.proc, self.sm_max = None, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA device through its UUID when CUDA_VISIBLE_DEVICES is set
            try:
                import torch
                uuid = torch.cuda.get_device_properties(index).uuid
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            # the first query of each kind initialises driver state: keep that outside the timed region
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def start(self):
        if self.nvml is None:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                              "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                time.sleep(1.0)  # its start-up stays outside the timed region
            except Exception:
                self.proc = None
            return
        super().start()

    def run(self):
        n = self.nvml
        masks = (n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                 n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap)
        while not self._stop_evt.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append((sm, self.sm_max, [name for name, m in zip(self.REASONS, masks) if r & m]))
            except Exception:
                pass
            self._stop_evt.wait(0.25)

    def stop(self):
        if self.nvml is None:
            if self.proc is not None:
                self.proc.terminate()
                try:
                    out = self.proc.communicate(timeout=3)[0]
                except Exception:
                    out = ""
                for line in out.strip().splitlines():
                    f = [x.strip() for x in line.split(",")]
                    if len(f) >= 6 and f[0].replace(".", "").isdigit():
                        self.rows.append((float(f[0]), float(f[1]),
                                          [nm for nm, v in zip(self.REASONS, f[2:6]) if v.lower().startswith("active")]))
        else:
            self._stop_evt.set()
            self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({x for r in self.rows for x in r[2]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi -lms 200"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def effective_cores():
    """host cores this process may actually use: min(os.cpu_count, affinity mask, cgroup CPU quota)"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(round(int(quota) / int(period)))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, int(round(q / p))))
        except Exception:
            pass
    return n


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
META_CONFIGS = ("cfg4", "cfg4-full", "cfg3")
SINGLE_CONFIGS = ("cfg2", "cfg5", "cfg5-tt4")


def workload(config, rank, contigs, world=1):
    """-> dict(flat, offsets, meta, label, ...) of the rank's input for `config`"""
    import refutil as R
    if config == "cfg4":
        ids = shard_ids(rank, world, contigs)
        flat, off = make_contigs(0, 0, ids=ids)
        return {"flat": flat, "offsets": off, "meta": True, "ids": ids, "n_total": min(world * contigs, 100_000),
                "label": "cfg4: 100k-contig synthetic metagenome (1-100 kbp, GC 0.30-0.70), meta mode, "
                         f"{contigs} contigs per GPU, contig-sharded"}
    if config == "cfg4-full":
        flat, off = make_contigs(0, 100_000)
        return {"flat": flat, "offsets": off, "meta": True,
                "label": "cfg4-full: all 100 000 contigs of the synthetic metagenome (5.05 Gbp) on ONE GPU, meta mode, "
                         "sub-batched by the library"}
    if config == "cfg3":
        flat, off = make_contigs(0, 1000, seed=3, universe=1000, contig_seed0=10_000)
        return {"flat": flat, "offsets": off, "meta": True,
                "label": "cfg3: 1 000 synthetic contigs 1-100 kbp (GC 0.30-0.70), meta mode, one batched call"}
    if config == "cfg2":
        seq = np.frombuffer(R.synth(4_641_652, 0.508, 2), np.uint8)
        return {"flat": np.ascontiguousarray(seq), "offsets": np.array([0, len(seq)], np.int64), "meta": False, "train": True,
                "label": "cfg2: E. coli K-12 stand-in (iid 4 641 652 bp, gc 0.508, seed 2; the genome itself is not available "
                         "offline), single mode, step = GeneFinder.train + find_genes"}
    if config in ("cfg5", "cfg5-tt4"):
        seq = np.frombuffer(R.synth(50_000_000, 0.5, 5), np.uint8)
        b = 20 if config == "cfg5" else 0
        return {"flat": np.ascontiguousarray(seq), "offsets": np.array([0, len(seq)], np.int64), "meta": False, "train": False,
                "bin": b,
                "label": f"{config}: one 50 Mbp synthetic chromosome (gc 0.50, seed 5), single mode with built-in bin {b}"
                         + (" (translation table 4: giant-ORF DP windows)" if b == 0 else "") + ", one DP chain"}
    raise SystemExit(f"unknown --config {config}")


# ---------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref = the unmodified reference build; C oracle port otherwise)
# ---------------------------------------------------------------------------------------------------
_REF_GF = None


def _genes_of(genes):
    return [(g.begin, g.end, g.strand) for g in genes]


def _ref_init(ref_dir):
    """process-pool worker initialiser: import the unmodified reference and build one GeneFinder"""
    global _REF_GF
    sys.path.insert(0, ref_dir)
    import pyrodigal
    _REF_GF = pyrodigal.GeneFinder(meta=True)


def _ref_work(seq):
    return _genes_of(_REF_GF.find_genes(seq))


def cpu_meta_runners():
    """[(label, kind, fn(list_of_bytes) -> per-contig [(begin, end, strand)], close)] -- the reference's own CPU
    implementation with every host core: its documented thread-pool recipe (docs/guide/parallel.rst:24-41, cli.py:286-300)
    and the process pool its CLI also offers (cli.py:292-293).  Falls back to the C oracle port when oracle/_ref is absent."""
    cores = effective_cores()  # cgroup quota aware (os.cpu_count() reports the whole host)
    from multiprocessing.pool import ThreadPool
    import multiprocessing as mp
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    out = []
    try:
        if not os.path.exists(os.path.join(ref_dir, "pyrodigal")):
            raise ImportError("oracle/_ref not present")
        sys.path.insert(0, ref_dir)
        import pyrodigal  # the unmodified reference
        gf = pyrodigal.GeneFinder(meta=True)  # backend="detect" (SSE2/AVX2 SIMD skip filter)
        tp = ThreadPool(cores)
        out.append((f"pyrodigal {pyrodigal.__version__} GeneFinder(meta=True) ThreadPool({cores})", "reference",
                    lambda seqs: tp.map(lambda s: _genes_of(gf.find_genes(s)), seqs), tp.close))
        try:
            pp = mp.get_context("spawn").Pool(cores, initializer=_ref_init, initargs=(ref_dir,))
            out.append((f"pyrodigal {pyrodigal.__version__} GeneFinder(meta=True) multiprocessing.Pool({cores})", "reference",
                        lambda seqs: pp.map(_ref_work, seqs, chunksize=1), pp.terminate))
        except Exception:
            pass
    except Exception as e:
        from oracle import oracle as orc
        import refutil as R
        blob = R.bins_blob()
        tp = ThreadPool(cores)

        def one(s):
            d, gc, unk = orc.encode(s)
            genes, nodes, _, _ = orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, blob)
            return [(int(a["begin"]), int(a["end"]), int(nodes[a["start_ndx"]]["strand"])) for a in genes]
        out.append((f"C oracle port ThreadPool({cores}) [{type(e).__name__}: {e}]", "port", lambda seqs: tp.map(one, seqs), tp.close))
    return out, cores


def time_cpu_meta(flat, offsets, steps, warmup, max_contigs):
    """times every available CPU configuration on the same bounded sample (the first contigs of the input) and reports
    the fastest; keeps the genes it found so that the GPU result can be checked against them"""
    runners, cores = cpu_meta_runners()
    n = min(len(offsets) - 1, max_contigs)
    # longest contigs first: the pools then finish without a long tail
    order = sorted(range(n), key=lambda k: int(offsets[k + 1] - offsets[k]), reverse=True)
    seqs = [flat[offsets[k]:offsets[k + 1]].tobytes() for k in order]
    bp = int(offsets[n] - offsets[0])
    best, tried = None, []
    for label, kind, run, close in runners:
        try:
            for _ in range(max(1, warmup)):
                run(seqs[: max(cores, n // 8)])
            t0 = time.perf_counter()
            got = None
            for _ in range(steps):
                got = run(seqs)
            dt = (time.perf_counter() - t0) / steps
            tried.append(f"{label}: {bp / dt / 1e6:.1f} Mbp/s")
            if best is None or dt < best["s_per_step"]:
                genes = [None] * n
                for k, g in zip(order, got):
                    genes[k] = g
                best = {"value": bp / dt / 1e6, "unit": "Mbp/s", "cores": cores, "kind": kind, "s_per_step": dt,
                        "genes": genes, "label": label}
        finally:
            try:
                close()
            except Exception:
                pass
    best["sample"] = (f"first {n} contigs of the rank-0 input ({bp / 1e6:.1f} Mbp); fastest of: " + "; ".join(tried))
    return best


def time_cpu_single(w, steps, sample_bp):
    """cfg2 / cfg5 on the CPU: one sequence cannot use more than one core in the reference (a single DP chain), so this
    is a one-core figure; cfg5 is timed on a bounded prefix of the chromosome"""
    import refutil as R
    seq = w["flat"][:sample_bp].tobytes() if sample_bp and sample_bp < len(w["flat"]) else w["flat"].tobytes()
    note = f"{len(seq) / 1e6:.2f} Mbp" + (" (prefix of the chromosome)" if len(seq) < len(w["flat"]) else " (whole input)")
    try:
        ref = R.reference()
        if w.get("train"):
            def run():
                gf = ref.GeneFinder()
                gf.train(seq)
                return _genes_of(gf.find_genes(seq))
            label = f"pyrodigal {ref.__version__} GeneFinder().train + find_genes, 1 core"
        else:
            gf = ref.GeneFinder(list(ref.METAGENOMIC_BINS)[w["bin"]].training_info)
            run = lambda: _genes_of(gf.find_genes(seq))
            label = f"pyrodigal {ref.__version__} GeneFinder(bin {w['bin']}).find_genes, 1 core"
        kind = "reference"
    except Exception as e:
        from oracle import oracle as orc
        if w.get("train"):
            raise RuntimeError(f"cfg2 needs the reference build for its CPU leg ({e})")
        blob = R.bin_blob(w["bin"])

        def run():
            d, gc, unk = orc.encode(seq)
            genes, nodes, _ = orc.find_genes_single(d, blob)
            return [(int(a["begin"]), int(a["end"]), int(nodes[a["start_ndx"]]["strand"])) for a in genes]
        label, kind = f"C oracle port find_genes_single(bin {w['bin']}), 1 core [{type(e).__name__}]", "port"
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        genes = None
        for _ in range(steps):
            genes = run()
        dt = (time.perf_counter() - t0) / steps
    return {"value": len(seq) / dt / 1e6, "unit": "Mbp/s", "cores": 1, "kind": kind, "s_per_step": dt, "genes": [genes],
            "label": label, "sample": f"{label}: {note}", "sample_bp": len(seq)}


def gpu_genes(res, k):
    a, b = int(res.gene_off[k]), int(res.gene_off[k + 1])
    g, gn = res.genes[a:b], res.gene_nodes[a:b]
    return list(zip(g["begin"].tolist(), g["end"].tolist(), gn[:, 0]["strand"].tolist()))


def check_parity(res, cpu_genes):
    """GPU genes of the CPU sample's contigs == the CPU implementation's (begin, end, strand); -> (n_checked, n_bad)"""
    bad = 0
    for k, want in enumerate(cpu_genes):
        if want is not None and gpu_genes(res, k) != [tuple(x) for x in want]:
            bad += 1
    return sum(1 for g in cpu_genes if g is not None), bad


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg4", choices=list(META_CONFIGS + SINGLE_CONFIGS))
    ap.add_argument("--contigs", type=int, default=CONTIGS_PER_GPU, help="cfg4: contigs per GPU (default = the 12 500-contig shard)")
    ap.add_argument("--cpu-contigs", type=int, default=0, help="contigs in the CPU sample (0 = 32 per core)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)
    sharded = args.config == "cfg4"
    metric = "Mbp/s find_genes (meta mode)" if args.config in META_CONFIGS else \
             ("Mbp/s train + find_genes (single mode)" if args.config == "cfg2" else "Mbp/s find_genes (single mode)")

    # nominal input size of the configuration (decides the L2 policy, identically in both arms)
    nominal_bp = {"cfg4": args.contigs * 50_500, "cfg4-full": 5_050_000_000, "cfg3": 50_500_000, "cfg2": 4_641_652,
                  "cfg5": 50_000_000, "cfg5-tt4": 50_000_000}[args.config]
    small = nominal_bp <= (200 << 20)

    def base_line(w):
        cfg = {"workload": w["label"], "name": args.config,
               "parallelism": (f"contig-shard x{n_gpus} (LPT by estimated cost; gene records gathered on rank 0 inside e2e)"
                               if sharded else f"replicas x{n_gpus}"),
               "l2_policy": "L2 flushed between timed steps (256 MB device memset)" if small
               else "inputs larger than L2 (per-step working set >> 126 MB)"}
        if sharded:
            cfg["contigs_per_gpu"] = args.contigs
        return {"metric": metric, "unit": "Mbp/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg}

    # ------------------------------------------------------------------ reference arm (CPU) --------
    if args.impl == "reference":
        if rank != 0:
            return
        cores = effective_cores()
        if args.config in META_CONFIGS:
            n_ref = args.cpu_contigs or 32 * cores
            if args.config == "cfg3":
                w = workload("cfg3", 0, 0)
            else:   # a bounded sample of the rank-0 shard: the first contigs of the configuration
                flat, off = make_contigs(0, min(max(n_ref, 64), 100_000))
                w = {"flat": flat, "offsets": off, "meta": True}
                w["label"] = ("cfg4: 100k-contig synthetic metagenome (1-100 kbp, GC 0.30-0.70), meta mode, "
                              f"{args.contigs} contigs per GPU, contig-sharded") if args.config == "cfg4" else \
                             ("cfg4-full: all 100 000 contigs of the synthetic metagenome (5.05 Gbp) on ONE GPU, meta mode, "
                              "sub-batched by the library")
            cb = time_cpu_meta(w["flat"], w["offsets"], args.steps, 1, n_ref)
        else:
            w = workload(args.config, 0, 0)
            cb = time_cpu_single(w, args.steps, 0 if args.config == "cfg2" else 5_000_000)
        line = base_line(w)
        line.update({"impl": "reference", "value": cb["value"], "ms_per_step": cb["s_per_step"] * 1e3,
                     "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                     "e2e": {"value": cb["value"], "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm --------------------
    import torch
    from pyrodigal_b200 import _capi
    import pyrodigal_b200
    import refutil as R

    torch.cuda.set_device(local_rank)
    D = Dist("nccl", f"cuda:{local_rank}")
    w = workload(args.config, rank, args.contigs, world)
    flat, offsets = w["flat"], w["offsets"]
    bp = int(offsets[-1])
    # pinned host staging of the step's input (the e2e leg copies from here every step)
    pinned = torch.empty(bp, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = flat
    host = pinned.numpy()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}") if small else None

    def flush_l2():
        if flush_buf is not None:
            flush_buf.zero_()          # 256 MB > the 126 MB L2; on torch's stream, so synchronise before the step
            torch.cuda.synchronize()

    ctx = _capi.Context(local_rank)
    tctx = None
    if w["meta"]:
        ctx.set_models(R.bins_blob(), 50)
        opts = _capi.make_opts(meta=True)
    else:
        opts = _capi.make_opts(meta=False, single_model=0)
        if w.get("train"):
            tctx = _capi.Context(local_rank)
        else:
            ctx.set_models(R.bin_blob(w["bin"]), 1)

    class Gatherer:
        """N > 1: the gene records of every step are gathered on rank 0 (pyrodigal_b200.distributed.gather_result: counts by
        all_gather, records over NCCL) by a helper thread, so that the gather of step k overlaps the kernels of step k+1;
        the last gather is waited for INSIDE the timed region."""
        def __init__(self):
            from pyrodigal_b200 import distributed as PD
            self.PD, self.t, self.out, self.err = PD, None, None, None

        def _run(self, res):
            try:
                torch.cuda.set_device(local_rank)
                self.out = self.PD.gather_result(res, w["ids"], w["n_total"], device=local_rank)
            except Exception as e:   # re-raised by wait()
                self.err = e

        def submit(self, res):
            self.wait()
            self.t = threading.Thread(target=self._run, args=(res,))
            self.t.start()

        def wait(self):
            if self.t is not None:
                self.t.join()
                self.t = None
            if self.err is not None:
                raise self.err
            return self.out

    def timed(fn, steps, per_step=None, gather=None):
        """K steps bracketed by barrier + synchronize; CUDA events on the library's stream; max over ranks.  Small
        inputs: L2 is flushed before every step, and the step times (CUDA events per step) are summed instead."""
        gc.collect()
        gc.disable()   # no cyclic collection inside the timed region (re-enabled below): ordinary benchmarking hygiene
        D.barrier()
        tot_ms = 0.0
        if not small:
            ctx.timer_start()
        t0 = time.perf_counter()
        wall = 0.0
        last = None
        for _ in range(steps):
            if gather is None:
                last = None  # release the previous result first: its pinned buffer is reused by the next step
            if small:
                flush_l2()
                ctx.timer_start()
            t1 = time.perf_counter()
            last = fn()
            if gather is not None:
                gather.submit(last)
            dt = (time.perf_counter() - t1) * 1e3
            if small:
                tot_ms += ctx.timer_stop()
                wall += dt
            if per_step is not None:
                per_step.append(dict(last.stats, wall_ms=dt))
        if gather is not None:
            gather.wait()
        if not small:
            tot_ms = ctx.timer_stop()
            wall = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        gc.enable()
        tot_ms, wall = D.reduce([tot_ms, wall], "max")
        D.barrier()
        return tot_ms, wall, last

    class TrainFind:
        """cfg2 step through the C ABI: pgpu_train on the training context, then pgpu_set_models + find_genes"""
        def __init__(self, resident):
            self.resident = resident
            self.batch = ctx.upload(host, offsets) if resident else None

        def __call__(self):
            blob, tst = tctx.train(host, opts)
            ctx.set_models(blob, 1)
            r = self.batch.run(opts) if self.resident else ctx.find_genes_batch(host, offsets, opts)
            for k in ("ms_total_device", "ms_dp", "ms_score", "kernel_launches"):
                r.stats["train_" + k] = tst[k]
            r.stats["kernel_launches"] += tst["kernel_launches"]
            return r

    # ---- device-resident leg ("value") ----
    if w.get("train"):
        step_resident = TrainFind(True)
        step_host = TrainFind(False)
        batch = step_resident.batch
    else:
        batch = ctx.upload(host, offsets)
        step_resident = lambda: batch.run(opts)
        step_host = lambda: ctx.find_genes_batch(host, offsets, opts)
    r = None
    for _ in range(args.warmup):
        r = None  # at most two results alive at a time: their pinned buffers are recycled by the library
        r = step_resident()
    r = None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if rank == 0:
        sampler.start()
    step_stats = []
    ms, wall_ms, res = timed(step_resident, args.steps, step_stats)
    clocks = sampler.stop() if rank == 0 else None
    stats = res.stats
    genes_rank = int(res.summary["n_genes"].sum())
    if w.get("train"):
        # cfg2: the training call uploads its input itself (there is no resident-input training entry point); `value`
        # is therefore the sum of the device times of the two calls (CUDA events around their kernels, input copy excluded)
        ms = sum(t["ms_total_device"] + t["train_ms_total_device"] for t in step_stats)
        ms = D.reduce([ms], "max")[0]

    # ---- end-to-end leg: C ABI call with host buffers, H2D + D2H inside ----
    res = None
    for _ in range(max(2, args.warmup)):
        r = None
        r = step_host()
    r = None
    e2e_steps = []
    gatherer = Gatherer() if (sharded and world > 1 and os.environ.get("BENCH_NO_GATHER") != "1") else None
    if gatherer is not None:   # untimed pipelined steps: NCCL connections, the rotation of page-locked result / receive buffers
        for _ in range(4):
            gatherer.submit(step_host())
        gatherer.wait()
    ms_e2e, wall_e2e, res2 = timed(step_host, args.steps, e2e_steps, gatherer)
    st2 = res2.stats
    gathered = gatherer.wait() if gatherer is not None else None
    if w.get("train") or gatherer is not None:
        ms_e2e = wall_e2e   # host work between / after the library calls: the wall clock is the end-to-end time
    # the same call from pageable (ordinary numpy) memory, as a Python caller holding `bytes` would make it
    res2 = None
    pe_steps = max(2, min(3, args.steps))
    step_page = (lambda: ctx.find_genes_batch(flat, offsets, opts)) if not w.get("train") else None
    e2e_page = None
    if step_page is not None:
        r = step_page(); r = None
        ms_pg, wall_pg, r = timed(step_page, pe_steps)
        e2e_page = {"value": None, "ms_per_step": ms_pg / pe_steps, "steps": pe_steps}
        r = None
    # ---- the Python surface (GeneFinder): object construction included ----
    e2e_py = None
    try:
        if w["meta"]:
            gf = pyrodigal_b200.GeneFinder(meta=True, device=local_rank)
            run_py = lambda: gf.find_genes_batch(host, offsets)
        elif w.get("train"):
            def run_py():
                gf = pyrodigal_b200.GeneFinder(device=local_rank)
                gf.train(host)
                return gf.find_genes_batch(host, offsets)
        else:
            gf = pyrodigal_b200.GeneFinder(pyrodigal_b200.METAGENOMIC_BINS[w["bin"]].training_info, device=local_rank)
            run_py = lambda: gf.find_genes_batch(host, offsets)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = run_py(); n_py = sum(len(g) for g in out); out = None
            out = run_py(); out = None
            D.barrier()
            t0 = time.perf_counter()
            for _ in range(pe_steps):
                flush_l2()
                out = None
                out = run_py()
            torch.cuda.synchronize()
            wall_py = (time.perf_counter() - t0) * 1e3
            wall_py = D.reduce([wall_py], "max")[0]
        t0 = time.perf_counter()
        n_touch = sum(len(g) for g in out)     # builds every Genes object of the (lazy) batch
        touch_ms = (time.perf_counter() - t0) * 1e3
        e2e_py = {"ms_per_step": wall_py / pe_steps, "steps": pe_steps, "genes": n_py,
                  "api": "GeneFinder.find_genes_batch -> sequence of Genes (views of the result buffers, built on access)",
                  "build_every_genes_object_ms": touch_ms, "genes_touched": n_touch}
        out = None
    except Exception as e:  # the Python surface is a secondary figure: never lose the line over it
        e2e_py = {"error": f"{type(e).__name__}: {e}"}
    batch.free()

    tot_bp, tot_pairs, tot_steps, tot_genes, tot_launch = D.reduce(
        [bp, stats["pairs"], stats["dp_steps"], genes_rank, stats["kernel_launches"]], "sum")

    if rank == 0:
        per_step = ms / args.steps
        peak, peak_src = measured_peak()
        # the DP kernel of this configuration, CUDA events on the launching stream, every timed step
        if w.get("train"):
            dp_kernel = "k_dp_dq<FINAL = 0> (training DP, one warp per chain)"
            dp_ms = float(np.mean([t["train_ms_dp"] for t in step_stats]))
            dp_bytes, dp_note = DP_BYTES_TRAIN, ("kernel_ms = first training phase (GC frame plot + bias + training DP + path + "
                                                 "dicodon counts) between CUDA events; the DP dominates it")
        else:
            dp_kernel = "k_dp_ml (connection-scoring DP, one lane per model)" if w["meta"] else \
                        "k_dp_dq (connection-scoring DP, one warp per chain)"
            dp_ms = float(np.mean([t["ms_dp"] for t in step_stats]))
            dp_bytes, dp_note = DP_BYTES_FINAL, ("kernel_ms = mean over the timed steps of the DP phase (DP kernel + k_chain_best) "
                                                 "between CUDA events on the launching stream")
        achieved = dp_bytes * stats["dp_steps"] / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else 0.0
        traffic = None  # dram__bytes_read+write of the DP kernel per launch, from the committed ncu capture of this build
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_dp_traffic.json")))
            if tj.get("config") == args.config and tj.get("contigs_per_gpu", args.contigs) == args.contigs:
                traffic = tj["traffic_bytes_per_launch"]
        except Exception:
            pass
        mbps = lambda t_ms: tot_bp / (t_ms * 1e-3) / 1e6
        line = base_line(w)
        line.update({
            "value": mbps(per_step), "ms_per_step": per_step,
            "e2e": {"value": mbps(ms_e2e / args.steps), "unit": "Mbp/s",
                    "h2d_bytes_per_step": int(st2["h2d_bytes"]) + (bp if w.get("train") else 0),
                    "d2h_bytes_per_step": int(st2["d2h_bytes"]),
                    "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall_e2e / args.steps,
                    "h2d_ms_per_step": [round(t["ms_h2d"], 2) for t in e2e_steps],   # the input copy alone, per timed step
                    "wall_ms_steps": [round(t["wall_ms"], 1) for t in e2e_steps],
                    "device_ms_steps": [round(t["ms_total_device"], 1) for t in e2e_steps],
                    "host_ms_last_step": [round(x, 1) for x in st2["host_ms"]],
                    "api": "pgpu_train + pgpu_set_models + pgpu_find_genes_batch (C ABI, pinned host input)" if w.get("train")
                           else "pgpu_find_genes_batch (C ABI, pinned host input)"
                           + (" + distributed.gather_result (gene records of all ranks on rank 0, NCCL)" if gathered is not None else "")},
            "gpu_launches": int(tot_launch * args.steps),
            "roofline": {"bound": "hbm", "kernel": dp_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "note": dp_note,
                         "algorithmic_bytes_per_dp_step": dp_bytes, "dp_steps_per_launch": int(stats["dp_steps"]),
                         "kernel_ms": dp_ms, "node_pairs_per_s": stats["pairs"] / (dp_ms * 1e-3) if dp_ms > 0 else None},
            "node_pairs_per_s_job": tot_pairs / (per_step * 1e-3),
            "phases_ms_rank0": dict({k: float(np.mean([t[k] for t in step_stats])) for k in
                                     ("ms_encode", "ms_extract", "ms_score", "ms_overlap", "ms_dp", "ms_trace", "ms_final",
                                      "ms_d2h", "ms_total_device")}, host_issue_ms=stats["host_ms"]),
            "wall_ms_per_step": wall_ms / args.steps,
            "wall_ms_steps": [round(t["wall_ms"], 1) for t in step_stats],
            "device_ms_steps": [round(t["ms_total_device"], 1) for t in step_stats],   # kernels only: wall - device = host between the syncs
            "totals": {"bp": int(tot_bp), "genes": int(tot_genes), "dp_steps": int(tot_steps), "pairs": int(tot_pairs),
                       "chains_rank0": int(stats["n_chains"]), "nodes_rank0": int(stats["total_nodes"])},
            "clocks": clocks,
        })
        if w.get("train"):
            line["phases_ms_rank0"].update({k: float(np.mean([t[k] for t in step_stats])) for k in
                                            ("train_ms_total_device", "train_ms_dp", "train_ms_score")})
        if gathered is not None:
            line["e2e"]["gathered_on_rank0"] = {"contigs": int(gathered.n), "genes": int(gathered.gene_off[-1]),
                                                "ranks": len(gathered.stats)}
        if e2e_page is not None:
            e2e_page["value"] = mbps(e2e_page["ms_per_step"])
            line["e2e"]["pageable_input"] = e2e_page
        if e2e_py is not None:
            if "ms_per_step" in e2e_py:
                e2e_py["value"] = mbps(e2e_py["ms_per_step"])
            line["e2e_python"] = e2e_py
        if not args.no_cpu and world == 1:
            cores = effective_cores()
            if w["meta"]:
                cb = time_cpu_meta(flat, offsets, 1, 1, args.cpu_contigs or 32 * cores)
            else:
                cb = time_cpu_single(w, 1, 0 if w.get("train") else 5_000_000)
            line["cpu_baseline"] = {k: v for k, v in cb.items() if k in ("value", "unit", "cores", "kind", "sample")}
            # what was timed == what the CPU implementation computes: genes (begin, end, strand) of the sample's contigs
            if w["meta"] or cb.get("sample_bp", bp) >= bp:
                chk = step_host()
                n_chk, n_bad = check_parity(chk, cb["genes"])
                line["parity_checked"] = n_chk
                line["parity_mismatches"] = n_bad
                line["parity_genes"] = int(sum(len(g) for g in cb["genes"] if g is not None))
            else:
                # a prefix of one chromosome has different genes at its cut end: run the GPU path on the same prefix
                sub = np.ascontiguousarray(flat[:cb["sample_bp"]])
                chk = ctx.find_genes_batch(sub, np.array([0, len(sub)], np.int64), opts)
                n_chk, n_bad = check_parity(chk, cb["genes"])
                line["parity_checked"] = n_chk
                line["parity_mismatches"] = n_bad
                line["parity_genes"] = int(len(cb["genes"][0]))
                line["parity_note"] = f"checked on the CPU sample ({cb['sample_bp']} bp prefix), run through the same GPU path"
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    D.close()


if __name__ == "__main__":
    main()
