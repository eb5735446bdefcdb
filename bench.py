#!/usr/bin/env python
"""bench.py -- Mbp/s of find_genes (meta mode) on synthetic contig batches, 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

Workload (BASELINE.json configs[3]): the 100 000-contig synthetic metagenome
(contig lengths uniform 1-100 kbp, GC uniform 0.30-0.70, iid bases, seeds 1 000 000 + k), sharded
over the GPUs of one box: every rank owns 12 500 contigs (~630 Mbp), so per-GPU work is fixed as N
grows ("weak" scaling) and N = 8 is exactly the 100k-contig / ~5 Gbp batch.  A "step" is one pass of
the whole hot path (encode -> add_nodes -> score_nodes -> overlapping starts -> connection DP for every
(contig, model) chain -> winner / traceback / genes -> final re-score) over the rank's shard.

Printed JSON line (rank 0): see the task contract; additionally `roofline` (dominant kernel = the
connection-scoring DP), `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`, `phases`.
"""
import argparse
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
DP_BYTES_PER_STEP = 72  # algorithmic HBM bytes of one DP step (final=1)


def make_contigs(first, count, seed=4):
    """cfg4 contigs [first, first+count): (flat uint8 ASCII array, int64 offsets)"""
    rng = np.random.default_rng(seed)
    # lengths / gc of the whole 100k-contig config are drawn once so that shards are disjoint slices of it
    lengths = rng.integers(1_000, 100_001, size=100_000)
    gcs = rng.uniform(0.30, 0.70, size=100_000)
    idx = np.arange(first, first + count) % 100_000
    lens = lengths[idx]
    offsets = np.zeros(count + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    flat = np.empty(int(offsets[-1]), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for k, i in enumerate(idx):
        g = np.random.default_rng(1_000_000 + int(i))
        gc = gcs[i]
        # P(A)=P(T)=(1-gc)/2, P(C)=P(G)=gc/2 through a single uniform draw per base
        u = g.random(int(lens[k]), dtype=np.float32)
        a = (1 - gc) / 2
        code = (u >= a).astype(np.uint8) + (u >= a + gc / 2) + (u >= a + gc)
        flat[offsets[k]:offsets[k + 1]] = lut[code]
    return flat, offsets


def shard_range(rank, contigs_per_gpu):
    """contigs owned by `rank`: a disjoint slice of the 100k-contig configuration (weak scaling)"""
    return rank * contigs_per_gpu, contigs_per_gpu


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


_REF_GF = None


def _ref_init(ref_dir):
    """process-pool worker initialiser: import the unmodified reference and build one GeneFinder"""
    global _REF_GF
    sys.path.insert(0, ref_dir)
    import pyrodigal
    _REF_GF = pyrodigal.GeneFinder(meta=True)


def _ref_work(seq):
    return len(_REF_GF.find_genes(seq))


def cpu_reference_runners():
    """[(label, kind, fn(list_of_bytes) -> total genes, close)] -- the reference's own CPU implementation with every
    host core: its documented thread-pool recipe (docs/guide/parallel.rst:24-41, cli.py:286-300) and the process
    pool its CLI also offers (cli.py:292-293).  Falls back to the C oracle port when oracle/_ref is absent."""
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
                    lambda seqs: sum(len(g) for g in tp.map(gf.find_genes, seqs)), tp.close))
        try:
            pp = mp.get_context("spawn").Pool(cores, initializer=_ref_init, initargs=(ref_dir,))
            out.append((f"pyrodigal {pyrodigal.__version__} GeneFinder(meta=True) multiprocessing.Pool({cores})", "reference",
                        lambda seqs: sum(pp.map(_ref_work, seqs, chunksize=1)), pp.terminate))
        except Exception:
            pass
    except Exception as e:
        from oracle import oracle as orc
        import refutil as R
        blob = R.bins_blob()
        tp = ThreadPool(cores)

        def one(s):
            d, gc, unk = orc.encode(s)
            return len(orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, blob)[0])
        out.append((f"C oracle port ThreadPool({cores}) [{type(e).__name__}: {e}]", "port",
                    lambda seqs: sum(tp.map(one, seqs)), tp.close))
    return out, cores


def time_cpu(flat, offsets, steps, warmup, max_contigs):
    """times every available CPU configuration on the same bounded sample and reports the fastest"""
    runners, cores = cpu_reference_runners()
    n = min(len(offsets) - 1, max_contigs)
    # longest contigs first: the pools then finish without a long tail
    seqs = sorted((flat[offsets[k]:offsets[k + 1]].tobytes() for k in range(n)), key=len, reverse=True)
    bp = int(offsets[n] - offsets[0])
    best = None
    tried = []
    for label, kind, run, close in runners:
        try:
            for _ in range(max(1, warmup)):
                run(seqs[: max(cores, n // 8)])
            t0 = time.perf_counter()
            genes = 0
            for _ in range(steps):
                genes = run(seqs)
            dt = (time.perf_counter() - t0) / steps
            tried.append(f"{label}: {bp / dt / 1e6:.1f} Mbp/s")
            if best is None or dt < best["s_per_step"]:
                best = {"value": bp / dt / 1e6, "unit": "Mbp/s", "cores": cores, "kind": kind, "s_per_step": dt,
                        "genes": genes, "label": label}
        finally:
            try:
                close()
            except Exception:
                pass
    best["sample"] = (f"first {n} contigs of the rank-0 shard ({bp / 1e6:.1f} Mbp); fastest of: " + "; ".join(tried))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--contigs", type=int, default=CONTIGS_PER_GPU, help="contigs per GPU (default = cfg4 shard)")
    ap.add_argument("--cpu-contigs", type=int, default=0, help="contigs in the CPU sample (0 = 16 per core)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)
    config = {"workload": "cfg4: 100k-contig synthetic metagenome (1-100 kbp, GC 0.30-0.70), meta mode, "
                          f"{args.contigs} contigs per GPU, contig-sharded", "contigs_per_gpu": args.contigs,
              "parallelism": f"contig-shard x{n_gpus}", "l2_policy": "inputs larger than L2 (shard >> 126 MB)"}

    # ------------------------------------------------------------------ reference arm (CPU) --------
    if args.impl == "reference":
        if rank != 0:
            return
        flat, offsets = make_contigs(0, args.contigs if args.contigs < CONTIGS_PER_GPU else 2048)
        cores = effective_cores()
        cb = time_cpu(flat, offsets, args.steps, min(args.warmup, 1), args.cpu_contigs or 32 * cores)
        line = {"impl": "reference", "metric": "Mbp/s find_genes (meta mode)", "value": cb["value"], "unit": "Mbp/s",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["s_per_step"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm --------------------
    import torch
    from pyrodigal_b200 import _capi
    import refutil as R

    torch.cuda.set_device(local_rank)
    D = Dist("nccl", f"cuda:{local_rank}")
    first, count = shard_range(rank, args.contigs)
    flat, offsets = make_contigs(first, count)
    bp = int(offsets[-1])
    # pinned host staging of the step's input (the e2e leg copies from here every step)
    pinned = torch.empty(bp, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = flat
    host = pinned.numpy()

    ctx = _capi.Context(local_rank)
    ctx.set_models(R.bins_blob(), 50)
    opts = _capi.make_opts(meta=True)

    def timed(fn, steps, per_step=None):
        """K steps bracketed by barrier + synchronize; CUDA events on the library's stream; max over ranks"""
        D.barrier()
        ctx.timer_start()
        t0 = time.perf_counter()
        last = None
        for _ in range(steps):
            last = None  # release the previous result first: its pinned buffer is reused by the next step
            t1 = time.perf_counter()
            last = fn()
            if per_step is not None:
                per_step.append(dict(last.stats, wall_ms=(time.perf_counter() - t1) * 1e3))
        ms = ctx.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        ms, wall = D.reduce([ms, wall], "max")
        D.barrier()
        return ms, wall, last

    # ---- device-resident leg ("value") ----
    batch = ctx.upload(host, offsets)
    r = None
    for _ in range(args.warmup):
        r = None  # at most two results alive at a time: their pinned buffers are recycled by the library
        r = batch.run(opts)
    r = None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if rank == 0:
        sampler.start()
    step_stats = []
    ms, wall_ms, res = timed(lambda: batch.run(opts), args.steps, step_stats)
    clocks = sampler.stop() if rank == 0 else None
    stats = res.stats
    genes_rank = int(res.summary["n_genes"].sum())
    n_contigs_rank = res.n

    # ---- end-to-end leg: C ABI call with host buffers, H2D + D2H inside ----
    res = None
    for _ in range(max(2, args.warmup)):
        r = None
        r = ctx.find_genes_batch(host, offsets, opts)
    r = None
    e2e_steps = []
    ms_e2e, wall_e2e, res2 = timed(lambda: ctx.find_genes_batch(host, offsets, opts), args.steps, e2e_steps)
    st2 = res2.stats
    batch.free()

    tot_bp, tot_pairs, tot_steps, tot_genes, tot_launch = D.reduce(
        [bp, stats["pairs"], stats["dp_steps"], genes_rank, stats["kernel_launches"]], "sum")

    if rank == 0:
        per_step = ms / args.steps
        peak, peak_src = measured_peak()
        dp_ms = float(np.mean([t["ms_dp"] for t in step_stats]))  # CUDA events on the launching stream, every timed step
        achieved = DP_BYTES_PER_STEP * stats["dp_steps"] / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else 0.0
        traffic = None  # dram__bytes_read+write of the DP kernel per launch, from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_dp_traffic.json")))
            if tj.get("contigs_per_gpu") == args.contigs:
                traffic = tj["traffic_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "Mbp/s find_genes (meta mode)", "value": tot_bp / (per_step * 1e-3) / 1e6, "unit": "Mbp/s",
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "e2e": {"value": tot_bp / (ms_e2e / args.steps * 1e-3) / 1e6, "unit": "Mbp/s",
                    "h2d_bytes_per_step": int(st2["h2d_bytes"]), "d2h_bytes_per_step": int(st2["d2h_bytes"]),
                    "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall_e2e / args.steps,
                    "h2d_ms_per_step": [round(t["ms_h2d"], 2) for t in e2e_steps],   # the input copy alone, per timed step
                    "wall_ms_steps": [round(t["wall_ms"], 1) for t in e2e_steps],
                    "device_ms_steps": [round(t["ms_total_device"], 1) for t in e2e_steps],
                    "api": "pgpu_find_genes_batch (C ABI, pinned host input)"},
            "gpu_launches": int(tot_launch * args.steps),
            "roofline": {"bound": "hbm", "kernel": "k_dp_ml (connection-scoring DP, one lane per model)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "note": "kernel_ms = mean over the timed steps of the DP phase (k_dp_ml + k_chain_best) between CUDA events "
                                 "on the launching stream; the DP is latency bound (ncu in profiles/: issue-active 39 %, "
                                 "~370 warp-instructions per warp step covering ~11 chains), not HBM bound; DRAM traffic "
                                 "1.2x the algorithmic bytes (suffix-maximum arrays of the window maximum)",
                         "algorithmic_bytes_per_dp_step": DP_BYTES_PER_STEP, "dp_steps_per_launch": int(stats["dp_steps"]),
                         "kernel_ms": dp_ms, "node_pairs_per_s": stats["pairs"] / (dp_ms * 1e-3) if dp_ms > 0 else None},
            "node_pairs_per_s_job": tot_pairs / (per_step * 1e-3),
            "phases_ms_rank0": dict({k: stats[k] for k in ("ms_encode", "ms_extract", "ms_score", "ms_overlap", "ms_dp",
                                                           "ms_trace", "ms_final", "ms_d2h", "ms_total_device")},
                                    host_issue_ms=stats["host_ms"]),
            "wall_ms_per_step": wall_ms / args.steps,
            "wall_ms_steps": [round(t["wall_ms"], 1) for t in step_stats],
            "totals": {"bp": int(tot_bp), "genes": int(tot_genes), "dp_steps": int(tot_steps), "pairs": int(tot_pairs),
                       "chains_rank0": int(stats["n_chains"]), "nodes_rank0": int(stats["total_nodes"])},
            "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            cores = effective_cores()
            line["cpu_baseline"] = {k: v for k, v in time_cpu(flat, offsets, 1, 1, args.cpu_contigs or 32 * cores).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    D.close()


if __name__ == "__main__":
    main()
