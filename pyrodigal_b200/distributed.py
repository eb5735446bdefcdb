"""Multi-GPU find_genes: contigs shard across the GPUs of one box, one process per GPU (torch.distributed).

The reference spreads `find_genes` over the records of a FASTA file with a thread / process pool
(/root/reference/src/pyrodigal/cli.py:286-300).  Contigs are independent, so the GPU equivalent has no collective on
the data path:

  1. partition: greedy longest-processing-time over an estimated cost per contig (length x node density(GC) x models
     inside the GC window), contigs kept whole -- `lpt_partition`;
  2. input: every rank either already holds the host buffer (`input="replicated"`, e.g. each process mapped the same
     FASTA; only its shard is copied to its GPU over PCIe) or receives its shard from the root rank
     (`input="root"`: one send per rank -- NCCL over NVLink when the group's backend is nccl, the received bytes
     stay in device memory and enter the library through pgpu_batch_wrap_device);
  3. every rank runs the hot path on its shard (`pgpu_find_genes_batch` / `pgpu_batch_run`);
  4. output: all_gather of the per-rank gene counts, then the fixed-size records (per-contig summaries, `pgpu_gene`,
     the (start, stop) `pgpu_node` pair of every gene) are gathered on the root rank and put back into input order.

torch is plumbing here (process group, device / pinned buffers); nothing in this module computes on genes.
`runner` lets the tests drive the same code with another engine (the CUDA-on-CPU emulation of tests/emu) over gloo."""
import heapq
import os
import time

import numpy as np

from . import _capi

__all__ = ["contig_cost", "lpt_partition", "estimate_gc", "ShardedResult", "find_genes_sharded", "gather_result"]


def estimate_gc(flat, offsets, sample=2048):
    """GC fraction per contig from a strided sample of at most `sample` bases (planning only, never a result)"""
    n = len(offsets) - 1
    gc = np.full(n, 0.5)
    lens = np.diff(offsets)
    isgc = np.zeros(256, dtype=bool)
    isgc[[ord(c) for c in "GCgcSs"]] = True
    for k in range(n):
        L = int(lens[k])
        if L <= 0:
            continue
        step = max(1, L // sample)
        s = flat[int(offsets[k]):int(offsets[k + 1]):step]
        gc[k] = float(isgc[s].mean())
    return gc


def contig_cost(lengths, gc, model_gc=None):
    """estimated work per contig ~ chain-nodes: length x nodes per kbp (grows with GC: fewer stop codons, longer ORFs,
    more starts; SURVEY.md 8a: 14-62 nodes / kbp) x number of models inside the GC window (lib.pyx:5335-5342)"""
    lengths = np.asarray(lengths, dtype=np.float64)
    gc = np.clip(np.asarray(gc, dtype=np.float64), 0.2, 0.8)
    density = 14.0 + (gc - 0.3) / 0.4 * 48.0
    if model_gc is None:
        models = np.full(len(lengths), 12.0)
    else:
        mg = np.sort(np.asarray(model_gc, dtype=np.float64))
        low = np.minimum(0.65, 0.88495 * gc - 0.0102337)
        high = np.maximum(0.35, 0.86596 * gc + 0.1131991)
        models = np.maximum(1, np.searchsorted(mg, high, "right") - np.searchsorted(mg, low, "left"))
    return lengths * np.maximum(density, 5.0) * models + 1.0


def lpt_partition(cost, world):
    """greedy longest-processing-time: contigs in decreasing cost order, each to the least loaded rank.
    -> owner[n] (int32).  Deterministic (ties: lower contig index first, lower rank first)."""
    cost = np.asarray(cost, dtype=np.float64)
    owner = np.zeros(len(cost), dtype=np.int32)
    if world <= 1 or len(cost) == 0:
        return owner
    order = np.lexsort((np.arange(len(cost)), -cost))
    heap = [(0.0, r) for r in range(world)]
    for k in order:
        load, r = heapq.heappop(heap)
        owner[k] = r
        heapq.heappush(heap, (load + float(cost[k]), r))
    return owner


class ShardedResult:
    """What the root rank gets back, addressed in the ORIGINAL contig order (the layout of `_capi.Result`).
    `summary` / `gene_off` are built at once (32 B per contig); the gene records stay in the per-rank buffers they
    arrived in -- `contig_genes(k)` / `contig_gene_nodes(k)` are zero-copy views -- and `genes` / `gene_nodes`
    (everything, contig-major) are put together on first use only."""

    def __init__(self, parts, n_total, stats):
        self.n, self.stats, self._parts = n_total, stats, parts
        self.summary = np.zeros(n_total, dtype=_capi.SUMMARY_DTYPE)
        self._rank = np.zeros(n_total, dtype=np.int32)
        self._first = np.zeros(n_total, dtype=np.int64)   # first gene of the contig inside its rank's gene array
        for r, (ids, rsum, rgen, rnod) in enumerate(parts):
            if len(ids) == 0:
                continue
            self.summary[ids] = rsum
            first = np.zeros(len(ids), dtype=np.int64)
            np.cumsum(rsum["n_genes"][:-1], out=first[1:])
            self._rank[ids], self._first[ids] = r, first
        self.gene_off = np.zeros(n_total + 1, dtype=np.int64)
        np.cumsum(self.summary["n_genes"], out=self.gene_off[1:])
        self._genes = self._gene_nodes = None

    def contig_genes(self, k):
        a = int(self._first[k])
        return self._parts[self._rank[k]][2][a:a + int(self.summary["n_genes"][k])]

    def contig_gene_nodes(self, k):
        a = int(self._first[k])
        return self._parts[self._rank[k]][3][a:a + int(self.summary["n_genes"][k])]

    def _materialise(self):
        ng = int(self.gene_off[-1])
        if len(self._parts) == 1 and np.array_equal(self._parts[0][0], np.arange(self.n)):
            self._genes, self._gene_nodes = self._parts[0][2], np.asarray(self._parts[0][3]).reshape(ng, 2)
            return
        genes = np.zeros(ng, dtype=_capi.GENE_DTYPE)
        gnodes = np.zeros((ng, 2), dtype=_capi.NODE_DTYPE)
        cnt = self.summary["n_genes"].astype(np.int64)
        contig_of = np.repeat(np.arange(self.n), cnt)
        src = self._first[contig_of] + (np.arange(ng) - self.gene_off[:-1][contig_of])
        rk = self._rank[contig_of]
        for r, (ids, rsum, rgen, rnod) in enumerate(self._parts):
            sel = rk == r
            if sel.any():
                genes[sel] = rgen[src[sel]]
                gnodes[sel] = np.asarray(rnod).reshape(-1, 2)[src[sel]]
        self._genes, self._gene_nodes = genes, gnodes

    @property
    def genes(self):
        if self._genes is None:
            self._materialise()
        return self._genes

    @property
    def gene_nodes(self):
        if self._gene_nodes is None:
            self._materialise()
        return self._gene_nodes


class _Dist:
    def __init__(self, group):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        self.backend = dist.get_backend(group) if self.on else None
        self.cuda = self.backend == "nccl"

    def dev(self, device):
        return self.torch.device(f"cuda:{device}") if self.cuda else self.torch.device("cpu")

    def global_rank(self, r):
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r


_pinned_cache = {}
_gather_seq = [0]   # the receive buffers alternate between two sets: a result stays valid during the NEXT gather


def _pinned(key, size, torch):
    """a page-locked uint8 host tensor of at least `size` bytes, recycled between calls (allocation costs milliseconds).
    Over nccl the arrays of a ShardedResult alias these buffers: they stay valid until the second-next gather of the
    process (copy what must live longer)."""
    t = _pinned_cache.get(key)
    if t is None or t.numel() < size:
        t = torch.empty(max(size + size // 4, 1 << 20), dtype=torch.uint8).pin_memory()
        _pinned_cache[key] = t
    return t


_buffer_cache = {}


def _buffer(key, size, dev, torch):
    """a uint8 tensor of at least `size` bytes on `dev`, recycled between calls"""
    t = _buffer_cache.get((key, str(dev)))
    if t is None or t.numel() < size:
        t = torch.empty(max(size + size // 4, 1 << 16), dtype=torch.uint8, device=dev)
        _buffer_cache[(key, str(dev))] = t
    return t


def _shard_view(flat, offsets, ids):
    """contigs `ids` back to back: (flat_shard, offsets_shard); a zero-copy slice when the ids are one contiguous run"""
    lens = (offsets[1:] - offsets[:-1])[ids]
    off = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    if len(ids) and np.array_equal(ids, np.arange(ids[0], ids[0] + len(ids))):
        return flat[int(offsets[ids[0]]):int(offsets[ids[-1] + 1])], off
    out = np.empty(int(off[-1]), dtype=np.uint8)
    for j, k in enumerate(ids):
        out[off[j]:off[j + 1]] = flat[int(offsets[k]):int(offsets[k + 1])]
    return out, off


def gather_result(local, ids, n_total, group=None, root=0, device=0):
    """Gather step (4): `local` is this rank's result (`.summary`, `.genes`, `.gene_nodes` for its contigs `ids`, in
    that order).  Returns a ShardedResult on the root rank, None elsewhere.  Over nccl the records travel GPU to GPU
    (staged through device tensors); over gloo they are sent from host memory."""
    D = _Dist(group)
    stats = dict(getattr(local, "stats", {}) or {})
    if D.world == 1:
        return ShardedResult([(np.asarray(ids, dtype=np.int64), local.summary, local.genes, local.gene_nodes)], n_total, [stats])
    torch, dist = D.torch, D.dist
    dev = D.dev(device)
    ng = int(local.summary["n_genes"].sum())
    # fixed-size header per rank: (#contigs, #genes)
    head = torch.tensor([len(ids), ng], dtype=torch.int64, device=dev)
    heads = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(D.world)]
    dist.all_gather(heads, head, group=group)
    heads = [tuple(int(x) for x in h.tolist()) for h in heads]

    def as_bytes(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))

    S, G, N = _capi.SUMMARY_DTYPE.itemsize, _capi.GENE_DTYPE.itemsize, _capi.NODE_DTYPE.itemsize
    trace = os.environ.get("PGPU_GATHER_TRACE") == "1"
    t0 = time.perf_counter()
    # one message per rank: [contig ids | summaries | genes | gene nodes], padded to the largest one, so that the whole
    # exchange is ONE gather collective (over nccl: device buffers, NVLink) instead of three point-to-point messages per
    # rank that the root would have to take one after the other
    size_of = [nc * (8 + S) + g * (G + 2 * N) for nc, g in heads]
    maxb = max(max(size_of), 16)
    pieces = [as_bytes(np.asarray(ids, dtype=np.int64)), as_bytes(local.summary)]
    if ng:
        pieces += [as_bytes(local.genes), as_bytes(local.gene_nodes)]
    send = _buffer(("send", _gather_seq[0] & 1), maxb, dev, torch)
    o = 0
    for t in pieces:   # from the result's page-locked buffers: asynchronous DMA when the destination is a device tensor
        send[o:o + t.numel()].copy_(t, non_blocking=True)
        o += t.numel()
    if D.rank != root:
        dist.gather(send[:maxb], None, dst=D.global_rank(root), group=group)
        if D.cuda:
            torch.cuda.current_stream(dev).synchronize()   # the result buffers may be recycled once this returns
        if trace:
            print(f"[gather rank {D.rank}] sent {o / 1e6:.1f} MB in {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
        _gather_seq[0] += 1
        return None
    recv = [_buffer(("recv", r, _gather_seq[0] & 1), maxb, dev, torch)[:maxb] for r in range(D.world)]
    dist.gather(send[:maxb], recv, dst=D.global_rank(root), group=group)
    parts, all_stats = [], [stats]
    hosts = {}
    for r in range(D.world):
        if r == root or size_of[r] == 0:
            continue
        if D.cuda:   # device -> host copies of what arrived over NVLink, into recycled page-locked memory, all in flight
            hosts[r] = _pinned(("gather", r, _gather_seq[0] & 1), size_of[r], torch)
            hosts[r][:size_of[r]].copy_(recv[r][:size_of[r]], non_blocking=True)
        else:
            hosts[r] = recv[r].clone()   # gloo: the receive buffers are recycled
    if D.cuda:
        torch.cuda.current_stream(dev).synchronize()
    _gather_seq[0] += 1
    for r in range(D.world):
        nc, g = heads[r]
        if r == root:
            parts.append((np.asarray(ids, dtype=np.int64), local.summary, np.asarray(local.genes), np.asarray(local.gene_nodes)))
            continue
        if size_of[r] == 0:
            parts.append((np.zeros(0, np.int64), np.zeros(0, _capi.SUMMARY_DTYPE), np.zeros(0, _capi.GENE_DTYPE),
                          np.zeros((0, 2), _capi.NODE_DTYPE)))
            continue
        bb = hosts[r][:size_of[r]].numpy()
        o = 0
        rid = bb[o:o + 8 * nc].view(np.int64); o += 8 * nc
        rsum = bb[o:o + S * nc].view(_capi.SUMMARY_DTYPE); o += S * nc
        rgen = bb[o:o + G * g].view(_capi.GENE_DTYPE); o += G * g
        rnod = bb[o:o + 2 * N * g].view(_capi.NODE_DTYPE).reshape(g, 2)
        parts.append((rid, rsum, rgen, rnod))
    if trace:
        print(f"[gather rank {D.rank}] gathered {sum(size_of) / 1e6:.1f} MB from {D.world} ranks in "
              f"{(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    return ShardedResult(parts, n_total, all_stats)


def find_genes_sharded(flat, offsets, *, group=None, root=0, input="replicated", device=None, finder=None,
                       runner=None, cost=None):
    """find_genes over contigs flat[offsets[k]:offsets[k+1]] on all ranks of `group`; see the module docstring.

    flat / offsets   uint8 ASCII buffer and int64[n+1] offsets (the C ABI's batch layout).  `input="replicated"`: every
                     rank passes the same buffers; `input="root"`: only the root rank's buffers are read, the other
                     ranks may pass None.
    finder           a `pyrodigal_b200.GeneFinder` that supplies the options and the model set (default: meta mode)
    runner           `(flat_shard_or_None, offsets_shard, device_tensor_or_None) -> result`, replaces the CUDA library
                     (tests); default: the finder's context on `device` (default: the rank's index in the group)
    cost             optional per-contig cost estimate (default: `contig_cost` on a sampled GC content)
    Returns a ShardedResult on the root rank, None on the others."""
    D = _Dist(group)
    torch, dist = D.torch, D.dist
    if device is None:
        device = D.rank if D.cuda else 0
    dev = D.dev(device)
    have_input = input == "replicated" or D.rank == root
    # ---- 1. plan ----
    if have_input:
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        if cost is None:
            model_gc = None
            if finder is not None and getattr(finder, "meta", False):
                model_gc = [b.training_info.gc for b in finder.metagenomic_bins]
            cost = contig_cost(np.diff(offsets), estimate_gc(flat, offsets), model_gc)
        owner = lpt_partition(cost, D.world)
    if input == "root" and D.world > 1:
        meta = torch.tensor([n if have_input else 0], dtype=torch.int64, device=dev)
        dist.broadcast(meta, src=D.global_rank(root), group=group)
        n = int(meta.item())
        plan = torch.empty(2 * n + 1, dtype=torch.int64, device=dev)
        if have_input:
            plan.copy_(torch.from_numpy(np.concatenate([offsets, owner.astype(np.int64)])))
        dist.broadcast(plan, src=D.global_rank(root), group=group)
        plan = plan.cpu().numpy()
        offsets, owner = plan[:n + 1], plan[n + 1:].astype(np.int32)
    ids = np.flatnonzero(owner == D.rank)
    # ---- 2. input ----
    shard_dev = None
    if input == "root" and D.world > 1:
        lens = (offsets[1:] - offsets[:-1])
        if D.rank == root:
            reqs, keep = [], []
            for r in range(D.world):
                if r == root:
                    continue
                rid = np.flatnonzero(owner == r)
                sh, _ = _shard_view(flat, offsets, rid)
                t = torch.from_numpy(np.ascontiguousarray(sh))
                if D.cuda:
                    t = t.pin_memory().to(dev, non_blocking=True)
                keep.append(t)
                if t.numel():
                    reqs.append(dist.isend(t, dst=D.global_rank(r), group=group))
            shard, shard_off = _shard_view(flat, offsets, ids)
            for q in reqs:
                q.wait()
        else:
            shard_off = np.zeros(len(ids) + 1, dtype=np.int64)
            np.cumsum(lens[ids], out=shard_off[1:])
            t = torch.empty(int(shard_off[-1]), dtype=torch.uint8, device=dev)
            if t.numel():
                dist.recv(t, src=D.global_rank(root), group=group)
            if D.cuda:
                shard, shard_dev = None, t      # stays in device memory
            else:
                shard = t.numpy()
    else:
        shard, shard_off = _shard_view(flat, offsets, ids)
    # ---- 3. run ----
    if runner is None:
        runner = _default_runner(finder, device)
    local = runner(shard, shard_off, shard_dev)
    # ---- 4. gather ----
    return gather_result(local, ids, n, group=group, root=root, device=device)


def _default_runner(finder, device):
    from . import lib as L
    if finder is None:
        finder = L.GeneFinder(meta=True, device=device)
    finder.device = device

    def run(shard, shard_off, shard_dev):
        ctx = finder._context()
        with ctx.lock:
            if shard_dev is not None:
                b = ctx.wrap_device(shard_dev.data_ptr(), shard_off, keepalive=shard_dev)
                try:
                    return b.run(finder._opts(False))
                finally:
                    b.free()
            return ctx.find_genes_batch(np.ascontiguousarray(shard), shard_off, finder._opts(False))
    return run
