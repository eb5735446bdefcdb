"""ctypes binding of libpyrodigal_b200.so (the C ABI declared in include/pyrodigal_b200.h).

There is no CPU fallback: importing this module fails loudly when the CUDA library has not been
built, and creating a context fails when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpyrodigal_b200.so")
TRAINING_SIZE = 558392

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C pyrodigal_b200/csrc` (nvcc, sm_100a).  pyrodigal_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)


class Opts(C.Structure):
    _fields_ = [
        ("meta", C.c_int32), ("single_model", C.c_int32), ("closed", C.c_int32), ("mask", C.c_int32),
        ("min_mask", C.c_int32), ("min_gene", C.c_int32), ("min_edge_gene", C.c_int32), ("max_overlap", C.c_int32),
        ("want_nodes", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_contigs", C.c_int64), ("total_bp", C.c_int64), ("total_nodes", C.c_int64),
        ("total_chain_nodes", C.c_int64), ("n_chains", C.c_int64), ("total_genes", C.c_int64),
        ("pairs", C.c_int64), ("dp_steps", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("ms_total_device", C.c_double), ("ms_encode", C.c_double), ("ms_extract", C.c_double),
        ("ms_score", C.c_double), ("ms_overlap", C.c_double), ("ms_dp", C.c_double), ("ms_trace", C.c_double),
        ("ms_final", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("reserved", C.c_double * 4),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["host_ms"] = [self.reserved[i] for i in range(4)]  # host time issuing work before each of the 4 syncs
        return d


class TrainOpts(C.Structure):
    _fields_ = [("translation_table", C.c_int32), ("force_nonsd", C.c_int32), ("start_weight", C.c_double),
                ("reserved", C.c_int32 * 4)]


GENE_DTYPE = np.dtype([("begin", "<i4"), ("end", "<i4"), ("start_ndx", "<i4"), ("stop_ndx", "<i4")])

NODE_DTYPE = np.dtype(
    {
        "names": ["ndx", "stop_val", "traceb", "tracef", "star_ptr", "strand", "type", "edge", "elim", "ov_mark",
                  "rbs", "mot_len", "mot_ndx", "mot_spacer", "mot_spacendx", "gc_cont", "mot_score", "cscore",
                  "uscore", "tscore", "rscore", "sscore", "score"],
        "formats": ["<i4", "<i4", "<i4", "<i4", ("<i4", (3,)), "i1", "u1", "u1", "u1", "i1", ("u1", (2,)), "u1",
                    "<u2", "u1", "u1", "<f4", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "<f8"],
        "offsets": [0, 4, 8, 12, 16, 28, 29, 30, 31, 32, 33, 35, 36, 38, 39, 40, 48, 56, 64, 72, 80, 88, 96],
        "itemsize": 104,
    }
)

# `struct _node` of the reference (src/Prodigal/node.h:41-76, 128 bytes as packed by Pyrodigal): what
# pgpu_result_nodes_struct writes.  `mot_bits` = the motif bit fields: ndx (12 bits) | spacer (4) << 12 | len (3) << 16 |
# spacendx (2) << 19.
NODE_STRUCT_DTYPE = np.dtype(
    {
        "names": ["mot_score", "mot_bits", "gc_score", "cscore", "uscore", "tscore", "rscore", "sscore", "score", "gc_cont",
                  "star_ptr", "traceb", "tracef", "ndx", "stop_val", "ov_mark", "strand", "rbs", "edge", "elim", "gc_bias",
                  "type"],
        "formats": ["<f8", "<u4", ("<f8", (3,)), "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "<f4", ("<i4", (3,)), "<i4", "<i4",
                    "<i4", "<i4", "i1", "i1", ("u1", (2,)), "u1", "u1", "u1", "u1"],
        "offsets": [0, 8, 16, 40, 48, 56, 64, 72, 80, 88, 92, 104, 108, 112, 116, 120, 121, 122, 124, 125, 126, 127],
        "itemsize": 128,
    }
)

SUMMARY_DTYPE = np.dtype(
    [("n_genes", "<i4"), ("n_nodes", "<i4"), ("winner", "<i4"), ("ipath", "<i4"), ("unknown", "<i4"),
     ("gc_count", "<i4"), ("score", "<f8")]
)
assert SUMMARY_DTYPE.itemsize == 32

_vp = C.c_void_p
lib.pgpu_create.argtypes = [C.c_int, C.POINTER(_vp)]
lib.pgpu_destroy.argtypes = [_vp]
lib.pgpu_destroy.restype = None
lib.pgpu_last_error.argtypes = [_vp]
lib.pgpu_last_error.restype = C.c_char_p
lib.pgpu_set_models.argtypes = [_vp, _vp, C.c_int, C.c_size_t]
lib.pgpu_num_models.argtypes = [_vp]
lib.pgpu_set_workspace_limit.argtypes = [_vp, C.c_size_t]
lib.pgpu_host_alloc.argtypes = [C.c_size_t]
lib.pgpu_host_alloc.restype = _vp
lib.pgpu_host_free.argtypes = [_vp]
lib.pgpu_host_free.restype = None
lib.pgpu_timer_start.argtypes = [_vp]
lib.pgpu_timer_stop.argtypes = [_vp, C.POINTER(C.c_double)]
lib.pgpu_find_genes_batch.argtypes = [_vp, _vp, _vp, C.c_int, C.POINTER(Opts), C.POINTER(_vp)]
lib.pgpu_batch_upload.argtypes = [_vp, _vp, _vp, C.c_int, C.POINTER(_vp)]
lib.pgpu_batch_wrap_device.argtypes = [_vp, _vp, _vp, C.c_int, C.POINTER(_vp)]
lib.pgpu_batch_run.argtypes = [_vp, _vp, C.POINTER(Opts), C.POINTER(_vp)]
lib.pgpu_batch_free.argtypes = [_vp]
lib.pgpu_batch_free.restype = None
lib.pgpu_result_num_contigs.argtypes = [_vp]
lib.pgpu_result_summaries.argtypes = [_vp, _vp]
lib.pgpu_result_genes.argtypes = [_vp, C.c_int, _vp]
lib.pgpu_result_all_genes.argtypes = [_vp, _vp]
lib.pgpu_result_gene_nodes.argtypes = [_vp, _vp]
lib.pgpu_result_num_segments.argtypes = [_vp]
lib.pgpu_result_segment.argtypes = [_vp, C.c_int, C.POINTER(C.c_longlong), C.POINTER(_vp), C.POINTER(_vp)]
lib.pgpu_result_segment.restype = C.c_longlong
lib.pgpu_result_nodes.argtypes = [_vp, C.c_int, _vp]
lib.pgpu_result_nodes_struct.argtypes = [_vp, C.c_int, _vp]
lib.pgpu_skippable.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, _vp]
lib.pgpu_skippable.restype = None
lib.pgpu_result_stats.argtypes = [_vp, C.POINTER(Stats)]
lib.pgpu_result_free.argtypes = [_vp]
lib.pgpu_result_free.restype = None
lib.pgpu_train.argtypes = [_vp, _vp, C.c_int64, C.POINTER(Opts), C.POINTER(TrainOpts), _vp, C.POINTER(Stats)]
lib.pgpu_extract_nodes.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(Opts), C.c_int, _vp, _vp, _vp, _vp, _vp]
lib.pgpu_score_nodes.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(Opts), C.c_int, C.c_int, C.c_int, _vp]
lib.pgpu_max_gc_frame_plot.argtypes = [_vp, _vp, C.c_int, _vp]
lib.pgpu_shine_dalgarno.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
lib.pgpu_score_connections.argtypes = [_vp, C.c_int] + [_vp] * 10 + [C.c_int, C.c_int] + [_vp] * 5
lib.pgpu_compute_skippable.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp]

PGPU_ENODEV, PGPU_ENOMEM, PGPU_EINVAL, PGPU_ESTATE, PGPU_ECUDA = -1, -2, -3, -4, -5
# error code -> the exception type the reference raises in the same situation
_EXC = {PGPU_ENODEV: RuntimeError, PGPU_ENOMEM: MemoryError, PGPU_EINVAL: ValueError, PGPU_ESTATE: RuntimeError,
        PGPU_ECUDA: RuntimeError}


def check(rc, ctx=None):
    if rc >= 0:
        return rc
    msg = lib.pgpu_last_error(ctx)
    raise _EXC.get(rc, RuntimeError)((msg or b"").decode() or f"pyrodigal_b200 error {rc}")


def ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


class _PinnedBlock:
    def __init__(self, p):
        self.p = p

    def __del__(self):
        try:
            if self.p:
                lib.pgpu_host_free(self.p)
                self.p = None
        except Exception:
            pass


def pinned_empty(n):
    """uint8 array of n bytes in page-locked host memory (pgpu_host_alloc), or an ordinary array when that is not
    available (no CUDA device / allocation failed); freed with the last view of it"""
    p = lib.pgpu_host_alloc(max(int(n), 1)) if n > 0 else None
    if not p:
        return np.empty(n, dtype=np.uint8)
    buf = (C.c_ubyte * int(n)).from_address(p)
    buf._owner = _PinnedBlock(p)
    return np.frombuffer(buf, dtype=np.uint8)


def make_opts(meta=False, single_model=0, closed=False, mask=False, min_mask=50, min_gene=90, min_edge_gene=60,
              max_overlap=60, want_nodes=False):
    o = Opts()
    o.meta, o.single_model, o.closed, o.mask = int(meta), int(single_model), int(closed), int(mask)
    o.min_mask, o.min_gene, o.min_edge_gene, o.max_overlap = min_mask, min_gene, min_edge_gene, max_overlap
    o.want_nodes = int(want_nodes)
    return o


class Context:
    """One pgpu_ctx: a device, a stream, a model set."""

    def __init__(self, device=0):
        h = _vp()
        rc = lib.pgpu_create(int(device), C.byref(h))
        if rc < 0:
            raise _EXC.get(rc, RuntimeError)(
                "pyrodigal_b200: " + (lib.pgpu_last_error(None) or b"").decode() + " (no CPU fallback exists)")
        self.handle = h
        self.device = device
        self.model_key = None

    def close(self):
        if self.handle:
            lib.pgpu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_models(self, blob, n, key=None):
        buf = np.frombuffer(blob, dtype=np.uint8)
        assert len(buf) == n * TRAINING_SIZE
        check(lib.pgpu_set_models(self.handle, ptr(buf), n, TRAINING_SIZE), self.handle)
        self.model_key = key

    def timer_start(self):
        check(lib.pgpu_timer_start(self.handle), self.handle)

    def timer_stop(self):
        ms = C.c_double(0)
        check(lib.pgpu_timer_stop(self.handle, C.byref(ms)), self.handle)
        return ms.value

    # ---- hot path -------------------------------------------------------------------------------
    def find_genes_batch(self, seq, offsets, opts):
        """seq: uint8 array (ASCII), offsets: int64[n+1] -> Result"""
        res = _vp()
        check(lib.pgpu_find_genes_batch(self.handle, ptr(seq), ptr(offsets), len(offsets) - 1, C.byref(opts),
                                        C.byref(res)), self.handle)
        return Result(res, self)

    def upload(self, seq, offsets):
        b = _vp()
        check(lib.pgpu_batch_upload(self.handle, ptr(seq), ptr(offsets), len(offsets) - 1, C.byref(b)), self.handle)
        return Batch(self, b)

    def wrap_device(self, device_ptr, offsets, keepalive=None):
        """a Batch over ASCII bytes that already sit in this device's memory (device_ptr: int); `keepalive` (e.g. the
        torch tensor that owns the memory) is held until the Batch is freed"""
        b = _vp()
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        check(lib.pgpu_batch_wrap_device(self.handle, _vp(int(device_ptr)), ptr(offsets), len(offsets) - 1, C.byref(b)),
              self.handle)
        bt = Batch(self, b)
        bt.keepalive = keepalive
        return bt

    def train(self, seq, opts, translation_table=11, start_weight=4.35, force_nonsd=False):
        """GeneFinder.train on one (already joined) ASCII sequence -> (raw training struct bytes, stats dict)"""
        t = TrainOpts()
        t.translation_table, t.force_nonsd, t.start_weight = int(translation_table), int(bool(force_nonsd)), float(start_weight)
        out = np.zeros(TRAINING_SIZE, dtype=np.uint8)
        st = Stats()
        check(lib.pgpu_train(self.handle, ptr(seq), len(seq), C.byref(opts), C.byref(t), ptr(out), C.byref(st)),
              self.handle)
        return out.tobytes(), st.as_dict()

    # ---- operators ------------------------------------------------------------------------------
    def extract_nodes(self, seq, translation_table, opts):
        n = check(lib.pgpu_extract_nodes(self.handle, ptr(seq), len(seq), translation_table, C.byref(opts), 0, None,
                                         None, None, None, None), self.handle)
        out = {"ndx": np.empty(n, np.int32), "stop_val": np.empty(n, np.int32), "strand": np.empty(n, np.int8),
               "type": np.empty(n, np.uint8), "edge": np.empty(n, np.uint8)}
        check(lib.pgpu_extract_nodes(self.handle, ptr(seq), len(seq), translation_table, C.byref(opts), n,
                                     ptr(out["ndx"]), ptr(out["stop_val"]), ptr(out["strand"]), ptr(out["type"]),
                                     ptr(out["edge"])), self.handle)
        return out

    def score_nodes(self, seq, model, opts, is_meta=False, first_pass=True):
        # every position can carry at most one node per strand (plus the closing STOP nodes of the six frames)
        cap = 2 * len(seq) + 16
        if cap > (1 << 16):   # long sequences: ask for the count first (dst = NULL) instead of reserving 208 B per base
            cap = check(lib.pgpu_score_nodes(self.handle, ptr(seq), len(seq), model, C.byref(opts), int(is_meta),
                                             int(first_pass), 0, None), self.handle)
        out = np.zeros(max(cap, 1), dtype=NODE_DTYPE)
        n = check(lib.pgpu_score_nodes(self.handle, ptr(seq), len(seq), model, C.byref(opts), int(is_meta),
                                       int(first_pass), cap, ptr(out)), self.handle)
        return out[:n].copy()

    def max_gc_frame_plot(self, seq):
        out = np.empty(len(seq), dtype=np.int8)
        check(lib.pgpu_max_gc_frame_plot(self.handle, ptr(seq), len(seq), ptr(out)), self.handle)
        return out

    def shine_dalgarno(self, seq, pos, start, model=0, strand=1, exact=True):
        out = C.c_int32(0)
        check(lib.pgpu_shine_dalgarno(self.handle, ptr(seq), len(seq), pos, start, model, strand, int(exact),
                                      C.byref(out)), self.handle)
        return out.value

    def score_connections(self, ndx, stop_val, strand, type_, cscore, sscore, rscore, uscore, gc_score, star_ptr,
                          model, final):
        n = len(ndx)
        a = lambda x, t: np.ascontiguousarray(x, dtype=t)
        ndx, stop_val, strand, type_ = a(ndx, np.int32), a(stop_val, np.int32), a(strand, np.int8), a(type_, np.uint8)
        cscore, sscore, rscore, uscore = (a(x, np.float64) for x in (cscore, sscore, rscore, uscore))
        gc_score = a(gc_score, np.float64) if gc_score is not None else None
        star_ptr = a(star_ptr, np.int32)
        score, traceb, ov = np.zeros(n, np.float64), np.zeros(n, np.int32), np.zeros(n, np.int8)
        pairs, ms = C.c_int64(0), C.c_double(0)
        check(lib.pgpu_score_connections(self.handle, n, ptr(ndx), ptr(stop_val), ptr(strand), ptr(type_), ptr(cscore),
                                         ptr(sscore), ptr(rscore), ptr(uscore), ptr(gc_score), ptr(star_ptr), model,
                                         int(final), ptr(score), ptr(traceb), ptr(ov),
                                         C.cast(C.byref(pairs), _vp), C.cast(C.byref(ms), _vp)), self.handle)
        return score, traceb, ov, pairs.value, ms.value

    def compute_skippable(self, strand, type_, ndx, mn, i):
        n = len(ndx)
        skip = np.zeros(n, np.uint8)
        check(lib.pgpu_compute_skippable(self.handle, n, ptr(np.ascontiguousarray(strand, np.int8)),
                                         ptr(np.ascontiguousarray(type_, np.uint8)),
                                         ptr(np.ascontiguousarray(ndx, np.int32)), mn, i, ptr(skip)), self.handle)
        return skip


def skippable_plugin(strands, types, frames, mn, i):
    """pgpu_skippable: the skip filter through the reference's plug-in signature (skippable_t), on the process-wide context"""
    skip = np.full(len(types), 0xAA, np.uint8)   # only [mn, i) is written
    lib.pgpu_skippable(ptr(np.ascontiguousarray(strands, np.uint8)), ptr(np.ascontiguousarray(types, np.uint8)),
                       ptr(np.ascontiguousarray(frames, np.uint8)), mn, i, ptr(skip))
    return skip


class Batch:
    """Device-resident input (pgpu_batch): upload once, run many times."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    def run(self, opts):
        res = _vp()
        check(lib.pgpu_batch_run(self.ctx.handle, self.handle, C.byref(opts), C.byref(res)), self.ctx.handle)
        return Result(res, self.ctx)

    def free(self):
        if self.handle:
            lib.pgpu_batch_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _ResultHandle:
    """Owns the pgpu_result: freed when the last reference (the Result, or a numpy view of its buffers) goes away."""

    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            if self.h:
                lib.pgpu_result_free(self.h)
                self.h = None
        except Exception:
            pass


class Result:
    """A pgpu_result.  `summary`, `gene_off` and `stats` are read eagerly (small); `genes` and `gene_nodes` are
    materialised on first use: zero-copy numpy views of the result's page-locked buffers when the batch ran as one
    sub-batch, a concatenated copy otherwise.  The views keep the underlying pgpu_result alive through a handle object
    -- not through this Result, which would form a reference cycle and leave the (large, page-locked) buffers to the
    cyclic garbage collector."""

    def __init__(self, handle, ctx=None):
        self._h = _ResultHandle(handle)
        self.ctx = ctx  # the result returns its pinned buffers to the context when freed: keep it alive
        self.n = lib.pgpu_result_num_contigs(handle)
        self.summary = np.zeros(self.n, dtype=SUMMARY_DTYPE)
        if self.n:
            lib.pgpu_result_summaries(handle, ptr(self.summary))
        self.gene_off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(self.summary["n_genes"], out=self.gene_off[1:])
        self._genes = self._gene_nodes = None
        st = Stats()
        lib.pgpu_result_stats(handle, C.byref(st))
        self.stats = st.as_dict()

    @property
    def handle(self):
        return self._h.h if self._h is not None else None

    def _materialise(self):
        if self.handle is None:
            raise RuntimeError("result already freed")
        ng = int(self.gene_off[-1])
        nseg = lib.pgpu_result_num_segments(self.handle)
        if ng and nseg == 1:
            g0, pg, pn = C.c_longlong(0), _vp(), _vp()
            cnt = lib.pgpu_result_segment(self.handle, 0, C.byref(g0), C.byref(pg), C.byref(pn))
            assert cnt == ng and g0.value == 0
            gbuf = (C.c_char * (ng * GENE_DTYPE.itemsize)).from_address(pg.value)
            nbuf = (C.c_char * (2 * ng * NODE_DTYPE.itemsize)).from_address(pn.value)
            gbuf._owner = nbuf._owner = self._h  # the buffers live as long as any view of them
            self._genes = np.frombuffer(gbuf, dtype=GENE_DTYPE)
            self._gene_nodes = np.frombuffer(nbuf, dtype=NODE_DTYPE).reshape(ng, 2)
            self._genes.flags.writeable = self._gene_nodes.flags.writeable = False
        else:
            self._genes = np.zeros(ng, dtype=GENE_DTYPE)
            self._gene_nodes = np.zeros((ng, 2), dtype=NODE_DTYPE)
            if ng:
                lib.pgpu_result_all_genes(self.handle, ptr(self._genes))
                lib.pgpu_result_gene_nodes(self.handle, ptr(self._gene_nodes))

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

    def nodes_struct(self, contig):
        """the final node array of one contig in the reference's own `struct _node` layout (NODE_STRUCT_DTYPE)"""
        n = int(self.summary["n_nodes"][contig])
        out = np.zeros(n, dtype=NODE_STRUCT_DTYPE)
        rc = lib.pgpu_result_nodes_struct(self.handle, contig, ptr(out))
        if rc < 0:
            raise RuntimeError("node arrays were not requested (want_nodes=False)")
        return out

    def nodes(self, contig):
        n = int(self.summary["n_nodes"][contig])
        out = np.zeros(n, dtype=NODE_DTYPE)
        rc = lib.pgpu_result_nodes(self.handle, contig, ptr(out))
        if rc < 0:
            raise RuntimeError("node arrays were not requested (want_nodes=False)")
        return out

    def free(self):
        """drop this object's reference; the buffers are released once no view of them is left"""
        self._genes = self._gene_nodes = None
        self._h = None
