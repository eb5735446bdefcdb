"""ctypes/numpy binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (pyrodigal_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

NODE_DTYPE = np.dtype(
    [
        ("ndx", "i4"), ("stop_val", "i4"), ("strand", "i4"), ("type", "i4"),
        ("edge", "i4"), ("elim", "i4"), ("gc_bias", "i4"),
        ("star_ptr", "i4", (3,)), ("traceb", "i4"), ("tracef", "i4"), ("ov_mark", "i4"),
        ("rbs", "i4", (2,)),
        ("mot_ndx", "i4"), ("mot_len", "i4"), ("mot_spacer", "i4"), ("mot_spacendx", "i4"),
        ("mot_score", "f8"), ("gc_score", "f8", (3,)),
        ("cscore", "f8"), ("uscore", "f8"), ("tscore", "f8"), ("rscore", "f8"), ("sscore", "f8"), ("score", "f8"),
        ("gc_cont", "f4"), ("_pad", "i4"),
    ],
    align=True,
)
GENE_DTYPE = np.dtype([("begin", "i4"), ("end", "i4"), ("start_ndx", "i4"), ("stop_ndx", "i4")])
TRAINING_SIZE = 558392


class Opts(C.Structure):
    _fields_ = [
        ("closed", C.c_int32), ("min_gene", C.c_int32), ("min_edge_gene", C.c_int32), ("max_overlap", C.c_int32),
        ("n_masks", C.c_int32), ("masks", C.POINTER(C.c_int32)),
    ]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "pyrodigal_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        assert L.orc_sizeof_node() == NODE_DTYPE.itemsize, (L.orc_sizeof_node(), NODE_DTYPE.itemsize)
        assert L.orc_sizeof_training() == TRAINING_SIZE
        _lib = L
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def make_opts(closed=False, min_gene=90, min_edge_gene=60, max_overlap=60, masks=None):
    o = Opts(int(closed), min_gene, min_edge_gene, max_overlap, 0, None)
    if masks is not None and len(masks):
        m = np.ascontiguousarray(np.asarray(masks, dtype=np.int32).reshape(-1))
        o._keep = m
        o.n_masks = len(m) // 2
        o.masks = m.ctypes.data_as(C.POINTER(C.c_int32))
    return o


def encode(seq):
    """ASCII bytes/str -> (digits u8[n], gc_count, unknown)"""
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    a = np.frombuffer(bytes(seq), dtype=np.uint8)
    d = np.empty(len(a), dtype=np.uint8)
    gc = C.c_int(0)
    unk = lib().orc_encode(_p(a), len(a), _p(d), C.byref(gc))
    return d, gc.value, unk


def find_masks(digits, mask_size=50):
    cap = max(16, len(digits) // max(1, mask_size) + 2)
    out = np.empty((cap, 2), dtype=np.int32)
    n = lib().orc_find_masks(_p(digits), len(digits), mask_size, _p(out), cap)
    return out[:n].copy()


def node_capacity(slen):
    return max(1024, slen // 4 + 64)


def extract(digits, tt, opts=None, sort=True):
    opts = opts or make_opts()
    cap = node_capacity(len(digits))
    nodes = np.zeros(cap, dtype=NODE_DTYPE)
    n = lib().orc_extract(_p(digits), len(digits), tt, C.byref(opts), _p(nodes), cap)
    assert n >= 0
    nodes = nodes[:n].copy()
    if sort:
        lib().orc_sort(_p(nodes), n)
    return nodes


def tinf_ptr(blob):
    assert len(blob) == TRAINING_SIZE
    buf = np.frombuffer(blob, dtype=np.uint8)
    return buf, _p(buf)


def reset_scores(nodes):
    lib().orc_reset_scores(_p(nodes), len(nodes))


def score(digits, nodes, tinf_blob, closed=False, is_meta=False):
    keep, tp = tinf_ptr(tinf_blob)
    lib().orc_score(_p(digits), len(digits), _p(nodes), len(nodes), tp, int(closed), int(is_meta))


def record_overlapping_starts(nodes, tinf_blob, flag=1, max_overlap=60):
    keep, tp = tinf_ptr(tinf_blob)
    lib().orc_record_overlapping_starts(_p(nodes), len(nodes), tp, flag, max_overlap)


def score_connections(nodes, tinf_blob, final=True):
    keep, tp = tinf_ptr(tinf_blob)
    pairs = C.c_int64(0)
    lib().orc_score_connections(_p(nodes), len(nodes), tp, int(final), C.byref(pairs))
    return pairs.value


def dynamic_programming(nodes, tinf_blob, final=True):
    keep, tp = tinf_ptr(tinf_blob)
    return lib().orc_dynamic_programming(_p(nodes), len(nodes), tp, int(final))


def skippable(nodes, j, i):
    return lib().orc_skippable(_p(nodes), j, i)


def shine_dalgarno(digits, pos, start, rbs_wt, strand=1, exact=True):
    w = np.ascontiguousarray(rbs_wt, dtype=np.float64)
    f = lib().orc_shine_dalgarno_exact if exact else lib().orc_shine_dalgarno_mm
    return f(_p(digits), len(digits), pos, start, _p(w), strand)


def find_genes_single(digits, tinf_blob, opts=None):
    opts = opts or make_opts()
    keep, tp = tinf_ptr(tinf_blob)
    cap = node_capacity(len(digits))
    nodes = np.zeros(cap, dtype=NODE_DTYPE)
    gcap = max(64, len(digits) // 60 + 16)
    genes = np.zeros(gcap, dtype=GENE_DTYPE)
    nn = C.c_int(0)
    ipath = C.c_int(-1)
    ng = lib().orc_find_genes_single(_p(digits), len(digits), tp, C.byref(opts), _p(nodes), cap, C.byref(nn),
                                     _p(genes), gcap, C.byref(ipath))
    assert ng >= 0
    return genes[:ng].copy(), nodes[: nn.value].copy(), ipath.value


def find_genes_meta(digits, gc, bins_blob, opts=None, n_bins=None):
    """bins_blob: concatenated raw training structs. returns (genes, nodes, winner, pairs)"""
    opts = opts or make_opts()
    buf = np.frombuffer(bins_blob, dtype=np.uint8)
    n_bins = n_bins or len(buf) // TRAINING_SIZE
    cap = node_capacity(len(digits))
    nodes = np.zeros(cap, dtype=NODE_DTYPE)
    gcap = max(64, len(digits) // 60 + 16)
    genes = np.zeros(gcap, dtype=GENE_DTYPE)
    nn = C.c_int(0)
    winner = C.c_int(-1)
    pairs = C.c_int64(0)
    f = lib().orc_find_genes_meta
    f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    ng = f(_p(digits), len(digits), float(gc), _p(buf), n_bins, C.byref(opts), _p(nodes), cap, C.byref(nn),
           _p(genes), gcap, C.byref(winner), C.byref(pairs))
    assert ng >= 0
    return genes[:ng].copy(), nodes[: nn.value].copy(), winner.value, pairs.value


# ---- training (GeneFinder.train) ------------------------------------------------------------------
TRAINING_DTYPE = np.dtype(
    [
        ("gc", "f8"), ("trans_table", "i4"), ("_pad0", "i4"), ("st_wt", "f8"), ("bias", "f8", (3,)),
        ("type_wt", "f8", (3,)), ("uses_sd", "i4"), ("_pad1", "i4"), ("rbs_wt", "f8", (28,)),
        ("ups_comp", "f8", (32, 4)), ("mot_wt", "f8", (4, 4, 4096)), ("no_mot", "f8"), ("gene_dc", "f8", (4096,)),
    ]
)
assert TRAINING_DTYPE.itemsize == TRAINING_SIZE


def gc_frame_plot(digits):
    gp = np.empty(len(digits), dtype=np.int8)
    lib().orc_gc_frame_plot(_p(digits), len(digits), _p(gp))
    return gp


def train(digits, gc, translation_table=11, start_weight=4.35, force_nonsd=False, opts=None, return_nodes=False):
    """GeneFinder.train on an encoded sequence -> raw training struct (bytes)"""
    opts = opts or make_opts()
    cap = node_capacity(len(digits))
    nodes = np.zeros(cap, dtype=NODE_DTYPE)
    out = np.zeros(1, dtype=TRAINING_DTYPE)
    f = lib().orc_train
    f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p]
    nn = f(_p(digits), len(digits), float(gc), translation_table, float(start_weight), int(force_nonsd),
           C.byref(opts), _p(nodes), cap, _p(out))
    assert nn >= 0
    blob = out.tobytes()
    return (blob, nodes[:nn].copy()) if return_nodes else blob
