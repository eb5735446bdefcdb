"""Host-side mirror of the reference's public surface for the find_genes path.

Same class names, constructor keywords, attributes and error behaviour as `pyrodigal.lib`
(src/pyrodigal/lib.pyx; stubs src/pyrodigal/lib.pyi) for: GeneFinder.find_genes (+ the added batched
find_genes_many), Genes, Gene, Nodes, Node, Sequence, Mask(s), TrainingInfo, MetagenomicBin(s) and
lib.ConnectionScorer.  All computation happens in libpyrodigal_b200.so on the GPU through the C ABI
(include/pyrodigal_b200.h); this module only marshals buffers and formats results.
"""
import datetime
import collections.abc
import itertools
import json
import lzma
import math
import os
import textwrap
import threading
import typing
import warnings

import numpy as np

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
TRAINING_SIZE = _capi.TRAINING_SIZE
MIN_SINGLE_GENOME = 20000
IDEAL_SINGLE_GENOME = 100000
TRANSLATION_TABLES = frozenset((1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 21, 22, 23, 24, 25, 26, 29, 30, 32, 33))

# lib.pyx:209-228
_RBS_MOTIF = [None, "GGA/GAG/AGG", "3Base/5BMM", "4Base/6BMM", "AGxAG", "AGxAG", "GGA/GAG/AGG", "GGxGG", "GGxGG",
              "AGxAG", "AGGAG(G)/GGAGG", "AGGA/GGAG/GAGG", "AGGA/GGAG/GAGG", "GGA/GAG/AGG", "GGxGG", "AGGA",
              "GGAG/GAGG", "AGxAGG/AGGxGG", "AGxAGG/AGGxGG", "AGxAGG/AGGxGG", "AGGAG/GGAGG", "AGGAG", "AGGAG",
              "GGAGG", "GGAGG", "AGGAGG", "AGGAGG", "AGGAGG"]
_RBS_SPACER = [None, "3-4bp", "13-15bp", "13-15bp", "11-12bp", "3-4bp", "11-12bp", "11-12bp", "3-4bp", "5-10bp",
               "13-15bp", "3-4bp", "11-12bp", "5-10bp", "5-10bp", "5-10bp", "5-10bp", "11-12bp", "3-4bp", "5-10bp",
               "11-12bp", "3-4bp", "5-10bp", "3-4bp", "5-10bp", "11-12bp", "3-4bp", "5-10bp"]
_NODE_TYPE = ["ATG", "GTG", "TTG", "Edge"]
_LETTERS = "AGCTNNN"
# the writers emit the reference's version tag so that their output is byte-identical to pyrodigal's
# (lib.pyx:3485, 3595, 3605, 3826); `pyrodigal_b200.__version__` is this package's own version
PYRODIGAL_COMPAT_VERSION = "3.7.1"

# Genetic codes in NCBI order (first base T, C, A, G; then second, then third): the published NCBI tables
# (https://www.ncbi.nlm.nih.gov/Taxonomy/Utils/wprintgc.cgi), written as differences from the standard code.
_NCBI_STANDARD = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
_NCBI_CHANGES = {
    1: {}, 11: {}, 2: {"AGA": "*", "AGG": "*", "ATA": "M", "TGA": "W"},
    3: {"ATA": "M", "CTT": "T", "CTC": "T", "CTA": "T", "CTG": "T", "TGA": "W"},
    4: {"TGA": "W"}, 5: {"AGA": "S", "AGG": "S", "ATA": "M", "TGA": "W"}, 6: {"TAA": "Q", "TAG": "Q"},
    9: {"AAA": "N", "AGA": "S", "AGG": "S", "TGA": "W"}, 10: {"TGA": "C"}, 12: {"CTG": "S"},
    13: {"AGA": "G", "AGG": "G", "ATA": "M", "TGA": "W"}, 14: {"AAA": "N", "AGA": "S", "AGG": "S", "TAA": "Y", "TGA": "W"},
    15: {"TAG": "Q"}, 16: {"TAG": "L"}, 21: {"TGA": "W", "ATA": "M", "AGA": "S", "AGG": "S", "AAA": "N"},
    22: {"TCA": "*", "TAG": "L"}, 23: {"TTA": "*"}, 24: {"AGA": "S", "AGG": "K", "TGA": "W"}, 25: {"TGA": "G"},
    26: {"CTG": "A"}, 27: {"TAG": "Q", "TAA": "Q", "TGA": "W"}, 28: {"TAA": "Q", "TAG": "Q", "TGA": "W"},
    29: {"TAA": "Y", "TAG": "Y"}, 30: {"TAA": "E", "TAG": "E"}, 31: {"TGA": "W", "TAG": "E", "TAA": "E"},
    32: {"TAG": "W"}, 33: {"TAA": "Y", "TGA": "W", "AGA": "S", "AGG": "K"},
}
_DIGIT_OF = {"A": 0, "G": 1, "C": 2, "T": 3}


def _translation_table(tt):
    """64-entry amino acid table indexed by (x0 << 4) + (x1 << 2) + x2 over the digit alphabet A0 G1 C2 T3
    (the indexing of _translation.h:37-40)"""
    tab = _TT_CACHE.get(tt)
    if tab is None:
        code = {}
        for k, aa in enumerate(_NCBI_STANDARD):
            code["TCAG"[k >> 4] + "TCAG"[(k >> 2) & 3] + "TCAG"[k & 3]] = aa
        code.update(_NCBI_CHANGES[tt])
        tab = np.zeros(64, dtype=np.uint8)
        for codon, aa in code.items():
            tab[(_DIGIT_OF[codon[0]] << 4) + (_DIGIT_OF[codon[1]] << 2) + _DIGIT_OF[codon[2]]] = ord(aa)
        _TT_CACHE[tt] = tab
    return tab


_TT_CACHE = {}
# stop codons per table exactly as the reference lists them (lib.pyx:174-201; the listing is only used to decide
# whether Gene.translate warns about a table with different stops, so its quirks -- e.g. table 6 -- are kept)
_STOP_LISTING = {
    1: ("TAA", "TAG", "TGA"), 2: ("TAA", "TAG", "AGA", "AGG"), 3: ("TAA", "TAG"), 4: ("TAA", "TAG"), 5: ("TAA", "TAG"),
    6: ("TAA", "TAG", "TGA"), 9: ("TAA", "TAG"), 10: ("TAA", "TAG"), 11: ("TAA", "TAG", "TGA"), 12: ("TAA", "TAG", "TGA"),
    13: ("TAA", "TAG"), 14: "TAG", 15: ("TAA", "TGA"), 16: ("TAA", "TGA"), 21: ("TAA", "TAG"), 22: ("TCA", "TAA", "TGA"),
    23: ("TTA", "TAA", "TGA"), 24: ("TAA", "TAG"), 25: ("TAA", "TAG"), 26: ("TAA", "TAG", "TGA"), 27: (), 28: (),
    29: "TGA", 30: "TGA", 31: (), 32: ("TAA", "TGA"), 33: "TAG",
}


def _stop_codons(tt):
    return _STOP_LISTING[tt]

# field offsets inside `struct _training` (vendor/Prodigal/training.h:29-51)
_T_DTYPE = np.dtype(
    [("gc", "<f8"), ("trans_table", "<i4"), ("_p0", "<i4"), ("st_wt", "<f8"), ("bias", "<f8", (3,)),
     ("type_wt", "<f8", (3,)), ("uses_sd", "<i4"), ("_p1", "<i4"), ("rbs_wt", "<f8", (28,)),
     ("ups_comp", "<f8", (32, 4)), ("mot_wt", "<f8", (4, 4, 4096)), ("no_mot", "<f8"), ("gene_dc", "<f8", (4096,))]
)
assert _T_DTYPE.itemsize == TRAINING_SIZE


# --- Training info ----------------------------------------------------------------------------------
class TrainingInfo:
    """Parameters of one model, stored in the reference's raw struct layout (lib.pyx:3898-4283)."""

    def __init__(self, gc, *, translation_table=11, start_weight=4.35, bias=None, type_weights=None, uses_sd=True,
                 rbs_weights=None, upstream_compositions=None, motif_weights=None, missing_motif_weight=0.0,
                 coding_statistics=None):
        if translation_table not in TRANSLATION_TABLES:
            raise ValueError(f"{translation_table} is not a valid translation table index")
        self._raw = np.zeros(1, dtype=_T_DTYPE)
        r = self._raw[0]
        r["gc"], r["trans_table"], r["st_wt"], r["uses_sd"], r["no_mot"] = gc, translation_table, start_weight, int(uses_sd), missing_motif_weight
        for name, val in (("bias", bias), ("type_wt", type_weights), ("rbs_wt", rbs_weights),
                          ("ups_comp", upstream_compositions), ("mot_wt", motif_weights), ("gene_dc", coding_statistics)):
            if val is not None:
                r[name] = np.asarray(val, dtype=np.float64).reshape(r[name].shape)

    @classmethod
    def _from_bytes(cls, raw):
        if len(raw) != TRAINING_SIZE:
            raise EOFError(f"Expected {TRAINING_SIZE} bytes, only read {len(raw)}")
        self = cls.__new__(cls)
        self._raw = np.frombuffer(bytes(raw), dtype=_T_DTYPE).copy()
        return self

    @classmethod
    def load(cls, fp):
        """Load a raw `struct _training` file as written by Prodigal / TrainingInfo.dump (lib.pyx:3911-3953)."""
        return cls._from_bytes(fp.read(TRAINING_SIZE))

    def dump(self, fp):
        fp.write(self._raw.tobytes())

    def __bytes__(self):
        return self._raw.tobytes()

    def __repr__(self):
        return f"<pyrodigal_b200.TrainingInfo gc={self.gc!r} translation_table={self.translation_table!r}>"

    def to_dict(self):
        """The constructor's keyword arguments as plain Python objects (lib.pyx:4829-4862), e.g. for JSON."""
        return {
            "gc": self.gc,
            "translation_table": self.translation_table,
            "start_weight": self.start_weight,
            "bias": self.bias.tolist(),
            "type_weights": self.type_weights.tolist(),
            "uses_sd": self.uses_sd,
            "rbs_weights": self.rbs_weights.tolist(),
            "upstream_compositions": self.upstream_compositions.tolist(),
            "motif_weights": self.motif_weights.tolist(),
            "missing_motif_weight": self.missing_motif_weight,
            "coding_statistics": self.coding_statistics.tolist(),
        }

    def __sizeof__(self):
        return TRAINING_SIZE + object.__sizeof__(self)

    def __getstate__(self):
        return {"raw": self._raw.tobytes()}

    def __setstate__(self, state):
        if "raw" in state:
            self._raw = np.frombuffer(state["raw"], dtype=_T_DTYPE).copy()
        else:  # the reference pickles the `to_dict` form (lib.pyx:4024-4045)
            state = dict(state)
            self.__init__(state.pop("gc"), **state)

    def __eq__(self, other):
        return isinstance(other, TrainingInfo) and self._raw.tobytes() == other._raw.tobytes()

    def __hash__(self):
        return hash(self._raw.tobytes()[:2048])

    def _scalar(field, cast):
        def fget(self):
            return cast(self._raw[0][field])

        def fset(self, value):
            if field == "trans_table" and value not in TRANSLATION_TABLES:
                raise ValueError(f"{value} is not a valid translation table index")
            self._raw[0][field] = value
        return property(fget, fset)

    def _array(field):
        def fget(self):
            return self._raw[0][field]

        def fset(self, value):
            dst = self._raw[0][field]
            src = np.asarray(value, dtype=np.float64)
            if src.size != dst.size:
                raise ValueError(f"expected {dst.size} values, got {src.size}")
            dst[...] = src.reshape(dst.shape)
        return property(fget, fset)

    translation_table = _scalar("trans_table", int)
    gc = _scalar("gc", float)
    start_weight = _scalar("st_wt", float)
    uses_sd = _scalar("uses_sd", bool)
    missing_motif_weight = _scalar("no_mot", float)
    bias = _array("bias")
    type_weights = _array("type_wt")
    rbs_weights = _array("rbs_wt")
    upstream_compositions = _array("ups_comp")
    motif_weights = _array("mot_wt")
    coding_statistics = _array("gene_dc")
    del _scalar, _array


class MetagenomicBin:
    def __init__(self, training_info, description):
        self.training_info = training_info
        self.description = description

    def __repr__(self):
        return f"<pyrodigal_b200.MetagenomicBin description={self.description!r}>"

    def __reduce__(self):
        return type(self), (self.training_info, self.description)


class MetagenomicBins(typing.Sequence):
    def __init__(self, iterable):
        self._bins = list(iterable)
        self._blob = None

    def __len__(self):
        return len(self._bins)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return MetagenomicBins(self._bins[index])
        return self._bins[index]

    def __reduce__(self):
        return type(self), (self._bins,)

    def _blob_bytes(self):
        if self._blob is None:
            self._blob = b"".join(bytes(b.training_info) for b in self._bins)
        return self._blob


def _load_builtin_bins():
    """The 50 pre-trained Prodigal models (data dumped by tools/dump_metagenomic_bins.py)."""
    with open(os.path.join(_HERE, "data", "metagenomic_bins.json")) as f:
        meta = json.load(f)
    with lzma.open(os.path.join(_HERE, "data", "metagenomic_bins.bin.xz")) as f:
        blob = f.read()
    n, stride = meta["count"], meta["stride"]
    assert len(blob) == n * stride and stride == TRAINING_SIZE
    bins = MetagenomicBins(
        MetagenomicBin(TrainingInfo._from_bytes(blob[i * stride:(i + 1) * stride]), meta["descriptions"][i])
        for i in range(n))
    bins._blob = blob
    return bins


class _LazyBins:
    _bins = None

    @classmethod
    def get(cls):
        if cls._bins is None:
            cls._bins = _load_builtin_bins()
        return cls._bins


def __getattr__(name):
    if name == "METAGENOMIC_BINS":
        return _LazyBins.get()
    raise AttributeError(name)


# --- contexts (one per device and model set) -----------------------------------------------------------
_ctx_lock = threading.Lock()
_ctx_cache = {}


class _CtxLock:
    """The per-context lock callers hold around a library call.  Leaving it also drops the use count that
    `_context_for` / `_training_context` took for the caller, and a context is only evicted from the cache (and
    destroyed) while its use count is zero -- so no thread can be inside, or on its way into, a call on a context
    that is being destroyed."""

    def __init__(self, ctx):
        self._lock, self._ctx = threading.Lock(), ctx

    def __enter__(self):
        self._lock.acquire()
        return self

    def __exit__(self, *exc):
        self._lock.release()
        with _ctx_lock:
            self._ctx.users = max(0, self._ctx.users - 1)
        return False


def _context_for(model_blob, n_models, device=0):
    """the cached context of (device, model set), with its use count taken: the caller enters `ctx.lock` exactly once"""
    key = (device, n_models, hash(model_blob))
    with _ctx_lock:
        ctx = _ctx_cache.get(key)
        if ctx is None:
            # keep the device footprint bounded: drop an IDLE model context (never the training context, never one
            # that a thread holds or is about to lock); if every context is busy the cache grows instead
            idle = [k for k, c in _ctx_cache.items() if k[1] != "train" and getattr(c, "users", 0) == 0]
            if len(_ctx_cache) >= 4 and idle:
                _ctx_cache.pop(idle[0]).close()
            ctx = _capi.Context(device)
            ctx.set_models(model_blob, n_models, key)
            ctx.users = 0
            ctx.lock = _CtxLock(ctx)
            _ctx_cache[key] = ctx
        ctx.users += 1
        return ctx


def _training_context(device=0):
    """training needs a device and a stream but no model set: one model-less context per device (never evicted)"""
    key = (device, "train")
    with _ctx_lock:
        ctx = _ctx_cache.get(key)
        if ctx is None:
            ctx = _capi.Context(device)
            ctx.users = 0
            ctx.lock = _CtxLock(ctx)
            _ctx_cache[key] = ctx
        ctx.users += 1
        return ctx


def _as_ascii(sequence):
    if isinstance(sequence, Sequence):
        return sequence._ascii
    if isinstance(sequence, str):
        return np.frombuffer(sequence.encode("ascii", "replace"), dtype=np.uint8)
    return np.frombuffer(memoryview(sequence).cast("B"), dtype=np.uint8)


# --- Sequence / masks -------------------------------------------------------------------------------
class Mask:
    def __init__(self, begin, end):
        self.begin, self.end = begin, end

    def __repr__(self):
        return f"<pyrodigal_b200.Mask begin={self.begin!r} end={self.end!r}>"

    def __eq__(self, other):
        return isinstance(other, Mask) and (self.begin, self.end) == (other.begin, other.end)

    def intersects(self, begin, end):
        return self.begin < end and begin < self.end


class Masks(list):
    """List of `Mask` regions (lib.pyx:345-470); pickles as the reference does, as a list of (begin, end) pairs."""

    def copy(self):
        return Masks(Mask(m.begin, m.end) for m in self)

    __copy__ = copy

    def __getstate__(self):
        return [(m.begin, m.end) for m in self]

    def __setstate__(self, state):
        self[:] = [Mask(b, e) for b, e in state]

    def __reduce__(self):
        return (Masks, (), self.__getstate__())

    def __sizeof__(self):
        return 8 * len(self) + object.__sizeof__(self)


_ENC = np.full(256, 6, dtype=np.uint8)
for _c, _v in ((b"Aa", 0), (b"Gg", 1), (b"Cc", 2), (b"Tt", 3)):
    for _b in _c:
        _ENC[_b] = _v


class Sequence(typing.Sized):
    """Input sequence.  The digitised form (`bytes(memoryview(seq))` in the reference, lib.pyx:642-658) is
    produced lazily on the host for accessors only; the compute path encodes on the GPU."""

    def __init__(self, sequence, mask=False, mask_size=50):
        if isinstance(sequence, Sequence):
            self._ascii = sequence._ascii
        else:
            self._ascii = _as_ascii(sequence)
        self._digits = None
        self._mask, self._mask_size = mask, mask_size
        self._masks = None
        self._gc_count = self._unknown = None

    def __len__(self):
        return len(self._ascii)

    @property
    def digits(self):
        if self._digits is None:
            self._digits = _ENC[self._ascii]
        return self._digits

    def __bytes__(self):
        return self.digits.tobytes()

    def __str__(self):
        return "".join(_LETTERS[d] for d in self.digits)

    def _counts(self):
        if self._gc_count is None:
            d = self.digits
            self._gc_count = int(np.count_nonzero((d == 1) | (d == 2)))
            self._unknown = int(np.count_nonzero(d == 6))
        return self._gc_count, self._unknown

    @property
    def gc(self):
        n = len(self)
        return self._counts()[0] / n if n else 0.0

    @property
    def unknown(self):
        return self._counts()[1]

    @property
    def gc_known(self):
        gc, unk = self._counts()
        return gc / (len(self) - unk) if len(self) > unk else 0.0

    def start_probability(self):
        """Start codon probability estimated from the GC content (lib.pyx:983-990)."""
        gc = self.gc_known
        p_atg = (1 - gc) * (1 - gc) * gc / 8
        p_gtg = gc * (1 - gc) * gc / 8
        p_ttg = (1 - gc) * (1 - gc) * gc / 8
        return p_atg + p_gtg + p_ttg

    def stop_probability(self):
        """Stop codon probability estimated from the GC content (lib.pyx:992-999)."""
        gc = self.gc_known
        p_tga = (1 - gc) * (1 - gc) * gc / 8.0
        p_tag = (1 - gc) * gc * (1 - gc) / 8.0
        p_taa = (1 - gc) * (1 - gc) * (1 - gc) / 8.0
        return p_tga + p_tag + p_taa

    def max_gc_frame_plot(self, window_size=120):
        """Frame with the highest GC content around every position (lib.pyx:1001-1026), computed on the GPU.

        Like the reference, the plot always uses a 120-base window: `window_size` is only validated."""
        if window_size < 0:
            raise ValueError(f"Invalid window size {window_size!r}")
        import array
        ctx = _context_for(_LazyBins.get()._blob_bytes(), 50)
        with ctx.lock:
            gp = ctx.max_gc_frame_plot(self._ascii)
        plot = array.array("i")
        plot.frombytes(gp.astype(np.intc).tobytes())
        return plot

    def shine_dalgarno(self, pos, start, training_info, strand=1, exact=True):
        """Bin of the highest scoring Shine-Dalgarno motif in the window at `pos` upstream of `start`
        (lib.pyx:1028-1072), evaluated on the GPU with the rbs weights of `training_info`."""
        if strand != 1 and strand != -1:
            raise ValueError(f"Invalid strand: {strand!r} (must be +1 or -1)")
        if pos < 0:
            raise ValueError("`pos` must be positive")
        if start < 0:
            raise ValueError("`start` must be positive")
        ctx = _context_for(bytes(training_info), 1)
        with ctx.lock:
            return ctx.shine_dalgarno(self._ascii, pos, start, 0, strand, exact)

    def __sizeof__(self):
        return len(self) + object.__sizeof__(self)

    def __getstate__(self):
        # the reference's keys (lib.pyx:616-630) plus what this mirror needs to rebuild itself
        return {"slen": len(self), "gc": self.gc, "masks": self.masks, "digits": bytearray(self.digits.tobytes()),
                "ascii": self._ascii.tobytes(), "mask": self._mask, "mask_size": self._mask_size}

    def __setstate__(self, state):
        if "ascii" in state:
            self._ascii = np.frombuffer(state["ascii"], dtype=np.uint8)
        else:  # a state pickled by the reference: digits only
            self._ascii = np.frombuffer(b"AGCTNNN", dtype=np.uint8)[np.frombuffer(bytes(state["digits"]), dtype=np.uint8)]
        self._digits = None
        self._mask, self._mask_size = state.get("mask", bool(len(state.get("masks", ())))), state.get("mask_size", 50)
        self._masks = state.get("masks")
        self._gc_count = self._unknown = None

    @property
    def masks(self):
        if self._masks is None:
            m = Masks()
            if self._mask:
                d = self.digits == 6
                if d.any():
                    edges = np.flatnonzero(np.diff(np.concatenate(([0], d.view(np.int8), [0]))))
                    for b, e in zip(edges[::2], edges[1::2]):
                        if e - b >= self._mask_size or e == len(d):
                            m.append(Mask(int(b), int(e)))
            self._masks = m
        return self._masks


# --- Nodes --------------------------------------------------------------------------------------------
class Node:
    """View on one node record (lib.pyx:1440-1552)."""

    __slots__ = ("owner", "_r")

    def __init__(self, owner, record):
        self.owner, self._r = owner, record

    index = property(lambda s: int(s._r["ndx"]))
    strand = property(lambda s: int(s._r["strand"]))
    type = property(lambda s: ["ATG", "GTG", "TTG", "Stop"][int(s._r["type"])])
    edge = property(lambda s: bool(s._r["edge"]))
    gc_bias = property(lambda s: 0)
    cscore = property(lambda s: float(s._r["cscore"]))
    gc_cont = property(lambda s: float(s._r["gc_cont"]))
    score = property(lambda s: float(s._r["score"]))
    rscore = property(lambda s: float(s._r["rscore"]))
    sscore = property(lambda s: float(s._r["sscore"]))
    tscore = property(lambda s: float(s._r["tscore"]))
    uscore = property(lambda s: float(s._r["uscore"]))

    def __repr__(self):
        return f"<pyrodigal_b200.Node index={self.index!r} strand={self.strand:+} type={self.type!r} edge={self.edge!r}>"


class Nodes(typing.Sequence):
    """Sorted node array of one contig as a numpy structured array (`_capi.NODE_DTYPE`).

    The operator-level methods of the reference (`extract`, `sort`, `reset_scores`, `score`, lib.pyx:2512-2595) are
    mirrored on top of the C ABI's operator twins (`pgpu_extract_nodes`, `pgpu_score_nodes`): the GPU extracts
    straight into sorted order, so `extract` already returns what the reference has after `extract` + `sort`."""

    def __init__(self, array=None):
        self.array = np.zeros(0, dtype=_capi.NODE_DTYPE) if array is None else array
        self._params = None     # how the nodes were extracted (pgpu_score_nodes re-extracts with the same options)
        self._scored = False    # nodes that were scored once carry converted edge flags

    def __len__(self):
        return len(self.array)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return Nodes(self.array[index])
        return Node(self, self.array[index])

    def copy(self):
        new = Nodes(self.array.copy())
        new._params, new._scored = self._params, self._scored
        return new

    __copy__ = copy

    def clear(self):
        """Remove all nodes from the node list (lib.pyx:2512-2516)."""
        self.array = np.zeros(0, dtype=_capi.NODE_DTYPE)
        self._params, self._scored = None, False

    def extract(self, sequence, *, closed=False, min_gene=90, min_edge_gene=60, translation_table=11):
        """Nodes.extract (lib.pyx:2518-2560, = add_nodes) on the GPU; returns the number of nodes added.

        The nodes arrive in (index, strand) order, i.e. as the reference has them after `sort()`."""
        if translation_table not in TRANSLATION_TABLES:
            raise ValueError(f"{translation_table} is not a valid translation table index")
        if min_gene <= 0 or min_edge_gene <= 0:
            raise ValueError("`min_gene` and `min_edge_gene` must be strictly positive")
        seq = sequence if isinstance(sequence, Sequence) else Sequence(sequence)
        opts = _capi.make_opts(meta=False, closed=closed, mask=seq._mask, min_mask=seq._mask_size, min_gene=min_gene,
                               min_edge_gene=min_edge_gene, max_overlap=min(60, min_gene))
        ctx = _context_for(_LazyBins.get()._blob_bytes(), 50)  # extraction needs no model; any context will do
        with ctx.lock:
            out = ctx.extract_nodes(seq._ascii, translation_table, opts)
        n = len(out["ndx"])
        new = np.zeros(n, dtype=_capi.NODE_DTYPE)
        for name in ("ndx", "stop_val", "strand", "type", "edge"):
            new[name] = out[name]
        first = len(self.array) == 0
        self.array = new if first else np.concatenate([self.array, new])
        self._params = dict(closed=bool(closed), min_gene=min_gene, min_edge_gene=min_edge_gene,
                            translation_table=translation_table) if first else None
        self._scored = False
        return n

    def _is_sorted(self):
        a = self.array
        if len(a) < 2:
            return True
        d = np.diff(a["ndx"].astype(np.int64))
        # compare_nodes (node.c:1578-1587): by index, forward strand first
        return bool(np.all((d > 0) | ((d == 0) & (a["strand"][:-1] >= a["strand"][1:]))))

    def sort(self):
        """Nodes.sort (lib.pyx:2591-2595).  Extraction on the GPU already emits the sorted order, so this only checks."""
        if not self._is_sorted():
            raise NotImplementedError("nodes of several extractions were concatenated: clear() before extract()")

    def reset_scores(self):
        """Nodes.reset_scores (node.c:176-197): scores to zero, trace pointers to -1; the edge flags are kept."""
        a = self.array
        for name in ("score", "cscore", "sscore", "rscore", "tscore", "uscore", "mot_score"):
            a[name] = 0.0
        for name in ("rbs", "mot_len", "mot_ndx", "mot_spacer", "mot_spacendx", "star_ptr", "elim"):
            a[name] = 0
        a["traceb"] = a["tracef"] = -1
        a["ov_mark"] = -1

    def score(self, sequence, training_info, *, closed=False, is_meta=False):
        """Nodes.score (lib.pyx:2569-2589, = score_nodes) on the GPU for the nodes of `sequence` held by this object."""
        seq = sequence if isinstance(sequence, Sequence) else Sequence(sequence)
        p = self._params or {}
        opts = _capi.make_opts(meta=False, single_model=0, closed=closed, mask=seq._mask, min_mask=seq._mask_size,
                               min_gene=p.get("min_gene", 90), min_edge_gene=p.get("min_edge_gene", 60),
                               max_overlap=min(60, p.get("min_gene", 90)))
        ctx = _context_for(bytes(training_info), 1)
        with ctx.lock:
            scored = ctx.score_nodes(seq._ascii, 0, opts, is_meta=is_meta, first_pass=not self._scored)
        a = self.array
        if (len(scored) != len(a) or not np.array_equal(scored["ndx"], a["ndx"])
                or not np.array_equal(scored["strand"], a["strand"]) or not np.array_equal(scored["type"], a["type"])):
            raise ValueError("the nodes were not extracted from this sequence with these options "
                             "(translation table, `closed`, minimum gene lengths)")
        for name in ("cscore", "sscore", "rscore", "tscore", "uscore", "rbs", "mot_len", "mot_ndx", "mot_spacer",
                     "mot_spacendx", "mot_score", "gc_cont", "edge"):
            a[name] = scored[name]
        self._scored = True

    def __getstate__(self):
        return {"array": self.array, "params": self._params, "scored": self._scored}

    def __setstate__(self, state):
        self.array = state["array"]
        self._params, self._scored = state.get("params"), state.get("scored", False)


def _codon_masks(tt):
    """(is_stop[64], is_start[64]) over the 6-bit codon index (x0 << 4) + (x1 << 2) + x2: the scanner rules of
    _sequence.h:45-73 (starts) and 117-157 (stops)"""
    m = _MASK_CACHE.get(tt)
    if m is None:
        A, G, C, T = 0, 1, 2, 3
        taa = tt in (1, 2, 3, 4, 5, 9, 10, 11, 12, 13, 15, 16, 21, 22, 23, 24, 25, 26, 32)
        tag = tt in (1, 2, 3, 4, 5, 9, 10, 11, 12, 13, 14, 21, 23, 24, 25, 26, 33)
        tga = tt in (1, 6, 11, 12, 15, 16, 22, 23, 26, 29, 30, 32)
        stop, start = np.zeros(64, dtype=bool), np.zeros(64, dtype=bool)
        for x0 in range(4):
            for x1 in range(4):
                for x2 in range(4):
                    k = (x0 << 4) + (x1 << 2) + x2
                    if (x0, x1, x2) == (T, A, G):
                        st = tag
                    elif (x0, x1, x2) == (T, G, A):
                        st = tga
                    elif (x0, x1, x2) == (T, A, A):
                        st = taa
                    elif tt == 2:
                        st = x0 == A and x1 == G and x2 in (A, G)
                    elif tt == 22:
                        st = (x0, x1, x2) == (T, C, A)
                    elif tt == 23:
                        st = (x0, x1, x2) == (T, T, A)
                    else:
                        st = False
                    sa = False
                    if x1 == T and x2 == G:
                        if x0 == A:
                            sa = True
                        elif tt in (6, 10, 14, 15, 16, 2):
                            sa = False
                        elif x0 == G:
                            sa = tt not in (1, 3, 12, 2)
                        elif x0 == T:
                            sa = not (tt < 4 or tt == 9 or 21 <= tt < 25)
                    stop[k], start[k] = bool(st), bool(sa)
        m = _MASK_CACHE[tt] = (stop, start)
    return m


_MASK_CACHE = {}


# --- Genes ----------------------------------------------------------------------------------------------
def calculate_confidence(score, start_weight):
    """vendor/Prodigal/gene.c:522-532"""
    if score / start_weight < 41:
        conf = math.exp(score / start_weight)
        conf = (conf / (conf + 1)) * 100.0
    else:
        conf = 99.99
    return 50.0 if conf <= 50.0 else conf


class Gene:
    """One predicted gene (lib.pyx:2610-3047); reads the start/stop node records copied back with it."""

    __slots__ = ("owner", "_g", "_start", "_stop", "_i")

    def __init__(self, owner, i):
        self.owner, self._i = owner, i
        self._g = owner._genes[i]
        self._start, self._stop = owner._gene_nodes[i, 0], owner._gene_nodes[i, 1]

    begin = property(lambda s: int(s._g["begin"]))
    end = property(lambda s: int(s._g["end"]))
    strand = property(lambda s: int(s._start["strand"]))
    gc_cont = property(lambda s: float(s._start["gc_cont"]))
    cscore = property(lambda s: float(s._start["cscore"]))
    rscore = property(lambda s: float(s._start["rscore"]))
    sscore = property(lambda s: float(s._start["sscore"]))
    tscore = property(lambda s: float(s._start["tscore"]))
    uscore = property(lambda s: float(s._start["uscore"]))
    score = property(lambda s: float(s._start["cscore"]) + float(s._start["sscore"]))
    translation_table = property(lambda s: s.owner.training_info.translation_table)
    start_node = property(lambda s: Node(None, s._start))
    stop_node = property(lambda s: Node(None, s._stop))

    @property
    def partial_begin(self):
        return bool((self._start if self.strand == 1 else self._stop)["edge"])

    @property
    def partial_end(self):
        return bool((self._stop if self.strand == 1 else self._start)["edge"])

    @property
    def start_type(self):
        return _NODE_TYPE[3 if self._start["edge"] else int(self._start["type"])]

    def _rbs_pick(self, table):
        # lib.pyx:2694-2720 / 2723-2751
        t, n = self.owner.training_info, self._start
        rbs1 = float(t.rbs_weights[int(n["rbs"][0])]) * t.start_weight
        rbs2 = float(t.rbs_weights[int(n["rbs"][1])]) * t.start_weight
        if t.uses_sd:
            return table[int(n["rbs"][0 if rbs1 > rbs2 else 1])], False
        mot = float(n["mot_score"]) * t.start_weight
        if t.missing_motif_weight > -0.5 and rbs1 > rbs2 and rbs1 > mot:
            return table[int(n["rbs"][0])], False
        if t.missing_motif_weight > -0.5 and rbs2 >= rbs1 and rbs2 > mot:
            return table[int(n["rbs"][1])], False
        if int(n["mot_len"]) == 0:
            return None, False
        return None, True

    @property
    def rbs_motif(self):
        val, use_motif = self._rbs_pick(_RBS_MOTIF)
        if not use_motif:
            return val
        ndx, ln = int(self._start["mot_ndx"]), int(self._start["mot_len"])
        return "".join("AGCT"[(ndx >> (2 * i)) & 3] for i in range(ln))  # sequence.c:624-636

    @property
    def rbs_spacer(self):
        val, use_motif = self._rbs_pick(_RBS_SPACER)
        return f"{int(self._start['mot_spacer'])}bp" if use_motif else val

    def confidence(self):
        return calculate_confidence(self.score, self.owner.training_info.start_weight)

    def _gene_data(self, sequence_id):
        return "ID={}_{};partial={}{};start_type={};rbs_motif={};rbs_spacer={};gc_cont={:.3f}".format(
            sequence_id, self._i + 1, int(self.partial_begin), int(self.partial_end), self.start_type, self.rbs_motif,
            self.rbs_spacer, self.gc_cont)

    def _score_data(self):
        return "conf={:.2f};score={:.2f};cscore={:.2f};sscore={:.2f};rscore={:.2f};uscore={:.2f};tscore={:.2f};".format(
            self.confidence(), self.score, self.cscore, self.sscore, self.rscore, self.uscore, self.tscore)

    def sequence(self):
        d = self.owner.sequence.digits
        if self.strand == 1:
            return "".join(_LETTERS[x] for x in d[self.begin - 1:self.end])
        seg = d[self.begin - 1:self.end][::-1]
        return "".join(_LETTERS[x ^ 3 if x < 4 else x] for x in seg)

    def translate(self, translation_table=None, unknown_residue="X", include_stop=True, strict=True):
        """Translate the predicted gene into a protein sequence (lib.pyx:2926-3047, _sequence.h:75-157)."""
        owner_table = self.owner.training_info.translation_table
        if translation_table is None:
            tt = owner_table
        elif translation_table not in _NCBI_CHANGES:
            raise ValueError(f"{translation_table} is not a valid translation table index")
        else:
            if _stop_codons(translation_table) != _stop_codons(owner_table):
                warnings.warn(
                    f"requested translation table ({translation_table!r}) has different STOP codons "
                    f"than the one these genes were called with ({owner_table!r}), consider calling "
                    "genes with the proper translation table instead. This may become an error "
                    "in the future.", stacklevel=2)
            tt = translation_table
        d = self.owner.sequence.digits
        seg = d[self.begin - 1:self.end]
        if self.strand != 1:
            seg = seg[::-1] ^ 3          # the reference complements with xor 3 here, unknown bases included
        start_edge, stop_edge = bool(self._start["edge"]), bool(self._stop["edge"])
        ncod = len(seg) // 3
        if not stop_edge and not include_stop:
            ncod -= 1
        if ncod <= 0:
            return ""
        cod = seg[:3 * ncod].reshape(ncod, 3).astype(np.int64)
        known = (cod <= 3).all(axis=1)
        tab = _translation_table(tt)
        idx = np.where(known, (cod[:, 0] << 4) + (cod[:, 1] << 2) + cod[:, 2], 0)
        unk = ord(unknown_residue) if isinstance(unknown_residue, str) else int(unknown_residue)
        aa = np.where(known, tab[idx], unk).astype(np.uint8)
        # the stop / start scanners of _sequence.h override the table: table-specific stops and the initial M
        stop_m, start_m = _codon_masks(tt)
        aa[known & stop_m[idx]] = ord("*")
        if not start_edge and known[0] and start_m[idx[0]] and not stop_m[idx[0]]:
            aa[0] = ord("M")
        if not strict:
            # third base unknown: translate when all four completions agree (_sequence.h:95-102); an unknown
            # middle base never resolves in the reference (its loop at 104-111 does not advance), nor does the first
            for k in np.flatnonzero(~known):
                x0, x1, x2 = (int(v) for v in cod[k])
                if x0 <= 3 and x1 <= 3 and x2 > 3:
                    opts = {int(tab[(x0 << 4) + (x1 << 2) + b]) for b in range(4)}
                    if len(opts) == 1:
                        v = opts.pop()
                        aa[k] = unk if v == ord("X") else v
        return aa.tobytes().decode("ascii")

    def __repr__(self):
        return f"<pyrodigal_b200.Gene begin={self.begin} end={self.end} strand={self.strand:+}>"


class Genes(typing.Sequence):
    """Predicted genes of one contig (lib.pyx:3049-3184)."""

    def __init__(self, genes, gene_nodes, *, sequence, training_info, metagenomic_bin, meta, nodes, ipath, num_seq=1):
        self._genes, self._gene_nodes = genes, gene_nodes
        self.sequence, self.training_info, self.metagenomic_bin = sequence, training_info, metagenomic_bin
        self.meta, self._nodes, self.ipath, self._num_seq = meta, nodes, ipath, num_seq

    def __len__(self):
        return len(self._genes)

    def __bool__(self):
        return len(self._genes) > 0

    def __getitem__(self, index):
        if isinstance(index, slice):
            return [Gene(self, i) for i in range(*index.indices(len(self)))]
        if index < 0:
            index += len(self)
        if index < 0 or index >= len(self):
            raise IndexError("genes index out of range")
        return Gene(self, index)

    def clear(self):
        """Remove all genes from the list (lib.pyx:3185-3191)."""
        self._genes = self._genes[:0]
        self._gene_nodes = self._gene_nodes[:0]

    @property
    def nodes(self):
        if self._nodes is None:
            raise RuntimeError("node arrays were not requested: use GeneFinder.find_genes or want_nodes=True")
        return self._nodes

    @property
    def score(self):
        """lib.pyx:3170-3184: nodes[ipath].score (0.0 after the meta-mode re-scoring)"""
        if self.ipath < 0 or self._nodes is None:
            return 0.0
        return float(self._nodes.array["score"][self.ipath])

    def __getstate__(self):
        return self.__dict__.copy()

    def __setstate__(self, state):
        self.__dict__.update(state)

    # ---- writers (lib.pyx:3405-3895): host-side formatting of the records the GPU path returned ----
    def _model(self):
        """(training info, model description) used in headers; the reference falls back to bin #5 when meta mode
        found no gene (lib.pyx:3583-3590)"""
        tinf, mbin = self.training_info, self.metagenomic_bin
        if self.meta:
            if mbin is None:
                mbin = _LazyBins.get()[5]
            if tinf is None:
                tinf = mbin.training_info
            return tinf, mbin.description
        return tinf, "Ab initio"

    def write_gff(self, file, sequence_id, header=True, include_translation_table=False, full_id=True,
                  version_separator="_v"):
        """GFF3 (lib.pyx:3529-3645)"""
        tinf, desc = self._model()
        run = "Metagenomic" if self.meta else "Single"
        n = 0
        if header:
            n += file.write("##gff-version  3\n")
        n += file.write(f'# Sequence Data: seqnum={self._num_seq};seqlen={len(self.sequence)};seqhdr="{sequence_id}"\n')
        n += file.write(f'# Model Data: version=pyrodigal.v{PYRODIGAL_COMPAT_VERSION};run_type={run};model="{desc}";'
                        f"gc_cont={tinf.gc * 100:.2f};transl_table={tinf.translation_table};uses_sd={int(tinf.uses_sd)}\n")
        for gene in self:
            fields = [sequence_id, f"pyrodigal{version_separator}{PYRODIGAL_COMPAT_VERSION}", "CDS", str(gene.begin),
                      str(gene.end), "{:.1f}".format(gene.sscore + gene.cscore), "+" if gene.strand > 0 else "-", "0"]
            attr = gene._gene_data(sequence_id if full_id else self._num_seq) + ";"
            if include_translation_table:
                attr += f"transl_table={tinf.translation_table};"
            n += file.write("\t".join(fields) + "\t" + attr + gene._score_data() + "\n")
        return n

    def _fasta_header(self, i, gene, sequence_id, full_id):
        return ">{}_{} # {} # {} # {} # {}\n".format(sequence_id, i + 1, gene.begin, gene.end, gene.strand,
                                                      gene._gene_data(sequence_id if full_id else self._num_seq))

    def write_genes(self, file, sequence_id, width=70, full_id=False):
        """nucleotide FASTA of the genes (lib.pyx:3647-3702)"""
        n = 0
        for i, gene in enumerate(self):
            n += file.write(self._fasta_header(i, gene, sequence_id, full_id))
            for line in textwrap.wrap(gene.sequence(), width=width):
                n += file.write(line + "\n")
        return n

    def write_translations(self, file, sequence_id, width=60, translation_table=None, include_stop=True,
                           strict_translation=True, full_id=False):
        """protein FASTA (lib.pyx:3704-3781)"""
        if translation_table is not None and translation_table not in TRANSLATION_TABLES:
            raise ValueError(f"{translation_table} is not a valid translation table index")
        n = 0
        for i, gene in enumerate(self):
            n += file.write(self._fasta_header(i, gene, sequence_id, full_id))
            trans = gene.translate(translation_table, include_stop=include_stop, strict=strict_translation)
            for line in textwrap.wrap(trans, width=width):
                n += file.write(line + "\n")
        return n

    def write_genbank(self, file, sequence_id, division="BCT", date=None, translation_table=None, strict_translation=True):
        """complete GenBank record (lib.pyx:3405-3527)"""
        if translation_table is None:
            if self.training_info is not None:
                translation_table = self.training_info.translation_table
        elif translation_table not in TRANSLATION_TABLES:
            raise ValueError(f"{translation_table} is not a valid translation table index")
        if date is None:
            date = datetime.date.today()
        elif not isinstance(date, datetime.date):
            raise TypeError(f"Expected datetime.date, found {type(date).__name__}")
        slen = len(self.sequence)
        n = file.write("LOCUS       {:<23} {} bp    DNA     linear   {} {}\n".format(
            sequence_id, slen, division, date.strftime("%d-%b-%y").upper()))
        n += file.write(
            f"REFERENCE   1  (bases 1 to {slen})\n"
            "  AUTHORS   Hyatt,D., Chen,G-L., LoCascio,P.F., Land,M.L., Larimer,F.W.\n"
            "            Hauser,L.J.\n"
            "  TITLE     Prodigal: prokaryotic gene recognition and translation initiation\n"
            "            site identification\n"
            "  JOURNAL   BMC Bioinformatics. 2010;11:119.\n"
            "   PUBMED   20211023\n"
            f"REFERENCE   2  (bases 1 to {slen})\n"
            "  AUTHORS   Larralde,M.\n"
            "  TITLE     Pyrodigal: Python bindings and interface to Prodigal, an efficient\n"
            "            method for gene prediction in prokaryotes\n"
            "  JOURNAL   Journal of Open Source Software, 7(72), 4296.\n"
            "FEATURES             Location/Qualifiers\n")
        pad = " " * 21
        for i, gene in enumerate(self):
            begin = f"<{gene.begin}" if gene.start_node.edge else f"{gene.begin}"
            end = f">{gene.end}" if gene.stop_node.edge else f"{gene.end}"
            loc = f"{begin}..{end}" if gene.strand == 1 else f"complement({begin}..{end})"
            n += file.write(f"     CDS             {loc}\n{pad}/codon_start=1\n"
                            f'{pad}/inference="ab initio prediction:pyrodigal:{PYRODIGAL_COMPAT_VERSION}"\n'
                            f'{pad}/locus_tag="{sequence_id}_{i + 1}"\n{pad}/transl_table={translation_table}\n')
            translation = '/translation="{}"'.format(
                gene.translate(translation_table=translation_table, include_stop=False, strict=strict_translation))
            for block in textwrap.wrap(translation, 59):
                n += file.write(pad + block + "\n")
        seq = str(self.sequence).lower()
        n += file.write("ORIGIN\n")
        for i in range(0, len(seq), 60):
            n += file.write("{:>9}".format(i + 1))
            for j in range(i, min(i + 60, len(seq)), 10):
                n += file.write(" " + seq[j:j + 10])
            n += file.write("\n")
        n += file.write("//\n")
        return n

    def write_scores(self, file, sequence_id, header=True):
        """every potential start with its scores, grouped by stop codon (lib.pyx:3783-3895)"""
        tinf = _LazyBins.get()[5].training_info if (self.meta and self.training_info is None) else self.training_info
        n = 0
        if header:
            n += file.write(f'# Sequence Data: seqnum={self._num_seq};seqlen={len(self.sequence)};seqhdr="{sequence_id}"\n')
            n += file.write(f"# Run Data: version=pyrodigal.v{PYRODIGAL_COMPAT_VERSION};gc_cont={tinf.gc * 100:.2f};"
                            f"transl_table={tinf.translation_table};uses_sd={int(tinf.uses_sd)}\n")
            n += file.write("Beg\tEnd\tStd\tTotal\tCodPot\tStrtSc\tCodon\tRBSMot\tSpacer\tRBSScr\tUpsScr\tTypeScr\tGCCont\n")
        a = self.nodes.array
        # stopcmp_nodes (vendor/Prodigal/node.c:1591-1602): stop position, forward strand first, then position
        order = np.lexsort((a["ndx"], -a["strand"].astype(np.int64), a["stop_val"]))
        st_wt, rbs_wt = tinf.start_weight, tinf.rbs_weights
        prev = None
        for k in order:
            nd = a[k]
            if nd["type"] == 3:
                continue
            key = (int(nd["stop_val"]), int(nd["strand"]))
            if key != prev:
                prev = key
                n += file.write("\n")
            if nd["strand"] == 1:
                n += file.write(f"{int(nd['ndx']) + 1:d}\t{int(nd['stop_val']) + 3:d}\t+\t")
            else:
                n += file.write(f"{int(nd['stop_val']) - 1:d}\t{int(nd['ndx']) + 1:d}\t-\t")
            cs, ss, rs = float(nd["cscore"]), float(nd["sscore"]), float(nd["rscore"])
            n += file.write(f"{cs + ss:.2f}\t{cs:.2f}\t{ss:.2f}\t{_NODE_TYPE[3 if nd['edge'] else int(nd['type'])]}\t")
            r0, r1 = int(nd["rbs"][0]), int(nd["rbs"][1])
            rbs1, rbs2 = float(rbs_wt[r0]) * st_wt, float(rbs_wt[r1]) * st_wt
            mot = float(nd["mot_score"]) * st_wt
            if tinf.uses_sd:
                pick = r0 if rbs1 > rbs2 else r1
                n += file.write(f"{_RBS_MOTIF[pick]}\t{_RBS_SPACER[pick]}\t{rs:.2f}\t")
            elif tinf.missing_motif_weight > -0.5 and rbs1 > rbs2 and rbs1 > mot:
                n += file.write(f"{_RBS_MOTIF[r0]}\t{_RBS_SPACER[r0]}\t{rs:.2f}\t")
            elif tinf.missing_motif_weight > -0.5 and rbs2 >= rbs1 and rbs2 > mot:
                n += file.write(f"{_RBS_MOTIF[r1]}\t{_RBS_SPACER[r1]}\t{rs:.2f}\t")
            elif int(nd["mot_len"]) == 0:
                n += file.write(f"None\tNone\t{rs:.2f}\t")
            else:
                ndx, ln = int(nd["mot_ndx"]), int(nd["mot_len"])
                word = "".join("AGCT"[(ndx >> (2 * i)) & 3] for i in range(ln))
                n += file.write(f"{word}\t{int(nd['mot_spacer']):d}bp\t{rs:.2f}\t")
            n += file.write(f"{float(nd['uscore']):.2f}\t{float(nd['tscore']):.2f}\t{float(nd['gc_cont']):.3f}\n")
        n += file.write("\n")
        return n


# --- ConnectionScorer (operator-level surface, lib.pyx:1086-1432) -------------------------------------------
class ConnectionScorer:
    """GPU replacement of the SIMD-filter + score_connection plug-in host.

    `backend` is accepted for signature compatibility; there is exactly one backend (CUDA, sm_100a)."""

    def __init__(self, backend="detect", device=0):
        if backend not in ("detect", "cuda", "generic", "sse", "avx", "avx512", "neon", "swar64", None):
            raise ValueError(f"Unsupported backend: {backend}")
        self.backend = "cuda"
        self.device = device
        self._idx = None

    def index(self, nodes):
        a = nodes.array
        self._idx = (a["strand"].copy(), a["type"].copy(), a["ndx"].copy())
        self.skip_connection = np.zeros(len(a), dtype=np.uint8)

    def compute_skippable(self, min, i):
        assert self._idx is not None and min <= i
        blob = _LazyBins.get()._blob_bytes()
        ctx = _context_for(blob, 50, self.device)
        with ctx.lock:
            skip = ctx.compute_skippable(*self._idx, min, i)
        self.skip_connection[min:i] = skip[min:i]

    def score_connections(self, nodes, tinf, final=False, gc_score=None):
        blob = bytes(tinf)
        ctx = _context_for(blob, 1, self.device)
        a = nodes.array
        if gc_score is None:
            gc_score = np.zeros((len(a), 3))
        with ctx.lock:
            score, traceb, ov, pairs, ms = ctx.score_connections(
                a["ndx"], a["stop_val"], a["strand"], a["type"], a["cscore"], a["sscore"], a["rscore"], a["uscore"],
                gc_score, a["star_ptr"], 0, final)
        a["score"], a["traceb"], a["ov_mark"] = score, traceb, ov
        a["tracef"] = -1
        self.last_pairs, self.last_ms = pairs, ms


# --- GeneFinder ------------------------------------------------------------------------------------------
class GeneFinder:
    """Drop-in for pyrodigal.GeneFinder on the find_genes path (lib.pyx:5073-5575)."""

    def __init__(self, training_info=None, *, meta=False, metagenomic_bins=None, closed=False, mask=False, min_mask=50,
                 min_gene=90, min_edge_gene=60, max_overlap=60, backend="detect", device=0):
        if meta and training_info is not None:
            raise ValueError("cannot use a training info in meta mode.")
        if min_gene <= 0:
            raise ValueError("`min_gene` must be strictly positive")
        if min_edge_gene <= 0:
            raise ValueError("`min_edge_gene` must be strictly positive")
        if min_mask < 0:
            raise ValueError("`min_mask` must be positive")
        if max_overlap < 0:
            raise ValueError("`max_overlap` must be positive")
        elif max_overlap > min_gene:
            raise ValueError("`max_overlap` must be lower than `min_gene`")
        self.meta, self.closed, self.mask = meta, closed, mask
        self.training_info = training_info
        self.min_mask, self.min_gene, self.min_edge_gene, self.max_overlap = min_mask, min_gene, min_edge_gene, max_overlap
        self.backend = backend
        self.device = device
        self.metagenomic_bins = _LazyBins.get() if metagenomic_bins is None else metagenomic_bins
        self.lock = threading.Lock()
        self._num_seq = 1

    def __repr__(self):
        return f"pyrodigal_b200.GeneFinder(meta={self.meta!r}, closed={self.closed!r}, mask={self.mask!r})"

    def __reduce__(self):
        import functools
        fn = functools.partial(type(self), meta=self.meta, metagenomic_bins=self.metagenomic_bins, closed=self.closed,
                               mask=self.mask, min_mask=self.min_mask, min_gene=self.min_gene,
                               min_edge_gene=self.min_edge_gene, max_overlap=self.max_overlap, backend=self.backend)
        return fn, (self.training_info,)

    # ---- internals ----
    def _context(self):
        if self.meta:
            bins = self.metagenomic_bins
            return _context_for(bins._blob_bytes(), len(bins), self.device)
        return _context_for(bytes(self.training_info), 1, self.device)

    def _opts(self, want_nodes):
        return _capi.make_opts(meta=self.meta, single_model=0, closed=self.closed, mask=self.mask,
                               min_mask=self.min_mask, min_gene=self.min_gene, min_edge_gene=self.min_edge_gene,
                               max_overlap=self.max_overlap, want_nodes=want_nodes)

    def _wrap(self, res, k, seq, want_nodes, num_seq):
        s = res.summary[k]
        a, b = int(res.gene_off[k]), int(res.gene_off[k + 1])
        if self.meta:
            w = int(s["winner"])
            mbin = self.metagenomic_bins[w] if w >= 0 else None
            tinf = mbin.training_info if mbin is not None else None
        else:
            mbin, tinf = None, self.training_info
        if not isinstance(seq, Sequence):
            seq = Sequence(seq, mask=self.mask, mask_size=self.min_mask)
        seq._gc_count, seq._unknown = int(s["gc_count"]), int(s["unknown"])
        nodes = Nodes(res.nodes(k)) if want_nodes else None
        return Genes(res.genes[a:b], res.gene_nodes[a:b], sequence=seq, training_info=tinf, metagenomic_bin=mbin,
                     meta=self.meta, nodes=nodes, ipath=int(s["ipath"]), num_seq=num_seq)

    # ---- public ----
    class _Batch(collections.abc.Sequence):
        """What `find_genes_many` / `find_genes_batch` return: a read-only sequence of `Genes`, one per contig, built on
        first access from the result buffers of the batched call (zero-copy views of the library's page-locked result
        memory).  Creating twelve thousand `Genes` / `Sequence` objects eagerly cost more host time than the GPU pass."""

        def __init__(self, finder, res, sequences, want_nodes, first):
            self._finder, self._res, self._sequences, self._want_nodes, self._first = finder, res, sequences, want_nodes, first
            self._cache = {}

        def __len__(self):
            return self._res.n

        def __getitem__(self, k):
            if isinstance(k, slice):
                return [self[i] for i in range(*k.indices(len(self)))]
            if k < 0:
                k += len(self)
            if not 0 <= k < len(self):
                raise IndexError("batch index out of range")
            g = self._cache.get(k)
            if g is None:
                g = self._cache[k] = self._finder._wrap(self._res, k, self._sequences(k), self._want_nodes, self._first + k)
            return g

    def find_genes(self, sequence):
        """Find all the genes in one input sequence (lib.pyx:5400-5469)."""
        return self.find_genes_many([sequence], want_nodes=True)[0]

    def find_genes_many(self, sequences, want_nodes=False):
        """Batched find_genes: one GPU pass over all the given contigs (the reference maps find_genes over
        records with a thread pool, cli.py:286-300; one contig cannot fill a B200)."""
        arrays = [_as_ascii(s) for s in sequences]
        n = len(arrays)
        offsets = np.zeros(n + 1, dtype=np.int64)
        if n:
            np.cumsum([len(a) for a in arrays], out=offsets[1:])
        flat = np.concatenate(arrays) if n > 1 else (arrays[0] if n else np.zeros(0, np.uint8))
        return self.find_genes_batch(flat, offsets, want_nodes=want_nodes, sequences=sequences)

    def find_genes_batch(self, flat, offsets, want_nodes=False, sequences=None):
        """find_genes over contigs that already sit back to back in one uint8 buffer (`fasta.read_batch`):
        contig k is flat[offsets[k]:offsets[k + 1]].  This is the layout of the C ABI, passed through unchanged."""
        if not self.meta and self.training_info is None:
            raise RuntimeError("cannot find genes without having trained in single mode")
        n = len(offsets) - 1
        flat = np.ascontiguousarray(flat, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        with self.lock:
            first = self._num_seq
            self._num_seq += n
        ctx = self._context()
        with ctx.lock:
            res = ctx.find_genes_batch(flat, offsets, self._opts(want_nodes))
        self.last_stats = res.stats
        if sequences is None:
            get = lambda k: flat[offsets[k]:offsets[k + 1]]
        else:
            get = sequences.__getitem__
        return GeneFinder._Batch(self, res, get, want_nodes, first)

    def train(self, sequence, *sequences, force_nonsd=False, start_weight=4.35, translation_table=11):
        """Search parameters for the ORF finder using a training sequence (lib.pyx:5471-5575).

        Several sequences are treated as contigs of one genome and joined by ``TTAATTAATTAA`` linkers, like the
        reference does.  The whole training pass (node extraction, GC frame bias, training DP, dicodon statistics,
        SD / non-SD start training) runs on the GPU through ``pgpu_train``; the returned `TrainingInfo` is
        byte-identical to the reference's and becomes ``self.training_info``."""
        if self.meta:
            raise RuntimeError("cannot use training sequence in metagenomic mode")
        if translation_table not in TRANSLATION_TABLES:
            raise ValueError(f"{translation_table} is not a valid translation table index")
        if isinstance(sequence, Sequence):
            if sequences:
                raise NotImplementedError("cannot use more than one `Sequence` object in `GeneFinder.train`")
            ascii_ = sequence._ascii
        elif isinstance(sequence, str):
            if sequences:
                sequence = "TTAATTAATTAA".join(itertools.chain([sequence], sequences, [""]))
            ascii_ = _as_ascii(sequence)
        else:
            if sequences:
                sequence = b"TTAATTAATTAA".join(bytes(memoryview(x)) for x in itertools.chain([sequence], sequences, [b""]))
            ascii_ = _as_ascii(sequence)
        slen = len(ascii_)
        if slen < MIN_SINGLE_GENOME:
            raise ValueError(f"sequence must be at least {MIN_SINGLE_GENOME} characters ({slen} found)")
        elif slen < IDEAL_SINGLE_GENOME:
            warnings.warn(f"sequence should be at least {IDEAL_SINGLE_GENOME} characters ({slen} found)")
        ctx = _training_context(self.device)
        with ctx.lock:
            blob, self.last_train_stats = ctx.train(np.ascontiguousarray(ascii_), self._opts(False),
                                                    translation_table=translation_table, start_weight=start_weight,
                                                    force_nonsd=force_nonsd)
        tinf = TrainingInfo._from_bytes(blob)
        with self.lock:
            self.training_info = tinf
        return tinf
