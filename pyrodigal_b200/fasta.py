"""FASTA ingest for the batched GPU path.

The reference reads records one by one through a pure-Python line loop (src/pyrodigal/tests/fasta.py:61-86, used by
cli.py:283-284) and hands each sequence to `find_genes` separately.  The GPU path wants the opposite shape: ONE
contiguous uint8 buffer holding every sequence of the file plus an offsets table, which is exactly what
`pgpu_find_genes_batch` takes.  `read_batch` therefore parses the whole file with vectorised byte operations
(numpy): header lines are located from the positions of '>' at line starts, newlines / carriage returns are
dropped with one boolean mask, and the per-record offsets follow from a cumulative sum -- no per-line Python loop.
Compressed inputs (gzip / bz2 / xz, and lz4 / zstd when their modules are installed) are detected by magic number
like the reference's `zopen` (cli.py:31-62)."""
import collections
import io
import os

import numpy as np

_MAGIC = ((b"\x1f\x8b", "gzip"), (b"BZh", "bz2"), (b"\xfd7zXZ", "lzma"), (b"\x04\x22\x4d\x18", "lz4"),
          (b"\x28\xb5\x2f\xfd", "zstd"))

Record = collections.namedtuple("Record", ["id", "seq", "description"])


def _read_bytes(source):
    """whole content of a path / binary file / text file / bytes as bytes, decompressed if needed"""
    if isinstance(source, (bytes, bytearray, memoryview)):
        data = bytes(source)
    elif isinstance(source, (str, os.PathLike)):
        with open(source, "rb") as f:
            data = f.read()
    else:
        data = source.read()
        if isinstance(data, str):
            data = data.encode("ascii", "replace")
    for magic, kind in _MAGIC:
        if data.startswith(magic):
            if kind == "gzip":
                import gzip
                return gzip.decompress(data)
            if kind == "bz2":
                import bz2
                return bz2.decompress(data)
            if kind == "lzma":
                import lzma
                return lzma.decompress(data)
            if kind == "lz4":
                try:
                    import lz4.frame
                except ImportError as err:
                    raise RuntimeError("File compression is LZ4 but lz4 is not installed") from err
                return lz4.frame.decompress(data)
            try:
                import zstandard
            except ImportError as err:
                raise RuntimeError("File compression is ZSTD but zstandard is not installed") from err
            return zstandard.ZstdDecompressor().stream_reader(io.BytesIO(data)).read()
    return data


class Batch:
    """All records of a FASTA file in the layout of the C ABI: `flat` (uint8 ASCII, sequences back to back),
    `offsets` (int64[n + 1]), `ids`, `descriptions`."""

    def __init__(self, flat, offsets, ids, descriptions):
        self.flat, self.offsets, self.ids, self.descriptions = flat, offsets, ids, descriptions

    def __len__(self):
        return len(self.ids)

    def sequence(self, k):
        return self.flat[self.offsets[k]:self.offsets[k + 1]]

    def records(self):
        for k in range(len(self)):
            yield Record(self.ids[k], self.sequence(k).tobytes().decode("ascii"), self.descriptions[k])


def _sequence_buffer(n, pinned):
    """where the sequence bytes go: page-locked memory of the CUDA library when asked for (asynchronous input copies
    that overlap the kernels of the previous sub-batch), ordinary memory otherwise"""
    if pinned:
        try:
            from . import _capi
            return _capi.pinned_empty(n)
        except Exception:
            pass
    return np.empty(n, dtype=np.uint8)


def read_batch(source, pinned=False):
    """Parse a FASTA file into one contiguous buffer + offsets (see module docstring).  `pinned=True` writes the sequence
    bytes straight into page-locked host memory (the only copy of them that is made)."""
    a = np.frombuffer(_read_bytes(source), dtype=np.uint8)
    n = len(a)
    if n == 0:
        return Batch(np.zeros(0, np.uint8), np.zeros(1, np.int64), [], [])
    nl = a == 10
    line_start = np.empty(n, dtype=bool)
    line_start[0] = True
    line_start[1:] = nl[:-1]
    hdr_pos = np.flatnonzero(line_start & (a == 62))  # '>' at the beginning of a line
    eol = np.flatnonzero(nl)
    # end of every header line (exclusive): the next newline, or the end of the file
    k = np.searchsorted(eol, hdr_pos)
    hdr_end = np.where(k < len(eol), eol[np.minimum(k, len(eol) - 1)] if len(eol) else n, n)
    # sequence bytes: the reference strips every line (tests/fasta.py:61-86: `line.strip()`), i.e. white space at the
    # ENDS of a line goes, white space inside a line stays (and later encodes as an unknown base)
    ws = (a == 13) | (a == 32) | (a == 9) | (a == 11) | (a == 12)
    keep = ~(nl | ws)
    if ws.any():
        c = np.cumsum(keep)                                  # non-blank bytes so far
        line_id = np.cumsum(nl) - nl                         # a newline belongs to the line it ends
        n_lines = int(line_id[-1]) + 1
        starts = np.concatenate(([0], np.flatnonzero(nl) + 1))[:n_lines]
        ends = np.concatenate((np.flatnonzero(nl), [n - 1]))[:n_lines]
        c_before = np.where(starts > 0, c[np.maximum(starts - 1, 0)], 0)
        c_end = c[ends]
        inner = ws & (c > c_before[line_id]) & (c < c_end[line_id])
        keep |= inner
    for b, e in zip(hdr_pos, hdr_end):
        keep[b:e] = False
    if len(hdr_pos) == 0:
        if keep.any():
            raise ValueError("not in FASTA format")
        return Batch(np.zeros(0, np.uint8), np.zeros(1, np.int64), [], [])
    if keep[:hdr_pos[0]].any():
        keep[:hdr_pos[0]] = False                            # text before the first header is ignored, like the reference
    csum = np.concatenate(([0], np.cumsum(keep, dtype=np.int64)))
    bounds = np.concatenate((hdr_pos, [n]))
    offsets = csum[bounds] - csum[hdr_pos[0]]
    flat = _sequence_buffer(int(csum[-1]), pinned)
    np.compress(keep, a, out=flat)
    ids, descs = [], []
    raw = a.tobytes()
    for b, e in zip(hdr_pos, hdr_end):
        line = raw[b + 1:e].decode("utf-8", "replace")
        fields = line.split(maxsplit=1)
        ids.append(fields[0] if fields else "")
        descs.append(fields[1].rstrip("\r\n") if len(fields) > 1 else "")
    return Batch(flat, offsets.astype(np.int64), ids, descs)


def parse(source):
    """Record iterator with the interface of the reference's `parse` (tests/fasta.py:61-86)."""
    return read_batch(source).records()
