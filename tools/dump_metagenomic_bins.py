"""Dump the 50 pre-trained metagenomic models of the reference into pyrodigal_b200/data/.

The models are DATA (parameter tables, `struct _training`, 558392 bytes each; in the reference
they are C initialisers in vendor/Prodigal/training.c:55-1354 surfaced as
`pyrodigal.METAGENOMIC_BINS`, lib.pyx:4891-5069).  This script reads them through the
reference's own buffer protocol (`bytes(memoryview(TrainingInfo))`, lib.pyx:4047-4063) from the
build in oracle/_ref and stores them xz-compressed, plus their descriptions.

Run once in the build container:  python tools/dump_metagenomic_bins.py
"""
import json
import lzma
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import pyrodigal  # noqa: E402  (the unmodified reference)

out = os.path.join(ROOT, "pyrodigal_b200", "data")
os.makedirs(out, exist_ok=True)
blobs, descs = [], []
for b in pyrodigal.METAGENOMIC_BINS:
    raw = bytes(memoryview(b.training_info))
    assert len(raw) == 558392
    blobs.append(raw)
    descs.append(b.description)
with lzma.open(os.path.join(out, "metagenomic_bins.bin.xz"), "wb", preset=9) as f:
    f.write(b"".join(blobs))
with open(os.path.join(out, "metagenomic_bins.json"), "w") as f:
    json.dump({"stride": 558392, "count": len(blobs), "descriptions": descs,
               "source": "pyrodigal %s METAGENOMIC_BINS (Prodigal v2.6.3 training.c)" % pyrodigal.__version__}, f, indent=1)
print("wrote", len(blobs), "models")
