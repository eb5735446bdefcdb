#!/usr/bin/env bash
# Builds the UNMODIFIED reference (pyrodigal, from /root/reference) into oracle/_ref/
# so it can (a) pin the C restatement in oracle/, (b) generate tests/golden fixtures,
# (c) serve as the `--impl reference` CPU arm of bench.py.  TEST INFRASTRUCTURE ONLY.
#
# The reference's hot path lives in a Cython module (src/pyrodigal/lib.pyx), so it
# cannot be compiled with a bare gcc line: Cython code generation + its CMake
# (scikit-build-core) build are required.  /root/reference is read-only, so the build
# runs from a scratch copy under /tmp; nothing from the reference is copied into the repo.
# `touch src/Prodigal/node.h`: in a checkout where every mtime is equal CMake's
# file(COPY) keeps the vendored 176-byte node.h instead of pyrodigal's packed
# 128-byte one.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "no reference at $REF; skipping" >&2; exit 0; fi
if [ -f "$OUT/pyrodigal/__init__.py" ]; then echo "oracle/_ref already built"; exit 0; fi
TMP="$(mktemp -d /tmp/pyrodigal_ref.XXXXXX)"
cp -r "$REF" "$TMP/src"
chmod -R u+w "$TMP/src"
# file(COPY) skips a file whose timestamp equals the destination's (whole seconds): make the two header
# generations differ by years, not by the few milliseconds between `cp` and `touch`
find "$TMP/src/vendor/Prodigal" -type f -exec touch -d '2001-01-01 00:00:00' {} +
touch -d '2020-01-01 00:00:00' "$TMP/src/src/Prodigal/node.h" "$TMP/src/src/Prodigal/CMakeLists.txt"
python -m pip wheel "$TMP/src" --no-index --no-build-isolation --no-deps -w "$TMP/whl" \
    --find-links /opt/wheelhouse
mkdir -p "$OUT"
python -m pip install --no-index --no-deps --target "$OUT" "$TMP"/whl/pyrodigal-*.whl
rm -rf "$TMP"
# the packed 128-byte node (float gc_cont) must be the one that was compiled
PYTHONPATH="$OUT" python - <<'PY'
import pyrodigal, struct
g = pyrodigal.GeneFinder(meta=True).find_genes("ATG" + "GCA" * 400 + "TAA" + "ACGT" * 300)
vals = [n.gc_cont for n in g.nodes if n.type != "Stop"]
assert vals and all(struct.unpack("f", struct.pack("f", v))[0] == v for v in vals), \
    "reference was built with the vendored node.h (double gc_cont): rebuild"
print("reference build uses the packed node layout (float gc_cont)")
PY
echo "built reference into $OUT"
