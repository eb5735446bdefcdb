"""TEST INFRASTRUCTURE ONLY: randomised comparison of the emulated CUDA library (tests/emu/libpgpu_emu.so) with the
oracle -- meta-mode find_genes on batches of odd contigs (many Ns, tiny and AT/GC-extreme contigs, low-complexity
repeats, lower case and IUPAC letters), both end modes, with and without masks.  Usage:
    python tests/emu/fuzz.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE), HERE]
import refutil as R  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import emu_capi  # noqa: E402


def odd_contig(rng, k):
    kind = int(rng.integers(0, 9))
    n = int(rng.choice([0, 1, 2, 3, 5, 60, 89, 90, 91, 120, 300, 900, 2999, 3000, 3001, 6000, 15000, 40000]))
    gc = float(rng.choice([0.05, 0.2, 0.3, 0.5, 0.7, 0.8, 0.95]))
    s = bytearray(R.synth(n, gc, int(rng.integers(1 << 30)), n_frac=float(rng.choice([0, 0, 0.001, 0.02, 0.3]))))
    if kind == 1 and n > 200:      # long run of Ns
        a = int(rng.integers(0, n - 100)); s[a:a + int(rng.integers(10, 100))] = b"N" * len(s[a:a + int(rng.integers(10, 100))])
    elif kind == 2 and n > 0:      # lower case / IUPAC letters
        idx = rng.integers(0, n, size=max(1, n // 50))
        for i in idx:
            s[i] = int(rng.choice(list(b"acgtRYKMSWnBDHV")))
    elif kind == 3 and n > 50:     # low-complexity repeat
        unit = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 7))).astype(np.uint8))
        s = bytearray((unit * (n // len(unit) + 1))[:n])
    elif kind == 4 and n > 300:    # stop-free stretch (giant ORF) inside
        a = int(rng.integers(0, n // 2)); ln = min(n - a, int(rng.integers(300, 9000)))
        s[a:a + ln] = (b"GCC" * (ln // 3 + 1))[:ln]
    elif kind == 5 and n > 100:    # dense starts / stops
        s = bytearray((b"ATGTAA" * (n // 6 + 1))[:n])
    return bytes(s)


def main():
    orc.node_capacity = lambda slen: slen + 1024   # dense start/stop repeats exceed the wrapper's default estimate
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    capi = emu_capi.load()
    ctx = capi.Context(0)
    ctx.set_models(R.bins_blob(), 50)
    t0, rounds, contigs, genes_total = time.time(), 0, 0, 0
    while time.time() - t0 < seconds:
        closed, mask = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        seqs = [odd_contig(rng, k) for k in range(int(rng.integers(1, 12)))]
        arrs = [np.frombuffer(s, np.uint8) for s in seqs]
        off = np.zeros(len(arrs) + 1, np.int64)
        np.cumsum([len(a) for a in arrs], out=off[1:])
        flat = np.ascontiguousarray(np.concatenate(arrs)) if off[-1] else np.zeros(0, np.uint8)
        res = ctx.find_genes_batch(flat, off, capi.make_opts(meta=True, closed=closed, mask=mask, want_nodes=True))
        lean = ctx.find_genes_batch(flat, off, capi.make_opts(meta=True, closed=closed, mask=mask, want_nodes=False))
        assert lean.genes.tobytes() == res.genes.tobytes() and lean.gene_nodes.tobytes() == res.gene_nodes.tobytes()
        for k, s in enumerate(seqs):
            d, gc, unk = orc.encode(s)
            masks = orc.find_masks(d, 50) if mask else None
            genes, nodes, winner, pairs = orc.find_genes_meta(d, gc / len(d) if len(d) else 0.0, R.bins_blob(),
                                                              orc.make_opts(closed=closed, masks=masks))
            a, b = res.gene_off[k], res.gene_off[k + 1]
            ok = int(res.summary["winner"][k]) == winner and b - a == len(genes)
            ok = ok and all(np.array_equal(res.genes[f][a:b], genes[f]) for f in ("begin", "end", "start_ndx", "stop_ndx"))
            if ok and winner >= 0:
                n = res.nodes(k)
                ok = len(n) == len(nodes) and all(np.array_equal(n[f], nodes[f]) for f in (
                    "ndx", "stop_val", "strand", "type", "edge", "cscore", "sscore", "rscore", "uscore", "tscore", "gc_cont", "rbs"))
            if not ok:
                path = os.path.join(HERE, f"fuzz_fail_{seed}_{rounds}_{k}.fna")
                open(path, "wb").write(b">fail closed=%d mask=%d\n" % (closed, mask) + s + b"\n")
                print("MISMATCH", path, "closed", closed, "mask", mask, "len", len(s), "winner", int(res.summary["winner"][k]), winner)
                return 1
            genes_total += b - a
        rounds += 1
        contigs += len(seqs)
    print(f"ok: {rounds} batches, {contigs} contigs, {genes_total} genes in {time.time() - t0:.0f} s (seed {seed})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
