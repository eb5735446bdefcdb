"""TEST INFRASTRUCTURE ONLY: randomised comparison of the emulated CUDA library with the oracle for the single-mode
path -- GeneFinder.train (pgpu_train) must produce a byte-identical training struct, and find_genes with that model the
same genes and node scores.  Usage: python tests/emu/fuzz_train.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE), HERE]
import refutil as R  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import emu_capi  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    capi = emu_capi.load()
    t0, rounds = time.time(), 0
    while time.time() - t0 < seconds:
        n = int(rng.choice([20000, 20001, 23456, 30000, 45000, 70000]))
        gc = float(rng.choice([0.25, 0.35, 0.5, 0.65, 0.75]))
        seq = R.synth(n, gc, int(rng.integers(1 << 30)), n_frac=float(rng.choice([0, 0, 0.002])))
        closed, mask, force = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        tt = int(rng.choice([11, 11, 4, 1, 25]))
        st_wt = float(rng.choice([4.35, 2.5, 6.0]))
        d, gcc, unk = orc.encode(seq)
        masks = orc.find_masks(d, 50) if mask else None
        oo = orc.make_opts(closed=closed, masks=masks)
        want = orc.train(d, gcc / len(d), translation_table=tt, start_weight=st_wt, force_nonsd=force, opts=oo)
        ctx = capi.Context(0)
        a = np.frombuffer(seq, np.uint8)
        got, stats = ctx.train(a, capi.make_opts(closed=closed, mask=mask), translation_table=tt, start_weight=st_wt,
                               force_nonsd=force)
        if bytes(got) != want:
            print("TRAINING MISMATCH", n, gc, closed, mask, force, tt, st_wt)
            return 1
        ctx.set_models(want, 1)
        off = np.array([0, len(a)], np.int64)
        res = ctx.find_genes_batch(a, off, capi.make_opts(meta=False, single_model=0, closed=closed, mask=mask, want_nodes=True))
        genes, nodes, ipath = orc.find_genes_single(d, want, oo)
        ok = len(res.genes) == len(genes) and all(np.array_equal(res.genes[f], genes[f]) for f in ("begin", "end", "start_ndx", "stop_ndx"))
        nn = res.nodes(0)
        ok = ok and len(nn) == len(nodes) and all(np.array_equal(nn[f], nodes[f]) for f in (
            "ndx", "cscore", "sscore", "rscore", "uscore", "tscore", "score", "traceb", "tracef", "ov_mark", "elim", "star_ptr"))
        if not ok:
            print("FIND_GENES MISMATCH", n, gc, closed, mask, force, tt, st_wt)
            return 1
        ctx.close()
        rounds += 1
    print(f"ok: {rounds} train + find_genes rounds in {time.time() - t0:.0f} s (seed {seed})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
