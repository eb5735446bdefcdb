"""CPU check of the TRAINING kernels' logic without a GPU: tests/emu/train_emu.cu runs the product's training
schedule (pyrodigal_b200/csrc/train_host.hpp: run_training) over a backend that executes the per-item functions
the CUDA kernels are made of (train_device.cuh) in host loops, and the result must be byte-identical to the
oracle's training struct.  The steps that are other, separately GPU-tested kernels of the product (overlapping
starts, training DP, coding score, SD bins) are answered by the oracle through callbacks.  This is test
infrastructure: nothing here is reachable from the product package."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import refutil as R
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "train_emu.cu")
LIB = os.path.join(HERE, "emu", "libtrain_emu.so")
CSRC = os.path.join(R.ROOT, "pyrodigal_b200", "csrc")


def _emu():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("train_device.cuh", "train_host.hpp", "common.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(p) for p in deps):
        subprocess.check_call(["nvcc", "-x", "cu", "-O2", "-std=c++17", "-fmad=false", "-Wno-deprecated-gpu-targets",
                               "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", LIB, SRC])
    return C.CDLL(LIB)


DP_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
SCORE_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)


def emu_train(seq, closed=False, mask=False, force_nonsd=False, tt=11, st_wt=4.35):
    lib = _emu()
    d, gc, unk = orc.encode(seq)
    masks = orc.find_masks(d, 50) if mask else None
    opts = orc.make_opts(closed=closed, masks=masks)
    nodes = orc.extract(d, tt, opts)
    nn = len(nodes)
    cls = (nodes["type"] | ((nodes["strand"] != 1) << 2) | ((nodes["edge"] != 0) << 3) | ((nodes["ndx"] % 3) << 5)).astype(np.uint8)
    ndx, sv = np.ascontiguousarray(nodes["ndx"]), np.ascontiguousarray(nodes["stop_val"])
    blob0 = np.zeros(1, dtype=orc.TRAINING_DTYPE)

    def dp_cb(p_gs, p_bias, p_tb, p_ov, p_sp, p_ipath):
        gs = np.ctypeslib.as_array(C.cast(p_gs, C.POINTER(C.c_double)), shape=(nn, 3))
        bias = np.ctypeslib.as_array(C.cast(p_bias, C.POINTER(C.c_double)), shape=(3,))
        nodes["gc_score"] = gs
        t = blob0.copy()
        t["bias"][0] = bias
        t["st_wt"], t["trans_table"] = st_wt, tt
        tb = t.tobytes()
        orc.record_overlapping_starts(nodes, tb, flag=0, max_overlap=60)
        orc.score_connections(nodes, tb, final=False)
        np.ctypeslib.as_array(C.cast(p_tb, C.POINTER(C.c_int32)), shape=(nn,))[:] = nodes["traceb"]
        np.ctypeslib.as_array(C.cast(p_ov, C.POINTER(C.c_int8)), shape=(nn,))[:] = nodes["ov_mark"]
        np.ctypeslib.as_array(C.cast(p_sp, C.POINTER(C.c_int32)), shape=(nn, 3))[:] = nodes["star_ptr"]
        # lib.pyx:1239-1251 + 1311: largest index among the best terminal nodes, -1 without a traceback
        term = ((nodes["strand"] == 1) & (nodes["type"] == 3)) | ((nodes["strand"] != 1) & (nodes["type"] != 3))
        ip = -1
        if term.any():
            sc = np.where(term, nodes["score"], -np.inf)
            ip = int(nn - 1 - np.argmax(sc[::-1]))
            if sc[ip] <= -1.0 or nodes["traceb"][ip] == -1:
                ip = -1
        C.cast(p_ipath, C.POINTER(C.c_int32))[0] = ip

    def score_cb(p_t, p_cs, p_rbs):
        tb = C.string_at(p_t, orc.TRAINING_SIZE)
        orc.lib().orc_raw_coding_score(orc._p(d), len(d), orc._p(nodes), nn, orc.tinf_ptr(tb)[1])
        orc.lib().orc_rbs_score(orc._p(d), len(d), orc._p(nodes), nn, orc.tinf_ptr(tb)[1])
        np.ctypeslib.as_array(C.cast(p_cs, C.POINTER(C.c_double)), shape=(nn,))[:] = nodes["cscore"]
        np.ctypeslib.as_array(C.cast(p_rbs, C.POINTER(C.c_uint8)), shape=(nn, 2))[:] = nodes["rbs"]

    out = np.zeros(orc.TRAINING_SIZE, dtype=np.uint8)
    gp = np.zeros(len(d) + 1, dtype=np.int8)
    gs = np.zeros((nn + 1, 3), dtype=np.float64)
    niv = C.c_int32(0)
    cb1, cb2 = DP_CB(dp_cb), SCORE_CB(score_cb)
    lib.emu_train.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_double, C.c_int, DP_CB, SCORE_CB, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.emu_train(orc._p(d), len(d), gc, nn, orc._p(ndx), orc._p(sv), orc._p(cls), tt, st_wt, int(force_nonsd),
                       cb1, cb2, orc._p(out), orc._p(gp), orc._p(gs), C.byref(niv))
    assert rc == 0
    ref = orc.train(d, gc / len(d), translation_table=tt, start_weight=st_wt, force_nonsd=force_nonsd, opts=opts)
    return out.tobytes(), ref, gp[:len(d)], orc.gc_frame_plot(d)


def fields_differing(a, b):
    x, y = (np.frombuffer(v, dtype=orc.TRAINING_DTYPE)[0] for v in (a, b))
    return [f for f in orc.TRAINING_DTYPE.names if not np.array_equal(x[f], y[f], equal_nan=True)]


CASES = [
    ("sd_real_100k", None, dict(closed=True)),
    ("nonsd_60k", (60000, .45, 501, 0.0), {}),
    ("forced_tt4", (30000, .38, 502, 0.0), dict(force_nonsd=True, tt=4)),
    ("N_mask", (40000, .55, 503, 0.002), dict(mask=True, st_wt=3.9)),
    ("closed_high_gc", (25000, .68, 504, 0.0), dict(closed=True)),
    ("len_mod3_1", (20002, .5, 505, 0.0), {}),
    ("len_mod3_2", (20003, .31, 506, 0.001), {}),
]


@pytest.mark.parametrize("name,spec,kw", CASES, ids=[c[0] for c in CASES])
def test_training_kernels_logic_matches_oracle(name, spec, kw):
    if spec is None:
        seq = np.load(os.path.join(HERE, "golden", "train_cases.npz"))["ref100k_closed/seq"].tobytes()
    else:
        seq = R.synth(spec[0], spec[1], seed=spec[2], n_frac=spec[3])
    mine, ref, gp, gp_ref = emu_train(seq, **kw)
    assert np.array_equal(gp, gp_ref)
    assert fields_differing(mine, ref) == []
    assert mine == ref
