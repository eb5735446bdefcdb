"""The 128-byte `struct _node` layout pgpu_result_nodes_struct writes (pyrodigal_b200/_capi.py: NODE_STRUCT_DTYPE) is pinned
against the reference's own header, compiled where it lies (build container only: /root/reference does not travel)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include <string.h>
#include <stdint.h>
#include "node.h"
#define O(f) printf(#f " %zu\n", offsetof(struct _node, f))
int main(void) {
    struct _node n;
    printf("size %zu\n", sizeof(n));
    O(mot); O(gc_score); O(cscore); O(uscore); O(tscore); O(rscore); O(sscore); O(score); O(gc_cont); O(star_ptr);
    O(traceb); O(tracef); O(ndx); O(stop_val); O(ov_mark); O(strand); O(rbs); O(edge); O(elim); O(gc_bias); O(type);
    memset(&n, 0, sizeof n);
    n.mot.score = 1.5; n.mot.ndx = 0xABC; n.mot.spacer = 0x9; n.mot.len = 0x5; n.mot.spacendx = 0x2;
    uint32_t bits; memcpy(&bits, (char *)&n + 8, 4);
    double sc; memcpy(&sc, &n, 8);
    printf("bits %u\nscore_first %d\n", bits, sc == 1.5);
    return 0;
}
"""


def _dtype():
    # the dtype only, without loading the CUDA library (this test runs without a GPU)
    src = open(os.path.join(ROOT, "pyrodigal_b200", "_capi.py")).read()
    m = re.search(r"NODE_STRUCT_DTYPE = np\.dtype\(\n(.*?)\n\)\n", src, re.S)
    return eval("np.dtype(" + m.group(1) + ")", {"np": np})


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "Prodigal")), reason="reference sources not present")
def test_node_struct_dtype_matches_reference_header(tmp_path):
    c = tmp_path / "probe.c"
    c.write_text(PROBE)
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(REF, "src", "Prodigal"), "-I", os.path.join(REF, "vendor", "Prodigal"),
                           str(c), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    dt = _dtype()
    assert int(out["size"]) == dt.itemsize == 128
    for name in dt.names:
        ref_name = {"mot_score": "mot", "mot_bits": None}.get(name, name)
        if ref_name is None:
            continue
        assert dt.fields[name][1] == int(out[ref_name]), name
    assert dt.fields["mot_bits"][1] == 8 and out["score_first"] == "1"
    # the bit fields as api.cu packs them: ndx | spacer << 12 | len << 16 | spacendx << 19
    assert int(out["bits"]) == 0xABC | (0x9 << 12) | (0x5 << 16) | (0x2 << 19)


def test_node_struct_dtype_is_dense():
    dt = _dtype()
    covered = np.zeros(dt.itemsize, bool)
    for name in dt.names:
        sub, off = dt.fields[name][:2]
        assert not covered[off:off + sub.itemsize].any(), name
        covered[off:off + sub.itemsize] = True
    # the only padding: 4 bytes behind the motif bit fields
    assert list(np.nonzero(~covered)[0]) == [12, 13, 14, 15]
