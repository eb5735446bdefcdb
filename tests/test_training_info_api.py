"""TrainingInfo mirror (lib.pyx:3898-4283 / lib.pyi:297-363): constructor keywords, to_dict / JSON round trip, property
setters, pickling (both this package's raw state and the reference's dict state), dump / load -- checked against
the reference's own TrainingInfo objects when its build is present (oracle/_ref)."""
import io
import json
import pickle

import numpy as np
import pytest

import refutil as R

L = pytest.importorskip("pyrodigal_b200.lib")


def builtin(k):
    return L.TrainingInfo._from_bytes(R.bin_blob(k))


@pytest.mark.parametrize("k", [0, 2, 11, 24, 49])
def test_to_dict_json_round_trip(k):
    t = builtin(k)
    d = t.to_dict()
    assert set(d) == {"gc", "translation_table", "start_weight", "bias", "type_weights", "uses_sd", "rbs_weights",
                      "upstream_compositions", "motif_weights", "missing_motif_weight", "coding_statistics"}
    again = L.TrainingInfo(**json.loads(json.dumps(d)))
    assert bytes(again) == bytes(t) and again == t and hash(again) == hash(t)
    assert pickle.loads(pickle.dumps(t)) == t
    buf = io.BytesIO()
    t.dump(buf)
    assert L.TrainingInfo.load(io.BytesIO(buf.getvalue())) == t
    assert t.__sizeof__() > L.TRAINING_SIZE


def test_setters_and_validation():
    src = builtin(11)
    t = L.TrainingInfo(0.5)
    t.gc, t.translation_table, t.start_weight = src.gc, src.translation_table, src.start_weight
    t.uses_sd, t.missing_motif_weight = src.uses_sd, src.missing_motif_weight
    t.bias, t.type_weights, t.rbs_weights = list(src.bias), tuple(src.type_weights), np.array(src.rbs_weights)
    t.upstream_compositions, t.motif_weights = src.upstream_compositions.tolist(), src.motif_weights
    t.coding_statistics = src.coding_statistics
    assert bytes(t) == bytes(src)
    with pytest.raises(ValueError):
        t.translation_table = 7
    with pytest.raises(ValueError):
        t.bias = [1.0, 2.0]
    with pytest.raises(ValueError):
        L.TrainingInfo(0.5, translation_table=99)
    with pytest.raises(EOFError):
        L.TrainingInfo.load(io.BytesIO(b"\0" * 100))


@pytest.mark.skipif(not R.have_reference(), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("k", [0, 7, 24, 33])
def test_against_reference_objects(k):
    ref = R.reference()
    rt = ref.METAGENOMIC_BINS[k].training_info
    mine = L.TrainingInfo._from_bytes(bytes(memoryview(rt)))
    assert mine.to_dict() == rt.to_dict()
    assert bytes(L.TrainingInfo(**rt.to_dict())) == bytes(memoryview(rt))
    t = L.TrainingInfo.__new__(L.TrainingInfo)
    t.__setstate__(rt.__getstate__())          # the reference pickles the dict form
    assert bytes(t) == bytes(memoryview(rt))
    buf = io.BytesIO()
    rt.dump(buf)
    assert bytes(L.TrainingInfo.load(io.BytesIO(buf.getvalue()))) == bytes(memoryview(rt))
