"""Pins the bag-of-words oracle (oracle/bow_oracle.cpp: vocabulary text loader + DBoW2 transform) to the REFERENCE's own
DBoW2 (Thirdparty/DBoW2 compiled into oracle/_ref/libbow_ref.so): per-feature word / weight / node, BowVector (ids and normalised values, bit for
bit) and FeatureVector, for the four weighting and three of the scoring types, with stopped (zero-weight) words and distance ties."""
import numpy as np
import pytest

from plvs_b200 import synth
from oracle import bow as OB, orb as O

pytestmark = pytest.mark.skipif(not OB.ref_available(), reason="oracle/_ref/libbow_ref.so not built (/root/reference absent)")


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        if a[k].dtype == np.float64:
            assert np.array_equal(a[k].view(np.uint64), b[k].view(np.uint64)), k
        else:
            assert np.array_equal(a[k], b[k]), k


@pytest.fixture(scope="module")
def descriptors():
    return np.concatenate([O.extract_port(synth.gray_frame(f), 1500)[1] for f in (3, 4)])


@pytest.mark.parametrize("k,L,levelsup,scoring,weighting,zero", [(10, 3, 2, 0, 0, 0.0), (10, 3, 1, 0, 0, 0.3), (6, 4, 4, 1, 1, 0.1), (4, 5, 3, 5, 2, 0.0),
                                                                   (5, 3, 0, 2, 3, 0.2), (10, 2, 4, 0, 0, 0.0)])
def test_transform(tmp_path, descriptors, k, L, levelsup, scoring, weighting, zero):
    path = tmp_path / "voc.txt"
    OB.write_vocabulary(path, k, L, seed=k * 10 + L, scoring=scoring, weighting=weighting, zero_weight_fraction=zero)
    o, r = OB.Vocabulary(path), OB.RefVocabulary(path)
    assert o.size() == r.size() == k ** L
    a, b = o.transform(descriptors, levelsup), r.transform(descriptors, levelsup)
    _same(a, b)
    assert len(a["bow_ids"]) > 10 and (np.diff(a["bow_ids"].astype(np.int64)) > 0).all()
    if scoring == 0:
        assert abs(a["bow_vals"].sum() - 1.0) < 1e-9
    if zero:
        assert (a["weight"] == 0).any() and a["fv_offsets"][-1] == (a["weight"] > 0).sum()


def test_ties_pick_the_first_child_and_empty_input(tmp_path):
    path = tmp_path / "voc.txt"
    OB.write_vocabulary(path, 8, 2, seed=1, clustered=False)
    o, r = OB.Vocabulary(path), OB.RefVocabulary(path)
    rng = np.random.default_rng(0)
    d = rng.integers(0, 2, (3000, 32), dtype=np.uint8) * 255                 # few distinct byte values: many equal distances
    _same(o.transform(d, 1), r.transform(d, 1))
    e = np.zeros((0, 32), np.uint8)
    _same(o.transform(e, 1), r.transform(e, 1))
