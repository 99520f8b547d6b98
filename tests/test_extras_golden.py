"""Golden vectors recorded from the REFERENCE's own code (tests/golden/extras_ref.npz, made by tests/golden/make_extras_golden.py from oracle/_ref) for the
rows widened last: SearchForInitialization, mesh read-out, bag-of-words transform.  They hold where /root/reference is absent: the oracle is checked
against them here (CPU), the product in tests/test_gpu_widened.py's GPU run and on the CPU model."""
import hashlib
import pathlib

import numpy as np
import pytest

from oracle import match as OM, tsdf as OT, bow as OB
from tests.golden import extras_cases as X

GOLD = np.load(pathlib.Path(__file__).resolve().parent / "golden" / "extras_ref.npz")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


def check_initialization(search):
    """search(F1, F2, prev, window, ratio, check) -> (n, matches12, prev)"""
    f1, f2, f3 = X.init_frames()
    prev = np.stack([f1.keys["x"], f1.keys["y"]], 1)
    for window, ratio, check in X.INIT_PARAMS:
        tag = "init_%d_%g_%d" % (window, ratio, check)
        n, m, p = search(f1, f2, prev, window, ratio, check)
        n2, m2, p2 = search(f1, f3, p, window, ratio, check)
        assert [n, n2] == list(GOLD[tag + "_n"]) and np.array_equal(np.stack([m, m2]), GOLD[tag + "_m"])
        assert np.array_equal(np.stack([p, p2]).view(np.uint32), GOLD[tag + "_p"].view(np.uint32))


def check_meshes(make_map, meshes_of):
    for name, kw, frames, color in X.MESH_CASES:
        keys, counts, V, N, C = meshes_of(make_map(kw, frames, color))
        assert np.array_equal(keys, GOLD["mesh_%s_keys" % name]) and np.array_equal(counts, GOLD["mesh_%s_counts" % name]), name
        assert np.array_equal(digest(V, N, C), GOLD["mesh_%s_digest" % name]), name


def check_bow(make_vocabulary, tmp_path):
    desc = X.bow_descriptors()
    for k, L, levelsup, scoring, weighting, zero in X.BOW_CASES:
        path = tmp_path / ("voc_%d_%d.txt" % (k, L))
        OB.write_vocabulary(path, k, L, seed=k * 10 + L, scoring=scoring, weighting=weighting, zero_weight_fraction=zero)
        r = make_vocabulary(path).transform(desc, levelsup)
        tag = "bow_%d_%d_%d_%d_%d" % (k, L, levelsup, scoring, weighting)
        for name in ("word", "node", "bow_ids", "fv_nodes", "fv_offsets"):
            assert np.array_equal(r[name], GOLD[tag + "_" + name]), (tag, name)
        assert np.array_equal(r["bow_vals"].view(np.uint64), GOLD[tag + "_bow_vals"].view(np.uint64)), tag
        assert np.array_equal(digest(r["weight"], r["fv_features"]), GOLD[tag + "_digest"]), tag


def test_oracle_initialization_against_golden():
    check_initialization(OM.search_for_initialization)


def test_oracle_meshes_against_golden():
    check_meshes(lambda kw, frames, color: X.mesh_map(OT.Map, kw, frames, color, threads=8), lambda m: m.extract_mesh())


def test_oracle_bow_against_golden(tmp_path):
    check_bow(OB.Vocabulary, tmp_path)
