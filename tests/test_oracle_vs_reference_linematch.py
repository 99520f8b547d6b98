"""CPU: the restatement of the line-descriptor 2-NN (oracle/linematch.py: brute force with the multi-index-hashing discovery order as a per-pair key)
against the reference's own matcher -- Thirdparty/line_descriptor/src/binary_descriptor_matcher_custom.cpp compiled unmodified into
oracle/_ref/liblinematch_ref.so -- on neighbour indices (tie order included), distances, the ratio test and the compact-result row order; and both
against the committed golden (tests/golden/linematch_ref.npz, recorded from the compiled reference by tests/golden/make_linematch_golden.py)."""
import pathlib
import numpy as np
import pytest

from oracle import linematch as L
from tests.linematch_cases import cases

GOLD = pathlib.Path(__file__).parent / "golden" / "linematch_ref.npz"


def _ref():
    try:
        return L.RefLineMatcher()
    except Exception:
        return None


def test_enumeration_rank_is_a_ranking_per_popcount():
    r = L.enumeration_rank()
    from math import comb
    for s in range(9):
        ranks = sorted(int(r[m]) for m in range(256) if bin(m).count("1") == s)
        assert ranks == list(range(comb(8, s)))


@pytest.mark.skipif(_ref() is None, reason="oracle/_ref/liblinematch_ref.so not built (no /root/reference here)")
def test_restatement_equals_the_compiled_matcher():
    R = L.RefLineMatcher()
    n_ties = 0
    for q, t, mask in cases(60, seed=1):
        a = L.knn2(q, t, mask, 0.78); b = R.knn2(q, t, mask, 0.78)
        for x, y in zip(a[:4], b[:4]):
            assert np.array_equal(x, y)
        assert a[4] == b[4]
        n_ties += int(np.sum(a[2][:, 0] == a[2][:, 1]))
    assert n_ties > 100          # the cases do exercise the order of equidistant neighbours


def test_restatement_equals_the_golden():
    g = np.load(GOLD)
    for i, (q, t, mask) in enumerate(cases(int(g["n_cases"]), seed=int(g["seed"]))):
        a = L.knn2(q, t, mask, float(g["nn_ratio"]))
        assert np.array_equal(a[0], g[f"qi{i}"]) and np.array_equal(a[1], g[f"ti{i}"]) and np.array_equal(a[2], g[f"di{i}"]) and np.array_equal(a[3], g[f"vi{i}"])
