"""Matcher goldens recorded from the REFERENCE's own ORBmatcher.cc (tests/golden/make_match_golden.py):
 * not gpu: the oracle restatement reproduces them bit-exactly (runs without /root/reference);
 * gpu: the CUDA searches through the C ABI reproduce them, and agree with the compiled reference run live when
        oracle/_ref travelled to the box."""
import importlib.util
import pathlib
import numpy as np
import pytest

from plvs_b200.matcher import ORBmatcher
from oracle import match as OM

G = pathlib.Path(__file__).parent / "golden"
spec = importlib.util.spec_from_file_location("make_match_golden", G / "make_match_golden.py")
mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)


@pytest.fixture(scope="module")
def setup():
    K, fr = mk.frames()
    return mk.cases(K, fr), np.load(G / "match_ref.npz")


def test_oracle_reproduces_reference_golden(setup):
    cases, gold = setup
    for name, (kind, a) in cases.items():
        n, arr = mk.run_oracle(kind, a)
        assert n == int(gold[name + "_n"]) and np.array_equal(arr, gold[name]), name


def run_cuda(kind, a):
    if kind == "map":
        return ORBmatcher(a["ratio"], True).SearchByProjectionMap(a["F"], a["q"], a["th"], a["far"], a["th_far"], a["claimed"])
    if kind == "last":
        return ORBmatcher(0.9, a["check"]).SearchByProjectionLast(a["C"], a["q"], a["th"], a["fwd"], a["bwd"], claimed=a["claimed"])
    return ORBmatcher(0.6, a["check"]).SearchForTriangulation(a["K1"], a["K2"], a["fv1"], a["fv2"], a["has1"], a["has2"], a["F12"], a["ep"], a["only"], a["coarse"])


@pytest.mark.gpu
def test_cuda_reproduces_reference_golden(gpu, setup):
    cases, gold = setup
    live = OM.ref_available()
    for name, (kind, a) in cases.items():
        n, arr = run_cuda(kind, a)
        assert n == int(gold[name + "_n"]) and np.array_equal(arr, gold[name]), name
        if live:
            rn, rarr = mk.run_ref(kind, a)
            assert n == rn and np.array_equal(arr, rarr), name
