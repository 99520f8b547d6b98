"""CPU: product host logic that needs no GPU -- the keypoint distributor against the list-based oracle, the glibc
sincosf port against libm, and the C-ABI library (loads, exports every declared symbol, fails loudly without a GPU)."""
import ctypes as C
import numpy as np
import pytest

from oracle import orb as O
from plvs_b200 import _lib, synth
from tests.native_build import build_host_checks


@pytest.fixture(scope="module")
def host():
    lib = C.CDLL(build_host_checks())
    lib.chk_sincosf_sweep.restype = C.c_long
    lib.chk_sincosf_sweep.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    return lib


def _distribute(host, xs, ys, rs, w, h, n_want):
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32); rs = np.ascontiguousarray(rs, np.int32)
    sel = np.empty(max(len(xs), 1), np.int32)
    m = host.chk_distribute(len(xs), xs.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), rs.ctypes.data_as(C.c_void_p),
                            16, w - 16, 16, h - 16, n_want, sel.ctypes.data_as(C.c_void_p))
    return sel[:m]


def test_distributor_matches_list_oracle_on_fast_candidates(host):
    for frame, (w, h), quotas in ((0, (640, 480), (434, 217, 60)), (3, (257, 193), (175, 87, 20)), (5, (179, 134), (122, 30, 400))):
        img = synth.gray_frame(frame, w, h)
        xs, ys, rs = O.fast_cells(img, 20, 7)
        for n_want in quotas:
            want = O.distribute_octree(xs, ys, rs, 16, w - 16, 16, h - 16, n_want)
            got = _distribute(host, xs, ys, rs, w, h, n_want)
            assert np.array_equal(got, want), (frame, w, h, n_want)


def test_distributor_random_sets_ties_and_degenerate(host):
    rng = np.random.default_rng(42)
    for trial in range(60):
        w, h = int(rng.integers(80, 700)), int(rng.integers(80, 500))
        n = int(rng.integers(0, 3000))
        xs = rng.integers(0, w - 32, n); ys = rng.integers(0, h - 32, n)
        rs = rng.integers(7, 12 if trial % 2 else 200, n)        # few distinct responses => many ties
        if trial % 7 == 0 and n:                                   # duplicates / clusters
            xs[: n // 2] = xs[0]; ys[: n // 3] = ys[0]
        n_want = int(rng.integers(1, 600))
        want = O.distribute_octree(xs, ys, rs, 16, w - 16, 16, h - 16, n_want)
        got = _distribute(host, xs, ys, rs, w, h, n_want)
        assert np.array_equal(got, want), trial


def test_distributor_tall_image_has_no_root(host):
    # nIni == 0 when the ROI is more than twice as tall as wide: the reference returns nothing (src/ORBextractor.cc:617-622)
    assert len(_distribute(host, [1, 2], [1, 2], [9, 9], 60, 200, 10)) == 0
    assert len(O.distribute_octree([1, 2], [1, 2], [9, 9], 16, 44, 16, 184, 10)) == 0


def test_stdsort_emulation_matches_libstdcxx_on_ties(host):
    """the distributor's tie order is whatever std::sort does with a key-only comparator: the emulation must permute alike"""
    rng = np.random.default_rng(7)
    for trial in range(400):
        n = int(rng.integers(0, 3000)) if trial % 5 else int(rng.integers(0, 40))
        kind = trial % 4
        if kind == 0:
            keys = rng.integers(0, 4, n)                       # almost everything ties
        elif kind == 1:
            keys = rng.integers(0, 1 << 20, n)
        elif kind == 2:
            keys = np.sort(rng.integers(0, 50, n))[::-1].copy()   # descending runs
        else:
            keys = (rng.integers(2, 30, n) << 16) | rng.integers(0, 600, n)   # (size, UL.x) like the real use
        keys = np.ascontiguousarray(keys, np.uint32)
        assert host.chk_stdsort(n, keys.ctypes.data_as(C.c_void_p), None) == 0, (trial, n)
        assert host.chk_stdsort_ranges(n, keys.ctypes.data_as(C.c_void_p)) == 0, ("range by range", trial, n)
    # adversarial for median-of-3 quicksort: organ-pipe and many-duplicates inputs big enough to hit the heapsort fallback
    for n in (5000, 20000):
        pipe = np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]).astype(np.uint32)
        assert host.chk_stdsort(len(pipe), pipe.ctypes.data_as(C.c_void_p), None) == 0
        assert host.chk_stdsort_ranges(len(pipe), pipe.ctypes.data_as(C.c_void_p)) == 0


def test_sincosf_port_equals_libm_exhaustively(host):
    """every float in [0, 6.4] (the steering angle range is [0, 2*pi]): 1.09e9 values, ~10 s"""
    first = C.c_float()
    hi = np.array([6.4], np.float32).view(np.uint32)[0]
    bad = host.chk_sincosf_sweep(0, int(hi), 1, C.byref(first))
    assert bad == 0, f"{bad} mismatches, first at {first.value!r}"


def test_sincosf_port_sampled_up_to_100(host):
    lo = np.array([6.4], np.float32).view(np.uint32)[0]; hi = np.array([100.0], np.float32).view(np.uint32)[0]
    assert host.chk_sincosf_sweep(int(lo), int(hi), 97, None) == 0


def test_abi_library_loads_and_exports_everything():
    lib = _lib.load()
    names = _lib.exported_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.plvs_version()


def test_no_cpu_fallback_without_gpu():
    lib = _lib.load()
    if lib.plvs_device_count() > 0:
        pytest.skip("a GPU is present")
    from plvs_b200.orb import ORBextractor
    from plvs_b200.matcher import ORBmatcher
    from plvs_b200 import tsdf
    for make in (lambda: ORBextractor(1000, 1.2, 8, 20, 7), lambda: ORBmatcher(0.8, True), lambda: tsdf.ChiselServer(tsdf.default_params())):
        with pytest.raises(_lib.PlvsError, match="no CUDA device"):
            make()


def test_hamming_host_helper():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (100, 32), dtype=np.uint8); b = rng.integers(0, 256, (100, 32), dtype=np.uint8)
    lib = _lib.load()
    for i in range(100):
        assert lib.plvs_hamming256(a[i].ctypes.data_as(C.c_void_p), b[i].ctypes.data_as(C.c_void_p)) == int(np.unpackbits(a[i] ^ b[i]).sum())
    z = np.zeros(32, np.uint8); o = np.full(32, 255, np.uint8)
    assert lib.plvs_hamming256(z.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p)) == 256


def test_bow_vector_host_step_matches_the_oracle(tmp_path):
    """plvs_bow_vector (the host half of plvs_voc_transform: BowVector accumulation and normalisation in the reference's order) against the oracle's
    BowVector, which is pinned to the compiled DBoW2 -- for every weighting and scoring type, with stopped words"""
    from oracle import bow as OB, orb as O
    from plvs_b200 import synth
    from plvs_b200.bow import bow_vector
    desc = O.extract_port(synth.gray_frame(3), 1500)[1]
    for scoring in range(6):
        for weighting in range(4):
            path = tmp_path / ("v%d%d.txt" % (scoring, weighting))
            OB.write_vocabulary(path, 6, 3, seed=scoring * 4 + weighting, scoring=scoring, weighting=weighting, zero_weight_fraction=0.2)
            o = OB.Vocabulary(path).transform(desc, 2)
            ids, vals = bow_vector(scoring, weighting, o["word"], o["weight"])
            assert np.array_equal(ids, o["bow_ids"]) and np.array_equal(vals.view(np.uint64), o["bow_vals"].view(np.uint64)), (scoring, weighting)
    ids, vals = bow_vector(0, 0, np.zeros(0, np.uint32), np.zeros(0, np.float64))
    assert len(ids) == 0
