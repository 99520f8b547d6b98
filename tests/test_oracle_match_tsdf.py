"""CPU: known-answer tests that pin the matcher and TSDF oracles themselves (no GPU)."""
import numpy as np
import pytest

from oracle import match as OM, tsdf as OT
from plvs_b200 import synth, tsdf as T
from plvs_b200.matcher import Frame, MP_QUERY, LAST_QUERY, featvec
from plvs_b200.orb import KP_DTYPE


def _frame(xy, octaves, descs, uright=None):
    kp = np.zeros(len(xy), KP_DTYPE)
    kp["x"], kp["y"] = np.array(xy, np.float32).T
    kp["octave"] = octaves
    sf = np.array([1.2 ** i for i in range(8)], np.float32)
    return Frame(kp, np.array(descs, np.uint8), 640, 480, sf, uright=uright)


def _desc(nbits):
    d = np.zeros(256, np.uint8); d[:nbits] = 1
    return np.packbits(d)


def test_hamming_against_numpy():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8); b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    for i in range(500):
        assert OM.hamming256(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())
    assert OM.hamming256(_desc(0), _desc(256)) == 256 and OM.hamming256(_desc(7), _desc(7)) == 0


def test_projection_map_sequential_claims():
    # two map points project onto the same spot; keypoint A is nearer (dist 10) than B (dist 40) for both.
    F = _frame([(100, 100), (102, 100)], [0, 0], [_desc(10), _desc(40)])
    q = np.zeros(2, MP_QUERY)
    q["proj_x"], q["proj_y"], q["view_cos"], q["level"], q["flags"] = 100.5, 100, 0.5, 0, 1
    q["desc"] = _desc(0)
    n, a = OM.search_by_projection_map(F, q, 3.0, 0.8)
    # first point takes A (10 <= 0.8*40); the second skips A (claimed) and takes B, whose runner-up no longer exists
    assert n == 2 and list(a) == [0, 1]
    q["flags"] = 0      # claimants without observations do not block: the second point overwrites A
    n, a = OM.search_by_projection_map(F, q, 3.0, 0.8)
    assert n == 2 and list(a) == [1, -1]
    # ratio test: both keypoints equally good on the same level -> rejected
    F2 = _frame([(100, 100), (102, 100)], [0, 0], [_desc(10), _desc(11)])
    q["flags"] = 1
    assert OM.search_by_projection_map(F2, q[:1], 3.0, 0.8)[0] == 0
    # ... but accepted when the runner-up lives on another pyramid level
    F3 = _frame([(100, 100), (102, 100)], [0, 1], [_desc(10), _desc(11)])
    q["level"] = 1
    assert OM.search_by_projection_map(F3, q[:1], 3.0, 0.8)[0] == 1


def test_projection_last_level_quirk_and_orientation():
    F = _frame([(200, 200)], [2], [_desc(5)])
    q = np.zeros(1, LAST_QUERY)
    q["u"], q["v"], q["invz"], q["last_octave"], q["flags"] = 200, 200, 0.5, 2, 1
    q["desc"] = _desc(0)
    assert OM.search_by_projection_last(F, q, 15.0)[0] == 1
    # forward motion asks for (minLevel=octave, maxLevel=-1): the level filter then rejects everything unless octave==0
    assert OM.search_by_projection_last(F, q, 15.0, forward=True)[0] == 0
    F0 = _frame([(200, 200)], [0], [_desc(5)])
    q["last_octave"] = 0
    assert OM.search_by_projection_last(F0, q, 15.0, forward=True)[0] == 1
    # distance above TH_HIGH
    Ffar = _frame([(200, 200)], [0], [_desc(101)])
    assert OM.search_by_projection_last(Ffar, q, 15.0)[0] == 0


def test_triangulation_last_tie_wins_and_mappoint_gate():
    K1 = _frame([(100, 100)], [0], [_desc(0)])
    K2 = _frame([(300, 100), (320, 100), (340, 100)], [0, 0, 0], [_desc(20), _desc(20), _desc(60)])
    fv1, fv2 = featvec([5]), featvec([5, 5, 5])
    F12 = np.zeros((3, 3), np.float32); ep = np.array([-1000, -1000], np.float32)
    n, m = OM.search_for_triangulation(K1, K2, fv1, fv2, [0], [0, 0, 0], F12, ep, coarse=True, check_ori=False)
    assert n == 1 and m[0] == 1          # equal distance: the later candidate replaces the earlier one
    n, m = OM.search_for_triangulation(K1, K2, fv1, fv2, [0], [0, 1, 0], F12, ep, coarse=True, check_ori=False)
    assert m[0] == 0                      # candidates that already have a map point are skipped
    n, m = OM.search_for_triangulation(K1, K2, fv1, featvec([6, 6, 6]), [0], [0, 0, 0], F12, ep, coarse=True, check_ori=False)
    assert n == 0                         # different vocabulary node


def test_tsdf_running_mean_and_block_set():
    w, h = 64, 48
    p = T.default_params(voxel_resolution=0.05, use_carving=0, near_plane=0.1, far_plane=3.0, use_color=0, max_blocks=1024)
    m = OT.Map(p)
    m.set_camera(60.0, 60.0, 32.0, 24.0, w, h)
    Twc = np.eye(4, dtype=np.float32)[:3]
    d = np.full((h, w), 1.0, np.float32)
    m.integrate(d, Twc)
    keys, sdf, wt, _ = m.download()
    known = wt > 0
    assert known.any() and np.all(wt[known] == 1.0)
    # every updated voxel holds depth - z, and |sdf| stays below trunc(1 m) + 2*sqrt(3)*res
    band = 6 * (0.0019 - 0.00152 + 0.001504) + 2 * np.sqrt(3) * 0.05
    assert np.abs(sdf[known]).max() < band
    zc = (keys[:, 2:3] * 16 + (np.arange(4096) // 256)[None, :]) * np.float32(0.05) + np.float32(0.025)
    assert np.allclose(sdf[known], (1.0 - zc)[known], atol=1e-5)
    m.integrate(np.full((h, w), 1.02, np.float32), Twc)
    _, sdf2, wt2, _ = m.download()
    both = (wt2 == 2.0)
    assert both.any() and np.allclose(sdf2[both], (1.01 - zc)[both], atol=1e-5)      # mean of the two observations
    st = m.stats()
    assert st["n_collected"] > 0 and st["n_blocks"] == len(keys) + st["n_new"]


def test_depth_u16_scale_matches_cv2_for_every_value():
    """the u16 -> f32 depth scaling in front of the TSDF (src/Tracking.cc:1812-1813): all 65536 values x the factors of the shipped yamls"""
    cv2 = pytest.importorskip("cv2")
    from oracle import tsdf as OT
    v = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    for yaml_factor in (5000.0, 1000.0, 1.0, 5208.0, 1031.0):
        f = np.float32(1.0) / np.float32(yaml_factor)                 # Tracking: mDepthMapFactor = 1.0f / mDepthMapFactor
        want = cv2.multiply(v, float(f), dtype=cv2.CV_32F)
        assert np.array_equal(OT.depth_u16_to_f32(v, f).view(np.uint32), want.view(np.uint32))
