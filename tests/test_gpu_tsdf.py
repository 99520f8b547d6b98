"""-m gpu: TSDF voxel-block integration on the GPU vs the brute-force CPU oracle.
Tolerance from north_star: |sdf|,|weight| within 1e-4, identical chunk-key sets (the CUDA path
follows the oracle's operation order, so in practice the arrays are bit-identical)."""
import numpy as np
import pytest

from plvs_b200 import synth, tsdf as T
from oracle import tsdf as OT

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _pair(w, h, **kw):
    K = synth.intrinsics(w, h)
    p = T.default_params(**kw)
    g = T.ChiselServer(p)
    g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    o = OT.Map(p, threads=8)
    o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    return g, o


def _check(g, o, color=False):
    gk, gs, gw, gc = g.download()
    ok, os_, ow, oc = o.download()
    assert gk.shape == ok.shape and np.array_equal(gk, ok), f"chunk key sets differ: {len(gk)} vs {len(ok)}"
    assert np.abs(gw - ow).max() <= TOL
    known = ow > 0
    assert np.abs(gs[known] - os_[known]).max() <= TOL
    assert np.array_equal(gs[~known], os_[~known])          # untouched voxels keep the 99999 sentinel
    if color:
        assert np.array_equal(gc, oc)
    so, sg = o.stats(), g.stats()
    for f in ("n_blocks", "n_range", "n_updated", "n_new"):
        assert so[f] == sg[f], (f, so, sg)
    return len(gk)


@pytest.mark.parametrize("carve", [1, 0])
def test_scan_sequence(gpu, carve):
    g, o = _pair(160, 120, voxel_resolution=0.04, use_carving=carve, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=0)
    for f in (0, 1, 2, 9):
        d = synth.depth_frame(f, 160, 120)
        g.integrate(d, synth.pose(f)); o.integrate(d, synth.pose(f))
        n = _check(g, o)
    assert n > 50


def test_scan_color_sequence(gpu):
    g, o = _pair(160, 120, voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=1)
    for f in (0, 1, 2, 3, 4, 5, 6):
        d, c = synth.depth_frame(f, 160, 120), synth.bgr_frame(f, 160, 120)
        g.integrate(d, synth.pose(f), c); o.integrate(d, synth.pose(f), c)
        _check(g, o, color=True)


def test_carving_moves_surface(gpu):
    """a wall that jumps back by 1 m: voxels in front of the new surface get carved / reset"""
    for color in (False, True):
        g, o = _pair(128, 96, voxel_resolution=0.05, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=int(color))
        Twc = np.eye(4, dtype=np.float32)[:3]
        for depth_m in (1.5, 1.5, 2.5, 2.5):
            d = np.full((96, 128), depth_m, np.float32)
            c = np.full((96, 128, 3), 90, np.uint8) if color else None
            g.integrate(d, Twc, c); o.integrate(d, Twc, c)
            _check(g, o, color)


def test_nan_zero_and_rotated_pose(gpu):
    g, o = _pair(160, 120, voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=0)
    rng = np.random.default_rng(0)
    d = synth.depth_frame(4, 160, 120)
    d[rng.random(d.shape) < 0.05] = np.nan
    d[:10] = 0.0
    a = 0.6
    R = np.array([[np.cos(a), 0, np.sin(a)], [0.1, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    R, _ = np.linalg.qr(R)
    Twc = np.concatenate([R, [[0.2], [-0.1], [0.3]]], 1).astype(np.float32)
    g.integrate(d, Twc); o.integrate(d, Twc)
    _check(g, o)


def test_vga_1cm_single_scan(gpu):
    """BASELINE config 2 geometry (640x480, 1 cm voxels): one scan, depth clipped to 2 m to bound the oracle's brute force"""
    g, o = _pair(640, 480, voxel_resolution=0.01, use_carving=1, near_plane=0.1, far_plane=2.2, max_blocks=16384, use_color=1)
    d = synth.depth_frame(0)
    d[d > 2.0] = 0.0
    c = synth.bgr_frame(0)
    g.integrate(d, synth.pose(0), c); o.integrate(d, synth.pose(0), c)
    n = _check(g, o, color=True)
    assert n > 100
    s = g.stats()
    assert s["n_candidates"] < 0.5 * s["n_range"]          # the screen-space cull actually prunes
    assert s["n_candidates"] < 8 * s["n_updated"], s       # ... and is reasonably tight even for a small isolated object


def test_reset_and_errors(gpu):
    g, o = _pair(160, 120, voxel_resolution=0.04, max_blocks=4096, use_color=0)
    g.integrate(synth.depth_frame(0, 160, 120), synth.pose(0))
    assert g.stats()["n_blocks"] > 0
    g.Reset()
    assert len(g.download()[0]) == 0
    with pytest.raises(Exception):
        T.ChiselServer(T.default_params()).integrate(synth.depth_frame(0, 160, 120), synth.pose(0))     # no camera info
    small = T.ChiselServer(T.default_params(voxel_resolution=0.04, max_blocks=16, use_color=0))
    K = synth.intrinsics(160, 120)
    small.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], 160, 120)
    with pytest.raises(Exception):
        small.integrate(synth.depth_frame(0, 160, 120), synth.pose(0))                                  # pool exhausted is reported, not silent:
        small.stats()                                                                                   # at the latest by the next call that needs the map


def _check_cloud(g, o, color):
    gk, gs, gw, gc = g.download()
    ok, os_, ow, oc = o.download()
    assert gk.shape == ok.shape and np.array_equal(gk, ok), f"chunk key sets differ: {len(gk)} vs {len(ok)}"
    assert np.abs(gw - ow).max() <= TOL
    known = ow > 0
    assert np.abs(gs[known] - os_[known]).max() <= TOL
    assert np.array_equal(gs[~known], os_[~known])
    if color:
        assert np.array_equal(gc, oc)
    assert g.stats()["n_blocks"] == o.stats()["n_blocks"]
    return len(gk)


@pytest.mark.parametrize("color,carve", [(True, 1), (False, 1), (True, 0)])
def test_cloud_sequence(gpu, color, carve):
    """a26: Chisel::IntegratePointCloudWidthDepth (PLVS's default Chisel route) -- carve pass over the existing chunks,
    then per-point truncation-band ray casting with chunk creation; several points hit the same voxel, so the
    per-voxel update ORDER (point index) is part of the contract."""
    from plvs_b200 import scenario
    w, h = 320, 240
    K = synth.intrinsics(w, h)
    g, o = _pair(w, h, voxel_resolution=0.02, use_carving=carve, carving_dist=0.05, near_plane=0.1, far_plane=5.0, max_blocks=16384, use_color=int(color))
    for f in (0, 1, 2, 5):
        d = synth.depth_frame(f, w, h)
        c = synth.bgr_frame(f, w, h) if color else None
        xyz, rgb = scenario.cloud_from_depth(d, c, K, step=2)
        assert len(xyz) > 10000
        g.integrate_cloud(xyz, rgb, synth.pose(f), d); o.integrate_cloud(xyz, rgb, synth.pose(f), d)
        n = _check_cloud(g, o, color)
    assert n > 200


def test_cloud_edge_cases(gpu):
    from plvs_b200 import scenario
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    g, o = _pair(w, h, voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=1)
    Twc = synth.pose(3)
    # empty cloud with a depth image: only the carve pass runs (on an empty map: nothing)
    e = np.zeros((0, 3), np.float32)
    g.integrate_cloud(e, e, Twc, synth.depth_frame(3, w, h)); o.integrate_cloud(e, e, Twc, synth.depth_frame(3, w, h))
    assert g.stats()["n_blocks"] == 0
    # cloud without a depth image and without colours
    d = synth.depth_frame(3, w, h)
    xyz, _ = scenario.cloud_from_depth(d, None, K, step=1)
    g.integrate_cloud(xyz, None, Twc); o.integrate_cloud(xyz, None, Twc)
    _check_cloud(g, o, True)
    # mixed with the projective depth-scan path on the same map
    g.integrate(d, Twc, synth.bgr_frame(3, w, h)); o.integrate(d, Twc, synth.bgr_frame(3, w, h))
    _check_cloud(g, o, True)
    xyz, rgb = scenario.cloud_from_depth(synth.depth_frame(4, w, h), synth.bgr_frame(4, w, h), K, step=3)
    g.integrate_cloud(xyz, rgb, synth.pose(4), synth.depth_frame(4, w, h)); o.integrate_cloud(xyz, rgb, synth.pose(4), synth.depth_frame(4, w, h))
    _check_cloud(g, o, True)


def test_1080p_5mm_scan_pair(gpu):
    """BASELINE config 3 geometry (1920x1080, 5 mm voxels): two consecutive scans (the second one exercises carving, the
    carvable-octant mask and the bulk-copy ring over existing chunks); depth clipped to 3.2 m to bound the oracle."""
    w, h = 1920, 1080
    g, o = _pair(w, h, voxel_resolution=0.005, use_carving=1, near_plane=0.1, far_plane=3.4, max_blocks=40000, use_color=1)
    o2 = OT.Map(T.default_params(voxel_resolution=0.005, use_carving=1, near_plane=0.1, far_plane=3.4, max_blocks=40000, use_color=1), threads=32)
    K = synth.intrinsics(w, h)
    o2.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    for f in (0, 1):
        d = synth.depth_frame(f, w, h)
        d[d > 3.2] = 0.0
        c = synth.bgr_frame(f, w, h)
        g.integrate(d, synth.pose(f), c); o2.integrate(d, synth.pose(f), c)
        n = _check(g, o2, color=True)
    assert n > 1000


def test_many_scans_carvable_mask_stays_exact(gpu):
    """20 scans of a moving camera + a surface that retreats: the per-octant carvable mask must never hide a voxel the
    reference would carve (sdf / weight / colour compared after every scan)."""
    g, o = _pair(160, 120, voxel_resolution=0.03, use_carving=1, carving_dist=0.02, near_plane=0.1, far_plane=4.5, max_blocks=8192, use_color=1)
    for f in range(20):
        d = synth.depth_frame(f, 160, 120)
        if f >= 10:
            d = d + np.float32(0.12 * (f - 9))           # everything moves away from the camera: old surface voxels get carved
        c = synth.bgr_frame(f, 160, 120)
        g.integrate(d, synth.pose(f), c); o.integrate(d, synth.pose(f), c)
        _check(g, o, color=True)
    # the no-colour variant carves with Carve() (weight += 1.5) instead of Reset()
    g, o = _pair(160, 120, voxel_resolution=0.03, use_carving=1, carving_dist=0.02, near_plane=0.1, far_plane=4.5, max_blocks=8192, use_color=0)
    for f in range(12):
        d = synth.depth_frame(f, 160, 120) + np.float32(0.1 * max(0, f - 5))
        g.integrate(d, synth.pose(f)); o.integrate(d, synth.pose(f))
        _check(g, o)
