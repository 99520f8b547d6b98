"""Pins the TSDF oracle (oracle/tsdf_oracle.cpp, the restatement the CUDA path is compared with) to the REFERENCE's
own open_chisel sources, compiled from /root/reference by oracle/ref_build.py into oracle/_ref/libchisel_ref.so
(against the Eigen stand-in in oracle/eigen_standin/ -- Eigen itself is not installed).  Bit-exact on chunk keys,
sdf, weight and colour for every mode the product implements.  Skipped when neither /root/reference nor a prebuilt
oracle/_ref is present."""
import contextlib
import os
import numpy as np
import pytest

from plvs_b200 import synth, scenario, tsdf as T
from oracle import tsdf as OT

pytestmark = pytest.mark.skipif(not OT.ref_available(), reason="oracle/_ref/libchisel_ref.so not built (/root/reference absent)")


@contextlib.contextmanager
def quiet():
    """the reference prints per scan ("CHISEL: Integrating a scan", chunk counts); keep the test log readable"""
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 1)
    try:
        yield
    finally:
        os.dup2(saved, 1); os.close(null); os.close(saved)


def rot_pose(f, ax, deg):
    P = synth.pose(f).copy().reshape(3, 4)
    a = np.deg2rad(deg); c, s = np.cos(a), np.sin(a)
    R = {0: [[1, 0, 0], [0, c, -s], [0, s, c]], 1: [[c, 0, s], [0, 1, 0], [-s, 0, c]], 2: [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[ax]
    P[:, :3] = (P[:, :3].astype(np.float64) @ np.array(R)).astype(np.float32)
    return P


def same(o, r, colour=True):
    ok, os_, ow, oc = o.download(); rk, rs, rw, rc = r.download()
    assert np.array_equal(ok, rk), (len(ok), len(rk))
    assert np.array_equal(ow, rw) and np.array_equal(os_, rs)
    if colour:
        assert np.array_equal(oc, rc)
    assert o.stats()["n_blocks"] == r.stats()["n_blocks"]
    return len(ok)


def pair(w, h, **kw):
    K = synth.intrinsics(w, h)
    p = T.default_params(max_blocks=8192, **kw)
    o = OT.Map(p, threads=8); r = OT.RefMap(p)
    for m in (o, r):
        m.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    return K, o, r


@pytest.mark.parametrize("colour", [0, 1])
@pytest.mark.parametrize("carving", [0, 1])
def test_depth_scan_sequence(colour, carving):
    w, h = 160, 120
    K, o, r = pair(w, h, voxel_resolution=0.04, near_plane=0.1, far_plane=4.0, use_color=colour, use_carving=carving)
    n = 0
    for f in (0, 1, 2, 7):
        d = synth.depth_frame(f, w, h)
        bgr = synth.bgr_frame(f, w, h) if colour else None
        o.integrate(d, synth.pose(f), bgr)
        with quiet():
            r.integrate(d, synth.pose(f), bgr)
        assert o.stats()["n_range"] == r.stats()["n_range"]
        n = same(o, r)
    assert n > 20


def test_depth_scan_nan_zero_and_rotated_poses():
    w, h = 128, 96
    K, o, r = pair(w, h, voxel_resolution=0.05, near_plane=0.1, far_plane=5.0, use_color=1, use_carving=1)
    rng = np.random.default_rng(5)
    for f, (ax, deg) in enumerate([(0, 17.0), (1, -33.0), (2, 48.0), (1, 170.0), (0, -91.0)]):
        d = synth.depth_frame(f, w, h).copy()
        d[rng.random(d.shape) < 0.05] = np.nan
        d[rng.random(d.shape) < 0.05] = 0.0
        if f == 3:
            d[:, : w // 2] -= 0.4           # something moved closer / further: carving has work to do
            d = np.maximum(d, 0.0)
        P = rot_pose(f, ax, deg)
        bgr = synth.bgr_frame(f, w, h)
        o.integrate(d, P, bgr)
        with quiet():
            r.integrate(d, P, bgr)
        assert o.stats()["n_range"] == r.stats()["n_range"]
        same(o, r)


def test_depth_scan_carve_no_colour():
    """DistVoxel::Carve() (no-colour path) instead of Reset()"""
    w, h = 128, 96
    K, o, r = pair(w, h, voxel_resolution=0.05, near_plane=0.1, far_plane=5.0, use_color=0, use_carving=1)
    for f, shift in enumerate([0.0, 0.5, -0.3, 0.8]):
        d = synth.depth_frame(0, w, h) + np.float32(shift)
        o.integrate(d, synth.pose(0))
        with quiet():
            r.integrate(d, synth.pose(0))
        same(o, r)


@pytest.mark.parametrize("carving", [0, 1])
def test_point_cloud_with_depth(carving):
    w, h = 160, 120
    K, o, r = pair(w, h, voxel_resolution=0.04, near_plane=0.1, far_plane=4.0, use_color=1, use_carving=carving)
    for f in (0, 1, 2, 5):
        d = synth.depth_frame(f, w, h); c = synth.bgr_frame(f, w, h)
        if f == 2:
            d = np.maximum(d - np.float32(0.35), 0).astype(np.float32)
        xyz, rgb = scenario.cloud_from_depth(d, c, K, step=2)
        P = rot_pose(f, f % 3, 11.0 * f)
        o.integrate_cloud(xyz, rgb, P, d)
        with quiet():
            r.integrate_cloud(xyz, rgb, P, d)
        n = same(o, r)
    assert n > 20


def test_mixed_routes_on_one_map():
    w, h = 128, 96
    K, o, r = pair(w, h, voxel_resolution=0.05, near_plane=0.1, far_plane=4.0, use_color=1, use_carving=1)
    for f in range(4):
        d = synth.depth_frame(f, w, h); c = synth.bgr_frame(f, w, h)
        with quiet():
            if f % 2:
                xyz, rgb = scenario.cloud_from_depth(d, c, K, step=1)
                o.integrate_cloud(xyz, rgb, synth.pose(f), d); r.integrate_cloud(xyz, rgb, synth.pose(f), d)
            else:
                o.integrate(d, synth.pose(f), c); r.integrate(d, synth.pose(f), c)
        same(o, r)


@pytest.mark.parametrize("kw", [
    dict(voxel_resolution=0.08, trunc_scale=4.0, weight=2.0, carving_dist=0.1),
    dict(voxel_resolution=0.03, trunc_scale=8.0, weight=1.0, carving_dist=0.0, trunc_const=0.003),
    dict(voxel_resolution=0.05, trunc_quad=0.0, trunc_linear=0.0, trunc_const=0.02, trunc_scale=1.0, weight=3.0),     # constant truncation
])
def test_parameter_variants(kw):
    """truncation polynomial, weight (ChiselServer casts it to uint16, ChiselServer.cpp:108), carving distance and resolution other than
    the bench's: scan with colour, then the point-cloud route on the same map"""
    w, h = 128, 96
    K, o, r = pair(w, h, near_plane=0.1, far_plane=3.5, use_color=1, use_carving=1, **kw)
    for f in (0, 2, 3):
        d = synth.depth_frame(f, w, h); c = synth.bgr_frame(f, w, h)
        if f == 3:
            d = np.maximum(d - np.float32(0.4), 0).astype(np.float32)
        o.integrate(d, synth.pose(f), c)
        with quiet():
            r.integrate(d, synth.pose(f), c)
        same(o, r)
    d = synth.depth_frame(5, w, h); c = synth.bgr_frame(5, w, h)
    xyz, rgb = scenario.cloud_from_depth(d, c, K, step=2)
    o.integrate_cloud(xyz, rgb, synth.pose(5), d)
    with quiet():
        r.integrate_cloud(xyz, rgb, synth.pose(5), d)
    same(o, r)
