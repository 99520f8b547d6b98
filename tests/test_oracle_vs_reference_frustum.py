"""Frame::isInFrustum + MapPoint::PredictScale + Pinhole::project (the query builder of Tracking::SearchLocalPoints):
the oracle restatement against the reference's own text (sliced out of Frame.cc / MapPoint.cc / Pinhole.cpp at build time), and the
logarithm-free PredictScale the device uses against the direct formula for every float ratio in [1/64, 64]."""
import ctypes as C
import numpy as np
import pytest

from plvs_b200 import synth
from oracle import match as OM

pytestmark = pytest.mark.skipif(not OM.frustum_ref_available(), reason="oracle/_ref/libfrustum_ref.so not built (/root/reference absent)")


def cloud(seed, Twc, n=20000):
    """points in a box in front of (and partly behind / beside) the camera; normals = mean viewing direction, i.e. roughly along P - Ow"""
    rng = np.random.default_rng(seed)
    p = np.zeros(n, OM.MAP_POINT)
    T = np.asarray(Twc, np.float64).reshape(3, 4)
    pc = rng.uniform([-3, -2, -1.0], [3, 2, 6], (n, 3))
    p["xw"] = pc @ T[:, :3].T + T[:, 3]
    nrm = (p["xw"].astype(np.float64) - T[:, 3]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm += rng.normal(0, 0.6, (n, 3))
    p["normal"] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    p["min_dist"] = rng.uniform(0.2, 2.0, n); p["max_dist"] = p["min_dist"] * rng.uniform(2.0, 8.0, n)
    p["flags"] = rng.integers(0, 2, n); p["desc"] = rng.integers(0, 256, (n, 32))
    return p


@pytest.mark.parametrize("frame,sf,nl", [(0, 1.2, 8), (7, 1.2, 8), (23, 1.1, 12), (5, 1.5, 5)])
def test_in_frustum_equals_reference_text(frame, sf, nl):
    K = synth.intrinsics(640, 480)
    fr = OM.make_frustum(synth.pose(frame), K, (0.0, 0.0, 640.0, 480.0), K["bf"], 0.5, sf, nl)
    pts = cloud(frame, synth.pose(frame))
    n, q, iv = OM.in_frustum(fr, pts)
    rn, rq, riv = OM.ref_in_frustum(fr, pts)
    assert n == rn and np.array_equal(iv, riv)
    for f in ("proj_x", "proj_y"):                      # written before the later gates: compared for every point
        assert np.array_equal(q[f].view(np.uint32), rq[f].view(np.uint32)), f
    m = iv.astype(bool)
    for f in ("proj_xr", "track_depth", "view_cos", "level", "flags", "desc"):
        assert np.array_equal(q[f][m], rq[f][m]), f
    assert 500 < n < len(pts) and len(set(q["level"][m].tolist())) >= min(nl, 4)


def test_predict_scale_without_logarithm():
    l = OM._setup()
    l.orc_scale_threshold_mismatches.restype = C.c_longlong
    l.orc_scale_threshold_mismatches.argtypes = [C.c_float, C.c_int, C.c_float, C.c_float]
    for sf, nl in ((1.2, 8), (1.1, 12), (1.5, 5)):
        assert l.orc_scale_threshold_mismatches(sf, nl, 1.0 / 64, 64.0) == 0        # ~100 million floats each
        T = OM.scale_thresholds(sf, nl)
        assert np.all(np.diff(T) > 0) and abs(T[0] - 1.0) < 1e-6
