"""-m gpu: Frame::isInFrustum on the device against the oracle, which tests/test_oracle_vs_reference_frustum.py pins to
the reference's own text (bit-exact mTrackProjX/Y/XR, mTrackDepth, mTrackViewCos, mnTrackScaleLevel, mbTrackInView)."""
import numpy as np
import pytest

from plvs_b200 import synth
from plvs_b200.matcher import ORBmatcher
from oracle import match as OM

pytestmark = pytest.mark.gpu


def _cloud(seed, Twc, n=20000):
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
def test_in_frustum_on_device(gpu, frame, sf, nl):
    K = synth.intrinsics(640, 480)
    fr = OM.make_frustum(synth.pose(frame), K, (0.0, 0.0, 640.0, 480.0), K["bf"], 0.5, sf, nl)
    pts = _cloud(frame, synth.pose(frame))
    n, q, iv = ORBmatcher(0.8, True).InFrustum(fr, pts)
    on, oq, oiv = OM.in_frustum(fr, pts)
    assert n == on and np.array_equal(iv, oiv)
    for f in ("proj_x", "proj_y", "flags", "desc"):
        assert np.array_equal(q[f], oq[f]), f
    m = iv.astype(bool)
    for f in ("proj_xr", "track_depth", "view_cos", "level"):
        assert np.array_equal(q[f][m].view(np.uint32) if q[f].dtype == np.float32 else q[f][m], oq[f][m].view(np.uint32) if oq[f].dtype == np.float32 else oq[f][m]), f
    n0, q0, iv0 = ORBmatcher(0.8, True).InFrustum(fr, pts[:0])
    assert n0 == 0 and len(q0) == 0
