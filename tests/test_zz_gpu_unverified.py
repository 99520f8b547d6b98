"""-m gpu: features written after the round's GPU budget was spent.  The kernels compile for sm_100a and their oracles are pinned on the CPU, but
these comparisons have not run on a GPU yet -- hence the non-strict xfail (a pass shows as XPASS, a failure does not fail the suite) and the
file name that sorts behind every other GPU test."""
import cv2
import numpy as np
import pytest

from plvs_b200 import synth
from plvs_b200.orb import ORBextractor
from oracle import orb as O

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(reason="first GPU run still pending", strict=False)]


def test_undistort_keypoints_on_device(gpu):
    """Frame::UndistortKeyPoints (§8f rank 2): device == oracle == the real cv2.undistortPoints, bit for bit, TUM1 / EuRoC / rational models"""
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    mono, kp, desc = ex(synth.gray_frame(3))
    xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float32)
    for K4, dist in (((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
                     ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
                     ((500.0, 500.0, 320.0, 240.0), (0.1, -0.2, 0.001, -0.002, 0.05, 0.01, -0.02, 0.003))):
        un, dptr = ex.UndistortKeyPoints(K4, np.array(dist, np.float32))
        want = O.undistort_points(xy, K4, np.array(dist, np.float32))
        Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
        cvw = cv2.undistortPoints(xy.reshape(-1, 1, 2), Km, np.array(dist, np.float32), None, Km).reshape(-1, 2)
        assert np.array_equal(want.view(np.uint32), cvw.view(np.uint32))
        assert np.array_equal(un["x"].view(np.uint32), want[:, 0].view(np.uint32)) and np.array_equal(un["y"].view(np.uint32), want[:, 1].view(np.uint32))
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(un[f], kp[f])
        assert dptr
    un, _ = ex.UndistortKeyPoints((517.3, 516.5, 318.6, 255.3), np.zeros(5, np.float32))
    assert np.array_equal(un, kp)
