"""Pins the ORB oracle (oracle/orb_oracle.cpp + the cv2-driven arm, which the CUDA extractor is compared with) to the
REFERENCE's own src/ORBextractor.cc, compiled from /root/reference by oracle/ref_build.py into oracle/_ref/liborb_ref.so
against the OpenCV stand-in of oracle/cv_standin (cv::Mat bookkeeping + the four OpenCV primitives, each pinned
bit-exactly to the real cv2 by tests/test_oracle_orb.py).  With it, everything that is PLVS/ORB-SLAM code -- tables,
pyramid, per-cell FAST loop and threshold fallback, DistributeOctTree with the real std::list/std::sort, IC_Angle,
steered rBRIEF, lapping-area assembly -- is the reference's own.  Skipped when neither /root/reference nor a prebuilt
oracle/_ref is present."""
import pathlib
import numpy as np
import pytest

from plvs_b200 import synth
from oracle import orb as O

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/liborb_ref.so not built (/root/reference absent)")
GOLD = pathlib.Path(__file__).resolve().parent / "golden"
FIELDS = ("x", "y", "size", "angle", "response", "octave")


def same(a, b):
    assert len(a[0]) == len(b[0])
    for f in FIELDS:
        assert np.array_equal(a[0][f], b[0][f]), f
    assert np.array_equal(a[1], b[1]) and a[2] == b[2]


@pytest.mark.parametrize("w,h,nfeat,frame", [(640, 480, 1000, 0), (640, 480, 2000, 3), (752, 480, 1200, 2), (333, 257, 700, 4), (1920, 1080, 4000, 1)])
def test_reference_equals_both_oracle_arms(w, h, nfeat, frame):
    img = synth.gray_frame(frame, w, h)
    ref = O.RefExtractor(nfeat)
    r = ref(img)
    same(r, O.extract_port(img, nfeat)[:3])
    c = O.extract_cv2(img, nfeat, angle_impl="c", return_internals=True)
    same(r, c[:3])
    assert set(r[0]["class_id"].tolist()) == {-1}
    # public members read by callers (a10): the pyramid and the blurred pyramid
    for l in range(8):
        assert np.array_equal(ref.level(l), c[4]["pyramid"][l])
        if c[4]["blurred"][l] is not None:
            assert np.array_equal(ref.level(l, filtered=True), c[4]["blurred"][l])
    tab, rt = O.Tables(nfeat), ref.tables()
    for k in ("scale", "inv_scale", "sigma2", "inv_sigma2"):
        assert np.array_equal(getattr(tab, k), rt[k]), k


def test_lapping_area_and_low_texture():
    img = synth.gray_frame(6, 480, 360)
    for lap in ((0, 1000), (100, 300)):                 # mono passes {0,1000} (src/Frame.cc:630): everything goes to the back
        r = O.RefExtractor(700)(img, lap)
        same(r, O.extract_port(img, 700, lapping=lap)[:3])
        assert r[2] < len(r[0])
    flat = np.full((240, 320), 127, np.uint8); flat[100:140, 150:200] = 30            # few corners: minThFAST fallback cells, tiny octree
    same(O.RefExtractor(500)(flat), O.extract_port(flat, 500)[:3])
    rng = np.random.default_rng(9)
    noise = rng.integers(0, 256, (200, 260), dtype=np.uint8)                          # response ties everywhere: std::sort tie order matters
    same(O.RefExtractor(300)(noise), O.extract_port(noise, 300)[:3])


@pytest.mark.parametrize("name", ["orb_qvga_f0_500", "orb_vga_f3_1000"])
def test_reference_reproduces_committed_goldens(name):
    g = np.load(GOLD / f"{name}.npz")
    nfeat, nlev, ini, mn = (int(v) for v in g["params"])
    kp, desc, mono = O.RefExtractor(nfeat, float(g["scale_factor"]), nlev, ini, mn)(g["image"])
    assert mono == int(g["mono_index"])
    for f in FIELDS:
        assert np.array_equal(kp[f], g["keypoints"][f]), f
    assert np.array_equal(desc, g["descriptors"])


@pytest.mark.parametrize("nfeat,sf,nl,ini,mn", [(800, 1.3, 5, 12, 5), (1500, 1.1, 8, 30, 10), (300, 1.2, 3, 20, 7), (600, 1.5, 6, 25, 7)])
def test_other_extractor_settings(nfeat, sf, nl, ini, mn):
    """scale factors / level counts / FAST thresholds other than PLVS's defaults: tables, quotas and the per-level geometry follow the constructor"""
    img = synth.gray_frame(8, 512, 384)
    r = O.RefExtractor(nfeat, sf, nl, ini, mn)(img)
    same(r, O.extract_port(img, nfeat, sf, nl, ini, mn)[:3])
    assert len(r[0]) > 0.5 * nfeat
