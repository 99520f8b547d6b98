"""CPU: pins the ORB oracle (C restatement) against the real OpenCV in this image and against the committed
golden vectors, and checks the constructor tables of SURVEY.md §8."""
import ctypes as C
import pathlib
import numpy as np
import pytest

from oracle import orb as O
from plvs_b200 import synth

cv2 = pytest.importorskip("cv2")
GOLD = pathlib.Path(__file__).resolve().parent / "golden"


def test_tables_match_survey():
    t = O.Tables(1000, 1.2, 8)
    assert list(t.quota) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(O.Tables(2000, 1.2, 8).quota) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(O.Tables(4000, 1.2, 8).quota) == [869, 724, 603, 503, 419, 349, 291, 242]
    assert list(t.umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert [t.level_size(640, 480, l) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
    assert [t.level_size(1920, 1080, l) for l in range(8)][-1] == (536, 301)
    assert [t.level_size(752, 480, l) for l in range(8)][1] == (627, 400)


@pytest.mark.parametrize("size", [(640, 480), (752, 480), (333, 257)])
def test_resize_matches_cv2(size):
    w, h = size
    img = synth.gray_frame(1, 1024, 768)[:h, :w].copy()
    t = O.Tables(1000)
    cur = img
    for l in range(1, 8):
        lw, lh = t.level_size(w, h, l)
        a = cv2.resize(cur, (lw, lh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(a, O.resize_linear(cur, lw, lh))
        cur = a


def test_blur_matches_cv2():
    for img in (synth.gray_frame(2), synth.gray_frame(2, 179, 134), np.random.default_rng(0).integers(0, 256, (97, 133), dtype=np.uint8)):
        a = cv2.GaussianBlur(img.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(a, O.gauss7(img))


@pytest.mark.parametrize("th", [7, 20])
def test_fast_rect_matches_cv2_including_order(th):
    img = synth.gray_frame(4)
    f = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for (x0, y0, x1, y1) in [(0, 0, 640, 480), (16, 16, 57, 57), (300, 200, 347, 241), (600, 440, 624, 464), (10, 10, 17, 17)]:
        k = f.detect(img[y0:y1, x0:x1], None)
        xs, ys, rs = O.fast_rect(img, x0, y0, x1, y1, th)
        assert len(k) == len(xs)
        assert all(int(kk.pt[0]) == x and int(kk.pt[1]) == y and int(kk.response) == r for kk, x, y, r in zip(k, xs, ys, rs))


def test_fast_cells_match_cv2_calls():
    img = synth.gray_frame(5)
    for im in (img, O.resize_linear(img, 257, 193), (img.astype(np.float32) * 0.15 + 90).astype(np.uint8)):
        xs, ys, rs = O.fast_cells(im, 20, 7)
        cx, cy, cr = O.candidates_cv2(im, 20, 7)
        assert np.array_equal(xs, np.array(cx, np.int32)) and np.array_equal(ys, np.array(cy, np.int32)) and np.array_equal(rs, np.array(cr, np.int32))


def test_fast_atan2_golden():
    g = np.load(GOLD / "fast_atan2.npz")
    got = np.array([O.fast_atan2(float(y), float(x)) for y, x in g["yx"]], np.float32)
    assert np.array_equal(got.view(np.uint32), g["angle"].view(np.uint32))
    for (y, x), want in {(0, 0): 0.0, (0, 5): 0.0, (5, 0): 90.0, (0, -5): 180.0, (-5, 0): 270.0}.items():
        assert abs(O.fast_atan2(y, x) - want) < 0.02


def test_fast_atan2_matches_cv2_live():
    rng = np.random.default_rng(11)
    yx = rng.integers(-300000, 300000, size=(5000, 2)).astype(np.float32)
    for y, x in yx:
        assert np.float32(O.fast_atan2(float(y), float(x))) == np.float32(cv2.fastAtan2(float(y), float(x)))


@pytest.mark.parametrize("name", ["orb_qvga_f0_500", "orb_vga_f3_1000"])
def test_port_matches_golden(name):
    g = np.load(GOLD / f"{name}.npz")
    nfeat, nlev, ini, mn = (int(v) for v in g["params"])
    kp, desc, mono, ncand = O.extract_port(g["image"], nfeat, float(g["scale_factor"]), nlev, ini, mn)
    assert mono == int(g["mono_index"]) and ncand == int(g["n_candidates"])
    for f in kp.dtype.names:
        assert np.array_equal(kp[f], g["keypoints"][f]), f
    assert np.array_equal(desc, g["descriptors"])


def test_port_matches_cv2_arm_live_and_lapping():
    img = synth.gray_frame(6, 480, 360)
    for lap in ((0, 0), (100, 300)):
        a = O.extract_cv2(img, 700, lapping=lap)
        b = O.extract_port(img, 700, lapping=lap)
        assert a[2] == b[2] and a[3] == b[3]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        if lap != (0, 0):
            assert a[2] < len(a[0])


def test_descriptor_bit_order_and_steering():
    # angle 0: no steering, bit k of byte i is pattern test 8*i+k, LSB first (src/ORBextractor.cc:156-177)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (64, 64), dtype=np.uint8)
    d = O.descriptor(img, 32, 32, 0.0)
    import re
    pat = np.array([int(v) for v in re.findall(r"-?\d+", (pathlib.Path(__file__).resolve().parent.parent / "plvs_b200/csrc/orb_pattern.inc").read_text().split("\n", 2)[2])]).reshape(256, 4)
    bits = [(img[32 + y0, 32 + x0] < img[32 + y1, 32 + x1]) for x0, y0, x1, y1 in pat]
    want = np.packbits(np.array(bits, np.uint8), bitorder="little")
    assert np.array_equal(d, want)
    a, b = O.steer(90.0)
    assert abs(a) < 1e-6 and abs(b - 1) < 1e-6


def test_color_to_gray_matches_cv2_for_every_colour():
    """cv::cvtColor(COLOR_BGR2GRAY / RGB2GRAY / BGRA2GRAY / RGBA2GRAY) (src/Tracking.cc:1797-1810): all 2^24 colours"""
    vals = np.arange(256, dtype=np.uint8)
    B, G = np.meshgrid(vals, vals, indexing="ij")
    for r0 in range(256):
        cube = np.stack([B, G, np.full_like(B, r0)], -1)
        assert np.array_equal(cv2.cvtColor(cube, cv2.COLOR_BGR2GRAY), O.color_to_gray(cube))
        if r0 % 16 == 0:
            assert np.array_equal(cv2.cvtColor(cube, cv2.COLOR_RGB2GRAY), O.color_to_gray(cube, rgb=True))
    rng = np.random.default_rng(4)
    img4 = rng.integers(0, 256, (120, 160, 4), dtype=np.uint8)
    assert np.array_equal(cv2.cvtColor(img4, cv2.COLOR_BGRA2GRAY), O.color_to_gray(img4))
    assert np.array_equal(cv2.cvtColor(img4, cv2.COLOR_RGBA2GRAY), O.color_to_gray(img4, rgb=True))


def test_undistort_points_matches_cv2():
    """Frame::UndistortKeyPoints (src/Frame.cc:1507-1553): the restatement of cv::undistortPoints against the real one, bit for bit"""
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(-20, 660, 50000), rng.uniform(-20, 500, 50000)], 1).astype(np.float32)
    cams = [((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),        # TUM1.yaml
            ((520.908620, 521.007327, 325.141442, 249.701764), (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),       # TUM2.yaml
            ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),                   # EuRoC, 4 coefficients
            ((500.0, 500.0, 320.0, 240.0), (0.1, -0.2, 0.001, -0.002, 0.05, 0.01, -0.02, 0.003))]                            # rational model
    for K4, dist in cams:
        K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
        d = np.array(dist, np.float32)
        want = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2)
        got = O.undistort_points(pts, K4, d)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(O.undistort_points(pts, cams[0][0], np.zeros(5, np.float32)), pts)      # mDistCoef(0) == 0: copy
