"""TEST INFRASTRUCTURE ONLY -- CPU oracle for ORB extraction.

Two interchangeable CPU arms, both restating PLVS2::ORBextractor::operator()
(reference: src/ORBextractor.cc:1245-1389):

* ``extract_cv2``  -- the OpenCV primitives the reference calls are executed by the
  real OpenCV in this image (cv2 4.13: resize / FastFeatureDetector per cell /
  GaussianBlur / scalar fastAtan2), exactly in the reference's call pattern; the
  reference's own C++ (octree distribution, descriptor) comes from liboracle.so.
  This is the arm closest to the reference and the one timed as the CPU baseline.
* ``extract_port`` -- everything in C (oracle/orb_oracle.cpp), no cv2.

tests/test_oracle_orb.py pins the C restatements against cv2 bit-exactly.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import ctypes as C
import math
import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_ic_angle.restype = C.c_float
        _lib.orc_ic_angle.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        _lib.orc_orb_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.orc_steer.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


EDGE = 19
PATCH = 31


class Tables:
    """ORBextractor constructor tables (src/ORBextractor.cc:446-523)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8):
        self.nfeatures, self.scale_factor, self.nlevels = nfeatures, np.float32(scale_factor), nlevels
        self.scale = np.zeros(nlevels, np.float32)
        self.inv_scale = np.zeros(nlevels, np.float32)
        self.sigma2 = np.zeros(nlevels, np.float32)
        self.inv_sigma2 = np.zeros(nlevels, np.float32)
        self.quota = np.zeros(nlevels, np.int32)
        self.umax = np.zeros(16, np.int32)
        lib().orc_orb_tables(nfeatures, C.c_float(scale_factor), nlevels, _p(self.scale), _p(self.inv_scale),
                             _p(self.sigma2), _p(self.inv_sigma2), _p(self.quota), _p(self.umax))

    def level_size(self, w0, h0, level):
        w, h = C.c_int(), C.c_int()
        lib().orc_level_size(w0, h0, C.c_float(self.inv_scale[level]), C.byref(w), C.byref(h))
        return w.value, h.value


# ---- C primitives ---------------------------------------------------------------------------

def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
    return dst


def gauss7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gauss7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast_score_map(img, min_th):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_fast_score_map(_p(img), img.shape[1], img.shape[0], img.strides[0], min_th, _p(out), out.strides[0])
    return out


def fast_rect(img, x0, y0, x1, y1, th, cap=65536):
    img = np.ascontiguousarray(img, np.uint8)
    xs, ys, rs = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orc_fast_rect(_p(img), img.strides[0], x0, y0, x1, y1, th, _p(xs), _p(ys), _p(rs), cap)
    assert n <= cap
    return xs[:n].copy(), ys[:n].copy(), rs[:n].copy()


def fast_cells(img, ini_th, min_th, cap=1 << 20):
    img = np.ascontiguousarray(img, np.uint8)
    xs, ys, rs = (np.empty(cap, np.int32) for _ in range(3))
    n = lib().orc_fast_cells(_p(img), img.shape[1], img.shape[0], img.strides[0], ini_th, min_th, _p(xs), _p(ys), _p(rs), cap)
    assert n <= cap
    return xs[:n].copy(), ys[:n].copy(), rs[:n].copy()


def distribute_octree(px, py, resp, min_x, max_x, min_y, max_y, n_want):
    px = np.ascontiguousarray(px, np.float32); py = np.ascontiguousarray(py, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    sel = np.empty(max(len(px), 1), np.int32)
    m = lib().orc_distribute_octree(len(px), _p(px), _p(py), _p(resp), min_x, max_x, min_y, max_y, n_want, _p(sel))
    return sel[:m].copy()


def fast_atan2(y, x):
    return lib().orc_fast_atan2(C.c_float(y), C.c_float(x))


def ic_angle(img, x, y, umax):
    return lib().orc_ic_angle(_p(img), img.strides[0], C.c_float(x), C.c_float(y), _p(umax))


def descriptor(img, x, y, angle):
    d = np.empty(32, np.uint8)
    lib().orc_orb_descriptor(_p(img), img.strides[0], C.c_float(x), C.c_float(y), C.c_float(angle), _p(d))
    return d


def steer(angle):
    a, b = C.c_float(), C.c_float()
    lib().orc_steer(C.c_float(angle), C.byref(a), C.byref(b))
    return a.value, b.value


# ---- whole extractor ------------------------------------------------------------------------

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"),
                     ("octave", "i4"), ("class_id", "i4")])


def extract_port(gray, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, lapping=(0, 0)):
    """All-C arm.  Returns (keypoints[KP_DTYPE], descriptors[n,32] u8, mono_index, n_candidates)."""
    gray = np.ascontiguousarray(gray, np.uint8)
    cap = max(4 * nfeatures, 1024)
    kps = np.empty((cap, 7), np.float32)
    desc = np.empty((cap, 32), np.uint8)
    mono, ncand = C.c_int(), C.c_int()
    n = lib().orc_orb_extract(_p(gray), gray.shape[1], gray.shape[0], gray.strides[0], nfeatures,
                              C.c_float(scale_factor), nlevels, ini_th, min_th, lapping[0], lapping[1],
                              _p(kps), _p(desc), cap, C.byref(mono), C.byref(ncand))
    assert n >= 0, "oracle keypoint capacity exceeded"
    out = np.zeros(n, KP_DTYPE)
    for i, f in enumerate(("x", "y", "size", "angle", "response")):
        out[f] = kps[:n, i]
    out["octave"] = kps[:n, 5].astype(np.int32)
    out["class_id"] = -1
    return out, desc[:n].copy(), mono.value, ncand.value


def pyramid_cv2(gray, tab):
    """ComputePyramid (src/ORBextractor.cc:1481-1506); the 19-px REFLECT_101 frame is
    not materialised: nothing on the extraction path reads it."""
    import cv2
    h0, w0 = gray.shape
    pyr = [np.ascontiguousarray(gray)]
    for l in range(1, tab.nlevels):
        w, h = tab.level_size(w0, h0, l)
        pyr.append(cv2.resize(pyr[l - 1], (w, h), interpolation=cv2.INTER_LINEAR))
    return pyr


def candidates_cv2(img, ini_th, min_th, fast_ini=None, fast_min=None):
    """Per-cell cv::FAST with threshold fallback (src/ORBextractor.cc:867-998)."""
    import cv2
    fast_ini = fast_ini or cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fast_min = fast_min or cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    h, w = img.shape
    W = np.float32(35)
    min_bx = min_by = EDGE - 3
    max_bx, max_by = w - EDGE + 3, h - EDGE + 3
    width, height = np.float32(max_bx - min_bx), np.float32(max_by - min_by)
    xs, ys, rs = [], [], []
    if width <= 0 or height <= 0:
        return xs, ys, rs
    n_cols, n_rows = int(width / W), int(height / W)
    if n_cols == 0 or n_rows == 0:
        return xs, ys, rs
    w_cell, h_cell = int(math.ceil(width / np.float32(n_cols))), int(math.ceil(height / np.float32(n_rows)))
    for i in range(n_rows):
        ini_y = min_by + i * h_cell
        max_y = ini_y + h_cell + 6
        if ini_y >= max_by - 3:
            continue
        max_y = min(max_y, max_by)
        for j in range(n_cols):
            ini_x = min_bx + j * w_cell
            max_x = ini_x + w_cell + 6
            if ini_x >= max_bx - 6:
                continue
            max_x = min(max_x, max_bx)
            cell = img[ini_y:max_y, ini_x:max_x]
            kps = fast_ini.detect(cell, None)
            if len(kps) == 0:
                kps = fast_min.detect(cell, None)
            for kp in kps:
                xs.append(kp.pt[0] + j * w_cell); ys.append(kp.pt[1] + i * h_cell); rs.append(kp.response)
    return xs, ys, rs


def extract_cv2(gray, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, lapping=(0, 0),
                return_internals=False, angle_impl="cv2"):
    """cv2-driven arm of ORBextractor::operator() (src/ORBextractor.cc:1245-1389)."""
    import cv2
    gray = np.ascontiguousarray(gray, np.uint8)
    tab = Tables(nfeatures, scale_factor, nlevels)
    pyr = pyramid_cv2(gray, tab)
    fast_ini = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fast_min = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    per_level, ncand = [], 0
    cand_per_level = []
    for l in range(nlevels):
        img = pyr[l]
        xs, ys, rs = candidates_cv2(img, ini_th, min_th, fast_ini, fast_min)
        ncand += len(xs)
        cand_per_level.append((np.array(xs, np.float32), np.array(ys, np.float32), np.array(rs, np.float32)))
        min_b = EDGE - 3
        sel = distribute_octree(xs, ys, rs, min_b, img.shape[1] - EDGE + 3, min_b, img.shape[0] - EDGE + 3,
                                int(tab.quota[l])) if len(xs) else np.zeros(0, np.int32)
        kp = np.zeros(len(sel), KP_DTYPE)
        if len(sel):
            kp["x"] = np.asarray(xs, np.float32)[sel] + np.float32(min_b)
            kp["y"] = np.asarray(ys, np.float32)[sel] + np.float32(min_b)
            kp["response"] = np.asarray(rs, np.float32)[sel]
        kp["octave"] = l
        kp["size"] = np.float32(int(np.float32(PATCH) * tab.scale[l]))
        kp["class_id"] = -1
        for k in range(len(kp)):     # IC_Angle (:110-137): OpenCV's own scalar fastAtan2, or the pinned C port
            if angle_impl == "cv2":
                kp["angle"][k] = _ic_angle_cv2(img, kp["x"][k], kp["y"][k], tab.umax)
            else:
                kp["angle"][k] = ic_angle(img, kp["x"][k], kp["y"][k], tab.umax)
        per_level.append(kp)
    total = sum(len(k) for k in per_level)
    out = np.zeros(total, KP_DTYPE)
    desc = np.zeros((total, 32), np.uint8)
    mono, stereo = 0, total - 1
    blurred = []
    for l in range(nlevels):
        kp = per_level[l]
        if len(kp) == 0:
            blurred.append(None)
            continue
        blur = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        blurred.append(blur)
        sc = tab.scale[l]
        for k in range(len(kp)):
            d = descriptor(blur, kp["x"][k], kp["y"][k], kp["angle"][k])
            rec = kp[k].copy()
            if l != 0:
                rec["x"] = np.float32(rec["x"] * sc); rec["y"] = np.float32(rec["y"] * sc)
            if lapping[0] <= rec["x"] <= lapping[1]:
                slot = stereo; stereo -= 1
            else:
                slot = mono; mono += 1
            out[slot] = rec
            desc[slot] = d
    if return_internals:
        return out, desc, mono, ncand, dict(pyramid=pyr, blurred=blurred, candidates=cand_per_level, per_level=per_level, tables=tab)
    return out, desc, mono, ncand


def _ic_angle_cv2(img, x, y, umax):
    import cv2
    cx, cy = int(np.rint(x)), int(np.rint(y))
    m01 = m10 = 0
    row = img[cy].astype(np.int64)
    u = np.arange(-15, 16)
    m10 += int((u * row[cx - 15:cx + 16]).sum())
    for v in range(1, 16):
        d = int(umax[v])
        uu = np.arange(-d, d + 1)
        lo = img[cy + v, cx - d:cx + d + 1].astype(np.int64)
        hi = img[cy - v, cx - d:cx + d + 1].astype(np.int64)
        m01 += v * int((lo - hi).sum())
        m10 += int((uu * (lo + hi)).sum())
    return cv2.fastAtan2(float(np.float32(m01)), float(np.float32(m10)))


# ---- the REFERENCE's own ORBextractor.cc, compiled by oracle/ref_build.py (oracle/_ref/liborb_ref.so) -----------------

def ref_available():
    from . import ref_build
    return ref_build.build_orb() is not None


class RefExtractor:
    """PLVS2::ORBextractor from /root/reference/src/ORBextractor.cc (OpenCV stand-in: oracle/cv_standin)."""
    _lib = None

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        if RefExtractor._lib is None:
            from . import ref_build
            so = ref_build.build_orb()
            if so is None:
                raise RuntimeError("oracle/_ref/liborb_ref.so missing and /root/reference not present")
            lib()                                                  # liboracle.so first: liborb_ref.so resolves the primitives there
            RefExtractor._lib = C.CDLL(so, mode=C.RTLD_GLOBAL)
            RefExtractor._lib.ref_orb_create.restype = C.c_void_p
            RefExtractor._lib.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
            RefExtractor._lib.ref_orb_destroy.argtypes = [C.c_void_p]
            RefExtractor._lib.ref_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            RefExtractor._lib.ref_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
            RefExtractor._lib.ref_orb_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = RefExtractor._lib.ref_orb_create(nfeatures, C.c_float(scale_factor), nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "_h", None):
            RefExtractor._lib.ref_orb_destroy(self._h)
            self._h = None

    def tables(self):
        out = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        RefExtractor._lib.ref_orb_tables(self._h, *[_p(a) for a in out])
        return dict(scale=out[0], inv_scale=out[1], sigma2=out[2], inv_sigma2=out[3])

    def __call__(self, gray, lapping=(0, 0)):
        """returns (keypoints[KP_DTYPE], descriptors[n,32], mono_index) like extract_port"""
        gray = np.ascontiguousarray(gray, np.uint8)
        cap = max(4 * self.nfeatures, 1024)
        kps = np.empty((cap, 7), np.float32); desc = np.empty((cap, 32), np.uint8)
        mono = C.c_int()
        n = RefExtractor._lib.ref_orb_extract(self._h, _p(gray), gray.shape[1], gray.shape[0], gray.strides[0], lapping[0], lapping[1],
                                              _p(kps), _p(desc), cap, C.byref(mono))
        assert n >= 0
        out = np.zeros(n, KP_DTYPE)
        for i, f in enumerate(("x", "y", "size", "angle", "response")):
            out[f] = kps[:n, i]
        out["octave"] = kps[:n, 5].astype(np.int32)
        out["class_id"] = kps[:n, 6].astype(np.int32)
        return out, desc[:n].copy(), mono.value

    def level(self, l, filtered=False):
        w, h = C.c_int(), C.c_int()
        RefExtractor._lib.ref_orb_level(self._h, l, int(filtered), None, 0, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        rc = RefExtractor._lib.ref_orb_level(self._h, l, int(filtered), _p(out), out.size, C.byref(w), C.byref(h))
        assert rc == 0
        return out


# ---- step before extraction: cv::cvtColor(.., COLOR_*2GRAY), src/Tracking.cc:1797-1810 ---------------------

def color_to_gray(img, rgb=False):
    """OpenCV 4 fixed-point grey conversion of an 8-bit BGR[A] / RGB[A] image: (R*9798 + G*19235 + B*3735 + 2^14) >> 15.
    Pinned: tests/test_oracle_orb.py compares it with the real cv2.cvtColor for all 2^24 colours."""
    img = np.asarray(img, np.uint8)
    c0, c1, c2 = (img[..., i].astype(np.int64) for i in range(3))
    b, r = (c2, c0) if rgb else (c0, c2)
    return ((r * 9798 + c1 * 19235 + b * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def undistort_points(xy, K4, dist):
    """Frame::UndistortKeyPoints = cv::undistortPoints(pts, K, dist, None, K): xy [n,2] float32, K4 = (fx, fy, cx, cy), dist = float32 coefficients"""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    k = np.ascontiguousarray(K4, np.float32); d = np.ascontiguousarray(dist, np.float32)
    out = np.empty_like(xy)
    lib().orc_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib().orc_undistort_points(_p(xy), len(xy), _p(k), _p(d), len(d), _p(out))
    return out
