"""Host-side mirror of PLVS2::ORBextractor (reference: include/ORBextractor.h:59-170).

Same constructor arguments, same call semantics (``__call__`` == ``operator()``: returns the
mono index, fills keypoints + descriptors, honours vLappingArea, returns -1 on an empty image)
and the same getters, on top of the C ABI of libplvs_b200.so.  Used by tests/bench; the C++ shim in
shim/ is the drop-in for the reference's C++ callers.
"""
import ctypes as C
import numpy as np

from . import _lib

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"),
                     ("octave", "i4"), ("class_id", "i4")])
assert KP_DTYPE.itemsize == C.sizeof(_lib.Keypoint) == 28


class ORBextractor:
    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0):
        self._lib = _lib.load()
        self.nfeatures, self.scaleFactor, self.nlevels = nfeatures, scaleFactor, nlevels
        self.iniThFAST, self.minThFAST = iniThFAST, minThFAST
        prm = _lib.OrbParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        self._h = C.c_void_p()
        _lib.check(self._lib.plvs_orb_create(C.byref(prm), device, C.byref(self._h)), "plvs_orb_create")
        self._cap = max(2 * nfeatures + 64 * nlevels, 256)
        t = [np.zeros(nlevels, np.float32) for _ in range(4)] + [np.zeros(nlevels, np.int32)]
        _lib.check(self._lib.plvs_orb_tables(self._h, *[a.ctypes.data_as(C.c_void_p) for a in t]), "plvs_orb_tables")
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2, self.mnFeaturesPerLevel = t

    def close(self):
        if getattr(self, "_h", None):
            self._lib.plvs_orb_destroy(self._h)
            self._h = None

    __del__ = close

    # getters (include/ORBextractor.h:90-113)
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self.mvScaleFactor.copy()
    def GetInverseScaleFactors(self): return self.mvInvScaleFactor.copy()
    def GetScaleSigmaSquares(self): return self.mvLevelSigma2.copy()
    def GetInverseScaleSigmaSquares(self): return self.mvInvLevelSigma2.copy()

    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """operator()(image, mask /*ignored*/, keypoints, descriptors, vLappingArea).
        Returns (monoIndex, keypoints[KP_DTYPE], descriptors[n,32])."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        mono, kps, desc = self.extract_batch(image[None], vLappingArea)
        return mono[0], kps[0], desc[0]

    def extract_batch(self, images, vLappingArea=(0, 0), pinned_out=None):
        """images: (B,H,W) uint8 host array (or a (ptr, B, H, W, stride, frame_stride) device tuple)."""
        on_device = isinstance(images, tuple)
        if on_device:
            ptr, B, H, W, stride, fstride = images
        else:
            assert images.dtype == np.uint8 and images.ndim == 3
            images = np.ascontiguousarray(images)
            B, H, W = images.shape
            ptr, stride, fstride = images.ctypes.data, images.strides[1], images.strides[0]
        cap = self._cap
        if pinned_out is not None:
            kps, desc = pinned_out
        else:
            kps = np.empty((B, cap), KP_DTYPE)
            desc = np.empty((B, cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        mono = np.zeros(B, np.int32)
        rc = self._lib.plvs_orb_extract_batch(self._h, B, C.c_void_p(ptr), W, H, stride, fstride, int(on_device),
                                              int(vLappingArea[0]), int(vLappingArea[1]),
                                              kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap,
                                              n.ctypes.data_as(C.c_void_p), mono.ctypes.data_as(C.c_void_p))
        _lib.check(rc, "plvs_orb_extract_batch")
        return mono, [kps[b, :n[b]] for b in range(B)], [desc[b, :n[b]] for b in range(B)]

    def extract_batch_color(self, images, rgb=False, vLappingArea=(0, 0)):
        """images: (B,H,W,3|4) uint8 host array; cv::cvtColor(COLOR_{BGR|RGB}[A]2GRAY) runs on the device in front of the pyramid
        (src/Tracking.cc:1797-1810)."""
        assert images.dtype == np.uint8 and images.ndim == 4 and images.shape[3] in (3, 4)
        images = np.ascontiguousarray(images)
        B, H, W, nch = images.shape
        cap = self._cap
        kps = np.empty((B, cap), KP_DTYPE); desc = np.empty((B, cap, 32), np.uint8)
        n = np.zeros(B, np.int32); mono = np.zeros(B, np.int32)
        rc = self._lib.plvs_orb_extract_batch_color(self._h, B, C.c_void_p(images.ctypes.data), W, H, images.strides[1], images.strides[0], nch, int(rgb), 0,
                                                    int(vLappingArea[0]), int(vLappingArea[1]), kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap,
                                                    n.ctypes.data_as(C.c_void_p), mono.ctypes.data_as(C.c_void_p))
        _lib.check(rc, "plvs_orb_extract_batch_color")
        return mono, [kps[b, :n[b]] for b in range(B)], [desc[b, :n[b]] for b in range(B)]

    def ComputeStereoFromRGBD(self, depth, bf, frame=0, n=None, keys_un_x=None):
        """Frame::ComputeStereoFromRGBD (src/Frame.cc:2251-2279) on the device-resident keypoints of `frame` of the last batch.
        depth: (H,W) float32 host array.  Returns (mvuRight, mvDepth, device pointer of mvuRight)."""
        depth = np.ascontiguousarray(depth, np.float32)
        n = self.device_result(frame).n if n is None else n
        ur = np.empty(max(n, 1), np.float32); dz = np.empty(max(n, 1), np.float32)
        dptr = C.c_void_p()
        un = None if keys_un_x is None else np.ascontiguousarray(keys_un_x, np.float32)
        rc = self._lib.plvs_orb_stereo_from_rgbd(self._h, frame, depth.ctypes.data_as(C.c_void_p), depth.shape[1], depth.shape[0], depth.strides[0], 0, bf,
                                                 un.ctypes.data_as(C.c_void_p) if un is not None else None, ur.ctypes.data_as(C.c_void_p), dz.ctypes.data_as(C.c_void_p),
                                                 C.byref(dptr))
        _lib.check(rc, "plvs_orb_stereo_from_rgbd")
        return ur[:n], dz[:n], dptr.value

    def UndistortKeyPoints(self, K4, dist, frame=0):
        """Frame::UndistortKeyPoints (src/Frame.cc:1507-1553) on the device-resident keypoints of `frame`: K4 = (fx, fy, cx, cy), dist = mDistCoef.
        Returns (mvKeysUn[KP_DTYPE], device pointer of mvKeysUn)."""
        n = self.device_result(frame).n
        k = np.ascontiguousarray(K4, np.float32); d = np.ascontiguousarray(dist, np.float32)
        out = np.empty(max(n, 1), KP_DTYPE)
        dptr = C.c_void_p()
        _lib.check(self._lib.plvs_orb_undistort(self._h, frame, k.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), len(d), out.ctypes.data_as(C.c_void_p),
                                                C.byref(dptr)), "plvs_orb_undistort")
        return out[:n], dptr.value

    def pyramid_level(self, level, blurred=False, frame=0):
        """Host copy of mvImagePyramid[level] / mvImagePyramidFiltered[level] of the last extract."""
        w, h, pitch, dptr = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
        _lib.check(self._lib.plvs_orb_pyramid_level(self._h, frame, level, int(blurred), C.byref(dptr), C.byref(w), C.byref(h), C.byref(pitch)),
                   "plvs_orb_pyramid_level")
        out = np.empty((h.value, w.value), np.uint8)
        _lib.check(self._lib.plvs_orb_download_level(self._h, frame, level, int(blurred), out.ctypes.data_as(C.c_void_p), out.strides[0]),
                   "plvs_orb_download_level")
        return out

    def candidates(self, level, frame=0, cap=1 << 20):
        """FAST candidates of one level in DistributeOctTree input order: (x, y, score) int32 arrays."""
        x, y, s = (np.empty(cap, np.int32) for _ in range(3))
        n = C.c_int()
        _lib.check(self._lib.plvs_orb_candidates(self._h, frame, level, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                                 s.ctypes.data_as(C.c_void_p), cap, C.byref(n)), "plvs_orb_candidates")
        return x[:n.value].copy(), y[:n.value].copy(), s[:n.value].copy()

    def pyramid_view(self, frame=0, blurred=False):
        v = _lib.PyramidView()
        _lib.check(self._lib.plvs_orb_pyramid_view(self._h, frame, int(blurred), C.byref(v)), "plvs_orb_pyramid_view")
        return v

    def set_frame_grid(self, width=None, height=None, bounds=None):
        """Frame::AssignFeaturesToGrid at frame construction (src/Frame.cc:598,716): from now on every extraction also bins the keypoints of
        each frame into the 64 x 48 grid over `bounds` = (mnMinX, mnMinY, mnMaxX, mnMaxY) (default: the image, i.e. no distortion), and
        device_result() carries the grid.  set_frame_grid() without arguments turns it off."""
        if width is None and bounds is None:
            _lib.check(self._lib.plvs_orb_set_frame_grid(self._h, None), "plvs_orb_set_frame_grid")
            return
        x0, y0, x1, y1 = bounds or (0.0, 0.0, float(width), float(height))
        b = np.array([x0, y0, x1, y1, np.float32(64) / np.float32(np.float32(x1) - np.float32(x0)),
                      np.float32(48) / np.float32(np.float32(y1) - np.float32(y0))], np.float32)
        _lib.check(self._lib.plvs_orb_set_frame_grid(self._h, b.ctypes.data_as(C.c_void_p)), "plvs_orb_set_frame_grid")

    def device_result(self, frame=0):
        v = _lib.OrbDeviceView()
        _lib.check(self._lib.plvs_orb_device_result(self._h, frame, C.byref(v)), "plvs_orb_device_result")
        return v

    def last_stats(self):
        s = _lib.OrbStats()
        _lib.check(self._lib.plvs_orb_last_stats(self._h, C.byref(s)), "plvs_orb_last_stats")
        return {f: getattr(s, f) for f, _ in s._fields_}
