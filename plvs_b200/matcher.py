"""Host-side mirror of PLVS2::ORBmatcher for the hot-path overloads (reference:
include/ORBmatcher.h:61-97, src/ORBmatcher.cc).  Frames are flat views (numpy arrays) instead of
the pointer-rich Frame/KeyFrame/MapPoint graph; the C++ shim (shim/) does the same gather from the
reference's objects.  Everything runs through the C ABI of libplvs_b200.so."""
import ctypes as C
import numpy as np

from . import _lib
from .orb import KP_DTYPE

MP_QUERY = np.dtype([("proj_x", "f4"), ("proj_y", "f4"), ("proj_xr", "f4"), ("track_depth", "f4"), ("view_cos", "f4"),
                     ("level", "i4"), ("flags", "u4"), ("desc", "u1", 32)])
LAST_QUERY = np.dtype([("u", "f4"), ("v", "f4"), ("invz", "f4"), ("last_octave", "i4"), ("angle", "f4"),
                       ("flags", "u4"), ("desc", "u1", 32)])
FUSE_QUERY = np.dtype([("u", "f4"), ("v", "f4"), ("ur", "f4"), ("level", "i4"), ("desc", "u1", 32)])
assert MP_QUERY.itemsize == 60 and LAST_QUERY.itemsize == 56 and FUSE_QUERY.itemsize == 48
Q_OBS_POSITIVE = 1
FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48


class Frame:
    """The Frame/KeyFrame members the matchers read (RGB-D, Nleft == -1)."""

    def __init__(self, keys, desc, width, height, scale_factors, level_sigma2=None, uright=None, bf=40.0,
                 bounds=None, device_ptrs=None):
        self.keys = np.ascontiguousarray(keys, KP_DTYPE) if keys is not None else None
        self.desc = np.ascontiguousarray(desc, np.uint8) if desc is not None else None
        self.uright = None if uright is None else np.ascontiguousarray(uright, np.float32)
        self.n = len(self.keys) if device_ptrs is None else device_ptrs[0]
        self.device_ptrs = device_ptrs          # (n, keys_ptr, desc_ptr, uright_ptr|0[, cache_key[, grid_cell_start, grid_sorted]])
        self.scale_factors = np.asarray(scale_factors, np.float32)
        self.level_sigma2 = np.asarray(level_sigma2 if level_sigma2 is not None else self.scale_factors ** 2, np.float32)
        # Frame::ComputeImageBounds without distortion (src/Frame.cc:1770-1776)
        self.min_x, self.min_y, self.max_x, self.max_y = bounds or (0.0, 0.0, float(width), float(height))
        self.grid_inv_w = np.float32(FRAME_GRID_COLS) / np.float32(np.float32(self.max_x) - np.float32(self.min_x))
        self.grid_inv_h = np.float32(FRAME_GRID_ROWS) / np.float32(np.float32(self.max_y) - np.float32(self.min_y))
        self.bf = bf

    def view(self):
        v = _lib.FrameView()
        v.n = self.n
        if self.device_ptrs is not None:
            v.keys, v.desc, v.uright = self.device_ptrs[1], self.device_ptrs[2], self.device_ptrs[3] or None
            v.on_device = 1
            if not self.device_ptrs[3] and self.uright is not None:      # mvuRight computed by the caller on the host
                v.uright = self.uright.ctypes.data
                v.on_device = 1 | 2
            v.cache_key = self.device_ptrs[4] if len(self.device_ptrs) > 4 else 0
            if len(self.device_ptrs) > 6 and self.device_ptrs[5] and self.device_ptrs[6]:      # grid built at frame construction
                v.grid_cell_start, v.grid_sorted = self.device_ptrs[5], self.device_ptrs[6]
        else:
            v.keys = self.keys.ctypes.data
            v.desc = self.desc.ctypes.data
            v.uright = self.uright.ctypes.data if self.uright is not None else None
            v.on_device = 0
        v.min_x, v.min_y, v.max_x, v.max_y = self.min_x, self.min_y, self.max_x, self.max_y
        v.grid_inv_w, v.grid_inv_h = float(self.grid_inv_w), float(self.grid_inv_h)
        nl = len(self.scale_factors)
        for i in range(nl):
            v.scale_factors[i] = float(self.scale_factors[i])
            v.level_sigma2[i] = float(self.level_sigma2[i])
        v.nlevels = nl
        v.bf = self.bf
        return v


def featvec(node_of_feature):
    """Flatten a per-feature node id array into the CSR form of DBoW2::FeatureVector
    (std::map<NodeId, std::vector<unsigned>>: nodes ascending, features ascending inside a node)."""
    node = np.asarray(node_of_feature, np.int64)
    valid = np.nonzero(node >= 0)[0]
    order = valid[np.argsort(node[valid], kind="stable")]
    ids, counts = np.unique(node[order], return_counts=True)
    offsets = np.zeros(len(ids) + 1, np.int32)
    offsets[1:] = np.cumsum(counts)
    return ids.astype(np.uint32), offsets, order.astype(np.int32)


def featvec_struct(fv):
    s = _lib.FeatVec()
    s.n_nodes = len(fv[0])
    s.node_ids, s.offsets, s.features = fv[0].ctypes.data, fv[1].ctypes.data, fv[2].ctypes.data
    return s


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 12

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self._lib = _lib.load()
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        self._h = C.c_void_p()
        _lib.check(self._lib.plvs_match_create(device, C.byref(self._h)), "plvs_match_create")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.plvs_match_destroy(self._h)
            self._h = None

    __del__ = close

    def last_stats(self):
        """(rounds of the claim resolution, kernel launches) of the last projection search on this handle"""
        r, k = C.c_int(), C.c_int()
        _lib.check(self._lib.plvs_match_last_stats(self._h, C.byref(r), C.byref(k)), "plvs_match_last_stats")
        return r.value, k.value

    @staticmethod
    def DescriptorDistance(a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return _lib.load().plvs_hamming256(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))

    def SearchByProjectionMap(self, F, queries, th=3.0, bFarPoints=False, thFarPoints=50.0, claimed=None, nnratio=None):
        """SearchByProjection(Frame&, vector<MapPointPtr>&, th, bFarPoints, thFarPoints) -> (nmatches, assign[N])."""
        q = np.ascontiguousarray(queries, MP_QUERY)
        assign = np.full(max(F.n, 1), -1, np.int32)
        nm = C.c_int()
        v = F.view()
        cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
        rc = self._lib.plvs_match_projection_map(self._h, C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), C.c_float(th),
                                                 C.c_float(self.mfNNratio if nnratio is None else nnratio), int(bFarPoints), C.c_float(thFarPoints),
                                                 cl.ctypes.data_as(C.c_void_p) if cl is not None else None,
                                                 assign.ctypes.data_as(C.c_void_p), C.byref(nm))
        _lib.check(rc, "plvs_match_projection_map")
        return nm.value, assign[:F.n]

    def SearchByProjectionLast(self, Cur, queries, th, bForward=False, bBackward=False, claimed=None):
        """SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) with the last-frame points pre-projected."""
        q = np.ascontiguousarray(queries, LAST_QUERY)
        assign = np.full(max(Cur.n, 1), -1, np.int32)
        nm = C.c_int()
        v = Cur.view()
        cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
        rc = self._lib.plvs_match_projection_last(self._h, C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), C.c_float(th),
                                                  int(bForward), int(bBackward), int(self.mbCheckOrientation),
                                                  cl.ctypes.data_as(C.c_void_p) if cl is not None else None,
                                                  assign.ctypes.data_as(C.c_void_p), C.byref(nm))
        _lib.check(rc, "plvs_match_projection_last")
        return nm.value, assign[:Cur.n]

    def SearchByProjectionReloc(self, Cur, queries, th, ORBdist, claimed=None):
        """SearchByProjection(Frame& Cur, KeyFramePtr& pKF, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1996-2122) with the
        keyframe's map points pre-projected (LAST_QUERY records: last_octave = predicted level, flags = 1)."""
        q = np.ascontiguousarray(queries, LAST_QUERY)
        assign = np.full(max(Cur.n, 1), -1, np.int32)
        nm = C.c_int()
        v = Cur.view()
        cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
        _lib.check(self._lib.plvs_match_projection_reloc(self._h, C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th, int(ORBdist), int(self.mbCheckOrientation),
                                                         cl.ctypes.data_as(C.c_void_p) if cl is not None else None, assign.ctypes.data_as(C.c_void_p), C.byref(nm)),
                   "plvs_match_projection_reloc")
        return nm.value, assign[:Cur.n]

    def SearchByProjectionSim3(self, KF, queries, th, ratioHamming=1.0, matched=None):
        """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (src/ORBmatcher.cc:509-615) with pre-projected points."""
        q = np.ascontiguousarray(queries, LAST_QUERY)
        assign = np.full(max(KF.n, 1), -1, np.int32)
        nm = C.c_int()
        v = KF.view()
        ml = None if matched is None else np.ascontiguousarray(matched, np.uint8)
        _lib.check(self._lib.plvs_match_projection_sim3(self._h, C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th, ratioHamming,
                                                        ml.ctypes.data_as(C.c_void_p) if ml is not None else None, assign.ctypes.data_as(C.c_void_p), C.byref(nm)),
                   "plvs_match_projection_sim3")
        return nm.value, assign[:KF.n]

    def SearchByBoW(self, KF, F, fv_kf, fv_f, has_mp_kf):
        """SearchByBoW(KeyFramePtr&, Frame&, vector<MapPointPtr>&) (src/ORBmatcher.cc:300-506) -> (nmatches, match_f[F.N]) with
        match_f[i] = keyframe feature whose map point goes to frame feature i, or -1."""
        m = np.full(max(F.n, 1), -1, np.int32)
        nm = C.c_int()
        vk, vf = KF.view(), F.view()
        sk, sf = featvec_struct(fv_kf), featvec_struct(fv_f)
        h = np.ascontiguousarray(has_mp_kf, np.uint8)
        _lib.check(self._lib.plvs_match_bow(self._h, C.byref(vk), C.byref(vf), C.byref(sk), C.byref(sf), h.ctypes.data_as(C.c_void_p),
                                            self.mfNNratio, int(self.mbCheckOrientation), m.ctypes.data_as(C.c_void_p), C.byref(nm)), "plvs_match_bow")
        return nm.value, m[:F.n]

    def FuseSim3(self, KF, queries, th):
        """Search part of Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:1437-1553): no chi-square gate."""
        q = np.ascontiguousarray(queries, FUSE_QUERY)
        bi = np.full(max(len(q), 1), -1, np.int32); bd = np.full(max(len(q), 1), 256, np.int32)
        nf = C.c_int()
        v = KF.view()
        _lib.check(self._lib.plvs_match_fuse_sim3(self._h, C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th,
                                                  bi.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p), C.byref(nf)), "plvs_match_fuse_sim3")
        return nf.value, bi[:len(q)], bd[:len(q)]

    def SearchByBoWKF(self, KF1, KF2, fv1, fv2, has_mp1, has_mp2):
        """SearchByBoW(KeyFramePtr&, KeyFramePtr&, vector<MapPointPtr>&) (src/ORBmatcher.cc:853-997) -> (nmatches, match12[N1])."""
        m = np.full(max(KF1.n, 1), -1, np.int32)
        nm = C.c_int()
        v1, v2 = KF1.view(), KF2.view()
        s1, s2 = featvec_struct(fv1), featvec_struct(fv2)
        h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
        _lib.check(self._lib.plvs_match_bow_kf(self._h, C.byref(v1), C.byref(v2), C.byref(s1), C.byref(s2), h1.ctypes.data_as(C.c_void_p),
                                               h2.ctypes.data_as(C.c_void_p), self.mfNNratio, int(self.mbCheckOrientation), m.ctypes.data_as(C.c_void_p),
                                               C.byref(nm)), "plvs_match_bow_kf")
        return nm.value, m[:KF1.n]

    def SearchForInitialization(self, F1, F2, vbPrevMatched, windowSize=100):
        """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:732-852)
        -> (nmatches, vnMatches12[N1], updated vbPrevMatched [N1, 2])."""
        prev = np.array(vbPrevMatched, np.float32).reshape(-1, 2).copy()
        if len(prev) != F1.n:
            raise ValueError("vbPrevMatched needs one point per F1 keypoint")
        m = np.full(max(F1.n, 1), -1, np.int32)
        nm = C.c_int()
        v1, v2 = F1.view(), F2.view()
        _lib.check(self._lib.plvs_match_initialization(self._h, C.byref(v1), C.byref(v2), prev.ctypes.data_as(C.c_void_p), int(windowSize), self.mfNNratio,
                                                       int(self.mbCheckOrientation), m.ctypes.data_as(C.c_void_p), C.byref(nm)), "plvs_match_initialization")
        return nm.value, m[:F1.n], prev

    def ComputeDistinctiveDescriptors(self, desc_lists):
        """MapPoint::ComputeDistinctiveDescriptors for a batch of map points: desc_lists = one [n_i, 32] uint8 array per point
        -> best index per point (src/MapPoint.cc:428-455)."""
        off = np.zeros(len(desc_lists) + 1, np.int32)
        off[1:] = np.cumsum([len(d) for d in desc_lists])
        flat = np.ascontiguousarray(np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in desc_lists]) if off[-1] else np.zeros((0, 32), np.uint8))
        best = np.full(max(len(desc_lists), 1), -1, np.int32)
        _lib.check(self._lib.plvs_distinctive_descriptors(self._h, flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(desc_lists),
                                                          best.ctypes.data_as(C.c_void_p)), "plvs_distinctive_descriptors")
        return best[:len(desc_lists)]

    def InFrustum(self, frustum, map_points):
        """Frame::isInFrustum over a local map (src/Frame.cc:955-1017): `frustum` = a ctypes struct laid out like plvs_frustum, map_points =
        array of plvs_map_point records -> (n_in_view, queries[MP_QUERY], in_view[uint8])."""
        pts = np.ascontiguousarray(map_points)
        assert pts.dtype.itemsize == 68
        q = np.zeros(max(len(pts), 1), MP_QUERY); iv = np.zeros(max(len(pts), 1), np.uint8)
        n = C.c_int()
        _lib.check(self._lib.plvs_match_in_frustum(self._h, C.byref(frustum), pts.ctypes.data_as(C.c_void_p), len(pts), q.ctypes.data_as(C.c_void_p),
                                                   iv.ctypes.data_as(C.c_void_p), C.byref(n)), "plvs_match_in_frustum")
        return n.value, q[:len(pts)], iv[:len(pts)]

    def SearchByProjectionMapResident(self, F, th=3.0, bFarPoints=False, thFarPoints=50.0, claimed=None, nnratio=None):
        """SearchByProjection(F, vpMapPoints, ...) on the queries the last InFrustum call of this matcher left on the device.
        -> (nmatches, assign[N]) with assign[i] = index into the map-point array given to InFrustum."""
        assign = np.full(max(F.n, 1), -1, np.int32)
        nm = C.c_int()
        v = F.view()
        cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
        _lib.check(self._lib.plvs_match_projection_map_resident(self._h, C.byref(v), th, self.mfNNratio if nnratio is None else nnratio, int(bFarPoints), thFarPoints,
                                                                cl.ctypes.data_as(C.c_void_p) if cl is not None else None, assign.ctypes.data_as(C.c_void_p), C.byref(nm)),
                   "plvs_match_projection_map_resident")
        return nm.value, assign[:F.n]

    def SearchBySim3(self, KF1, KF2, q12, q21, valid1, valid2, th):
        """SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (src/ORBmatcher.cc:1555-1772): q12[i1] = map point i1 of KF1 projected into
        KF2 (FUSE_QUERY; `ur` unused), q21 the reverse; valid* = map point present, not bad, not already matched.  Two device searches
        (the Fuse-Sim3 search, accepted if <= TH_HIGH) and the mutual-agreement check -> (nFound, match12[N1])."""
        _, bi12, bd12 = self.FuseSim3(KF2, q12, th)
        _, bi21, bd21 = self.FuseSim3(KF1, q21, th)
        vn1 = np.where(np.asarray(valid1, bool) & (bd12 <= 100), bi12, -1)
        vn2 = np.where(np.asarray(valid2, bool) & (bd21 <= 100), bi21, -1)
        m12 = np.full(len(vn1), -1, np.int32)
        ok = vn1 >= 0
        back = np.where(ok, vn2[np.where(ok, vn1, 0)], -2)
        sel = ok & (back == np.arange(len(vn1)))
        m12[sel] = vn1[sel]
        return int(sel.sum()), m12

    def Fuse(self, KF, queries, th=3.0, inv_level_sigma2=None):
        """Search part of Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:1340-1406) -> (nFused, best_idx[nq], best_dist[nq]).
        The caller applies bestDist<=TH_LOW and the Replace/AddObservation bookkeeping in query order."""
        q = np.ascontiguousarray(queries, FUSE_QUERY)
        inv = np.ascontiguousarray(inv_level_sigma2 if inv_level_sigma2 is not None else 1.0 / KF.level_sigma2, np.float32)
        bi = np.full(max(len(q), 1), -1, np.int32); bd = np.full(max(len(q), 1), 256, np.int32)
        nf = C.c_int()
        v = KF.view()
        _lib.check(self._lib.plvs_match_fuse(self._h, C.byref(v), inv.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), len(q), th,
                                             bi.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p), C.byref(nf)), "plvs_match_fuse")
        return nf.value, bi[:len(q)], bd[:len(q)]

    def SearchForTriangulation(self, KF1, KF2, fv1, fv2, has_mp1, has_mp2, F12, ep, bOnlyStereo=False, bCoarse=False):
        """-> (nmatches, vMatches12[N1]); vMatchedPairs = [(i, m) for i, m in enumerate(vMatches12) if m >= 0]."""
        m12 = np.full(max(KF1.n, 1), -1, np.int32)
        nm = C.c_int()
        v1, v2 = KF1.view(), KF2.view()
        s1, s2 = featvec_struct(fv1), featvec_struct(fv2)
        h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
        F = np.ascontiguousarray(F12, np.float32).reshape(9); e = np.ascontiguousarray(ep, np.float32)
        rc = self._lib.plvs_match_triangulation(self._h, C.byref(v1), C.byref(v2), C.byref(s1), C.byref(s2),
                                                h1.ctypes.data_as(C.c_void_p), h2.ctypes.data_as(C.c_void_p),
                                                F.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                                int(bOnlyStereo), int(bCoarse), int(self.mbCheckOrientation),
                                                m12.ctypes.data_as(C.c_void_p), C.byref(nm))
        _lib.check(rc, "plvs_match_triangulation")
        return nm.value, m12[:KF1.n]


class LineMatcher:
    """The data-parallel part of PLVS2::LineMatcher (include/LineMatcher.h): ComputeDescriptorMatches = 2-NN over 256-bit LBD descriptors with the
    tie order of the vendored multi-index hashing + the ratio test (src/LineMatcher.cc:2567-2615)."""

    def __init__(self, nnratio=0.78, device=0):
        self._m = ORBmatcher(nnratio, False, device=device)
        self.mfNNratio = float(nnratio)

    def ComputeDescriptorMatches(self, ldesc_q, ldesc_t, queryMask=None):
        """-> (numValidMatches, query_idx[M], train_idx[M, 2], distance[M, 2], vValidMatch[M]): lmatches[i] = (DMatch(query_idx[i], train_idx[i, 0], distance[i, 0]),
        DMatch(query_idx[i], train_idx[i, 1], distance[i, 1]))"""
        q = np.ascontiguousarray(ldesc_q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(ldesc_t, np.uint8).reshape(-1, 32)
        nq = len(q)
        mk = None if queryMask is None else np.ascontiguousarray(queryMask, np.uint8).reshape(-1)
        qi = np.zeros(max(nq, 1), np.int32); ti = np.zeros((max(nq, 1), 2), np.int32); di = np.zeros((max(nq, 1), 2), np.float32); vi = np.zeros(max(nq, 1), np.uint8)
        rows, nv = C.c_int(), C.c_int()
        _lib.check(self._m._lib.plvs_line_knn2(self._m._h, q.ctypes.data, nq, t.ctypes.data, len(t), None if mk is None else mk.ctypes.data, self.mfNNratio,
                                               qi.ctypes.data, ti.ctypes.data, di.ctypes.data, vi.ctypes.data, C.byref(rows), C.byref(nv)), "plvs_line_knn2")
        m = rows.value
        return nv.value, qi[:m].copy(), ti[:m].copy(), di[:m].copy(), vi[:m].copy()


def ComputeStereoMatches(matcher, left, right, pyr_left, pyr_right, inv_scale, mb, mbf):
    """Frame::ComputeStereoMatches (src/Frame.cc:1780): left/right are Frame views whose keys are mvKeys / mvKeysRight,
    pyr_* the extractors' device pyramid views.  Returns (mvuRight, mvDepth, number of stereo points kept)."""
    lib = _lib.load()
    ur = np.full(max(left.n, 1), -1, np.float32); dp = np.full(max(left.n, 1), -1, np.float32)
    vl, vr = left.view(), right.view()
    inv = np.ascontiguousarray(inv_scale, np.float32)
    nv = C.c_int()
    rc = lib.plvs_stereo_match(matcher._h, C.byref(vl), C.byref(vr), C.byref(pyr_left), C.byref(pyr_right), inv.ctypes.data_as(C.c_void_p),
                               C.c_float(mb), C.c_float(mbf), ur.ctypes.data_as(C.c_void_p), dp.ctypes.data_as(C.c_void_p), C.byref(nv))
    _lib.check(rc, "plvs_stereo_match")
    return ur[:left.n], dp[:left.n], nv.value
