"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/match_oracle.cpp (sequential CPU restatement
of the three ORBmatcher searches).  Takes the same flat views as plvs_b200.matcher."""
import ctypes as C
import numpy as np

from .orb import lib
from plvs_b200 import _lib as _abi
from plvs_b200.matcher import MP_QUERY, LAST_QUERY, FUSE_QUERY, featvec_struct


def _setup():
    l = lib()
    l.orc_search_by_projection_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    l.orc_search_by_projection_last.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    l.orc_search_for_triangulation.argtypes = [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_int, C.c_void_p]
    l.orc_hamming256.argtypes = [C.c_void_p, C.c_void_p]
    return l


def hamming256(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return _setup().orc_hamming256(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))


def search_by_projection_map(F, queries, th, nn_ratio, far_points=False, th_far=50.0, claimed=None):
    l = _setup()
    q = np.ascontiguousarray(queries, MP_QUERY)
    v = F.view()
    assign = np.full(max(F.n, 1), -1, np.int32)
    cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
    n = l.orc_search_by_projection_map(C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th, nn_ratio, int(far_points), th_far,
                                       cl.ctypes.data_as(C.c_void_p) if cl is not None else None, assign.ctypes.data_as(C.c_void_p))
    return n, assign[:F.n]


def search_by_projection_last(Cur, queries, th, forward=False, backward=False, check_ori=True, claimed=None):
    l = _setup()
    q = np.ascontiguousarray(queries, LAST_QUERY)
    v = Cur.view()
    assign = np.full(max(Cur.n, 1), -1, np.int32)
    cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
    n = l.orc_search_by_projection_last(C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th, int(forward), int(backward), int(check_ori),
                                        cl.ctypes.data_as(C.c_void_p) if cl is not None else None, assign.ctypes.data_as(C.c_void_p))
    return n, assign[:Cur.n]


def search_for_triangulation(K1, K2, fv1, fv2, has1, has2, F12, ep, only_stereo=False, coarse=False, check_ori=True):
    l = _setup()
    v1, v2 = K1.view(), K2.view()
    s1, s2 = featvec_struct(fv1), featvec_struct(fv2)
    h1 = np.ascontiguousarray(has1, np.uint8); h2 = np.ascontiguousarray(has2, np.uint8)
    F = np.ascontiguousarray(F12, np.float32).reshape(9); e = np.ascontiguousarray(ep, np.float32)
    m12 = np.full(max(K1.n, 1), -1, np.int32)
    n = l.orc_search_for_triangulation(C.byref(v1), C.byref(v2), C.byref(s1), C.byref(s2), h1.ctypes.data_as(C.c_void_p),
                                       h2.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                       int(only_stereo), int(coarse), int(check_ori), m12.ctypes.data_as(C.c_void_p))
    return n, m12[:K1.n]


class _Level(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("stride", C.c_int)]


def compute_stereo_matches(left, right, pyr_left, pyr_right, scale, inv_scale, mb, mbf):
    """oracle of Frame::ComputeStereoMatches; pyr_* are lists of contiguous uint8 level images (unblurred)."""
    l = _setup()
    nl = len(pyr_left)
    L = (_Level * nl)(); R = (_Level * nl)()
    keep = []
    for i in range(nl):
        a = np.ascontiguousarray(pyr_left[i], np.uint8); b = np.ascontiguousarray(pyr_right[i], np.uint8); keep += [a, b]
        L[i] = _Level(a.ctypes.data, a.shape[1], a.shape[0], a.strides[0]); R[i] = _Level(b.ctypes.data, b.shape[1], b.shape[0], b.strides[0])
    sc = np.ascontiguousarray(scale, np.float32); isc = np.ascontiguousarray(inv_scale, np.float32)
    ur = np.full(max(left.n, 1), -1, np.float32); dp = np.full(max(left.n, 1), -1, np.float32)
    l.orc_compute_stereo_matches.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    kept = l.orc_compute_stereo_matches(left.n, left.keys.ctypes.data, left.desc.ctypes.data, right.n, right.keys.ctypes.data, right.desc.ctypes.data,
                                        C.cast(L, C.c_void_p), C.cast(R, C.c_void_p), sc.ctypes.data, isc.ctypes.data, mb, mbf,
                                        ur.ctypes.data, dp.ctypes.data)
    return ur[:left.n], dp[:left.n], kept


# ---- the REFERENCE's own ORBmatcher.cc, compiled by oracle/ref_build.py (oracle/_ref/libmatch_ref.so) -----------------
_ref = None


def ref_available():
    from . import ref_build
    return ref_build.build_match() is not None


def _ref_lib():
    global _ref
    if _ref is None:
        from . import ref_build
        so = ref_build.build_match()
        if so is None:
            raise RuntimeError("oracle/_ref/libmatch_ref.so missing and /root/reference not present")
        lib()
        _ref = C.CDLL(so)
        _ref.ref_search_by_projection_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        _ref.ref_search_by_projection_last.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _ref.ref_search_for_triangulation.argtypes = [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_int, C.c_void_p]
        _ref.ref_hamming256.argtypes = [C.c_void_p, C.c_void_p]
    return _ref


def ref_hamming256(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return _ref_lib().ref_hamming256(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))


def ref_search_by_projection_map(F, queries, th, nn_ratio, far_points=False, th_far=50.0, claimed=None):
    l = _ref_lib()
    q = np.ascontiguousarray(queries, MP_QUERY)
    v = F.view()
    assign = np.full(max(F.n, 1), -1, np.int32)
    cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
    n = l.ref_search_by_projection_map(C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th, nn_ratio, int(far_points), th_far,
                                       cl.ctypes.data_as(C.c_void_p) if cl is not None else None, assign.ctypes.data_as(C.c_void_p))
    return n, assign[:F.n]


def canonical_last_queries(queries):
    """The reference derives invzc from the camera-frame depth (`1.0/x3Dc(2)`, src/ORBmatcher.cc:1811); the flat query carries
    invz.  Returns (queries', z) with z a float depth and queries'.invz == float(1.0/z), so both sides see the same numbers."""
    q = np.array(queries, LAST_QUERY, copy=True)
    with np.errstate(divide="ignore"):
        z = (1.0 / q["invz"].astype(np.float64)).astype(np.float32)
        q["invz"] = (1.0 / z.astype(np.float64)).astype(np.float32)
    return q, z


def ref_search_by_projection_last(Cur, queries, z, th, forward=False, backward=False, check_ori=True, claimed=None):
    l = _ref_lib()
    q = np.ascontiguousarray(queries, LAST_QUERY)
    z = np.ascontiguousarray(z, np.float32)
    v = Cur.view()
    assign = np.full(max(Cur.n, 1), -1, np.int32)
    cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
    n = l.ref_search_by_projection_last(C.byref(v), q.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), len(q), th, int(forward), int(backward),
                                        int(check_ori), cl.ctypes.data_as(C.c_void_p) if cl is not None else None, assign.ctypes.data_as(C.c_void_p))
    return n, assign[:Cur.n]


def ref_search_for_triangulation(K1, K2, fv1, fv2, has1, has2, F12, ep, only_stereo=False, coarse=False, check_ori=True):
    l = _ref_lib()
    v1, v2 = K1.view(), K2.view()
    s1, s2 = featvec_struct(fv1), featvec_struct(fv2)
    h1 = np.ascontiguousarray(has1, np.uint8); h2 = np.ascontiguousarray(has2, np.uint8)
    F = np.ascontiguousarray(F12, np.float32).reshape(9); e = np.ascontiguousarray(ep, np.float32)
    m12 = np.full(max(K1.n, 1), -1, np.int32)
    n = l.ref_search_for_triangulation(C.byref(v1), C.byref(v2), C.byref(s1), C.byref(s2), h1.ctypes.data_as(C.c_void_p),
                                       h2.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                                       int(only_stereo), int(coarse), int(check_ori), m12.ctypes.data_as(C.c_void_p))
    return n, m12[:K1.n]


# ---- ORBmatcher::Fuse (search part) ------------------------------------------------------------------------------------

def fuse_queries(u, v, z, level, desc, bf):
    """queries as the caller-side shim builds them: ur = u - bf * invz with `const float invz = 1/p3Dc(2)` (src/ORBmatcher.cc:1303,1315)"""
    q = np.zeros(len(u), FUSE_QUERY)
    z = np.asarray(z, np.float32)
    invz = np.float32(1) / z
    q["u"] = u; q["v"] = v; q["ur"] = np.asarray(u, np.float32) - np.float32(bf) * invz
    q["level"] = level; q["desc"] = desc
    return q


def fuse(K, queries, th, inv_level_sigma2):
    l = _setup()
    q = np.ascontiguousarray(queries, FUSE_QUERY)
    inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
    v = K.view()
    bi = np.full(max(len(q), 1), -1, np.int32); bd = np.full(max(len(q), 1), 256, np.int32)
    l.orc_fuse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    n = l.orc_fuse(C.byref(v), inv.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), len(q), th, bi.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p))
    return n, bi[:len(q)], bd[:len(q)]


def ref_fuse(K, queries, z, bf, th, inv_level_sigma2):
    """the reference's own Fuse: returns (nFused, fused_idx[nq]) with fused_idx = keypoint the map point was fused into or -1"""
    l = _ref_lib()
    q = np.ascontiguousarray(queries, FUSE_QUERY)
    inv = np.ascontiguousarray(inv_level_sigma2, np.float32); z = np.ascontiguousarray(z, np.float32)
    v = K.view()
    out = np.full(max(len(q), 1), -1, np.int32)
    l.ref_fuse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_float, C.c_void_p]
    n = l.ref_fuse(C.byref(v), inv.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), bf, len(q), th, out.ctypes.data_as(C.c_void_p))
    return n, out[:len(q)]


# ---- ORBmatcher::SearchByBoW(KF, F) ------------------------------------------------------------------------------------

def _bow_call(fn, K, F, fvK, fvF, has_mp, nn_ratio, check_ori):
    vK, vF = K.view(), F.view()
    sK, sF = featvec_struct(fvK), featvec_struct(fvF)
    h = np.ascontiguousarray(has_mp, np.uint8)
    m = np.full(max(F.n, 1), -1, np.int32)
    fn.argtypes = [C.c_void_p] * 5 + [C.c_float, C.c_int, C.c_void_p]
    n = fn(C.byref(vK), C.byref(vF), C.byref(sK), C.byref(sF), h.ctypes.data_as(C.c_void_p), nn_ratio, int(check_ori), m.ctypes.data_as(C.c_void_p))
    return n, m[:F.n]


def search_by_bow(K, F, fvK, fvF, has_mp, nn_ratio=0.7, check_ori=True):
    return _bow_call(_setup().orc_search_by_bow, K, F, fvK, fvF, has_mp, nn_ratio, check_ori)


def ref_search_by_bow(K, F, fvK, fvF, has_mp, nn_ratio=0.7, check_ori=True):
    return _bow_call(_ref_lib().ref_search_by_bow, K, F, fvK, fvF, has_mp, nn_ratio, check_ori)


# ---- ORBmatcher::SearchByProjection(Cur, KF, sAlreadyFound, th, ORBdist) (relocalisation) -----------------------------

def _reloc_call(fn, Cur, queries, th, orb_dist, check_ori, claimed):
    q = np.ascontiguousarray(queries, LAST_QUERY)
    v = Cur.view()
    assign = np.full(max(Cur.n, 1), -1, np.int32)
    cl = None if claimed is None else np.ascontiguousarray(claimed, np.uint8)
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = fn(C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th, int(orb_dist), int(check_ori), cl.ctypes.data_as(C.c_void_p) if cl is not None else None,
           assign.ctypes.data_as(C.c_void_p))
    return n, assign[:Cur.n]


def search_by_projection_reloc(Cur, queries, th, orb_dist, check_ori=True, claimed=None):
    return _reloc_call(_setup().orc_search_by_projection_reloc, Cur, queries, th, orb_dist, check_ori, claimed)


def ref_search_by_projection_reloc(Cur, queries, th, orb_dist, check_ori=True, claimed=None):
    return _reloc_call(_ref_lib().ref_search_by_projection_reloc, Cur, queries, th, orb_dist, check_ori, claimed)


def fuse_sim3(K, queries, th):
    l = _setup()
    q = np.ascontiguousarray(queries, FUSE_QUERY)
    v = K.view()
    bi = np.full(max(len(q), 1), -1, np.int32); bd = np.full(max(len(q), 1), 256, np.int32)
    l.orc_fuse_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    n = l.orc_fuse_sim3(C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th, bi.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p))
    return n, bi[:len(q)], bd[:len(q)]


def ref_fuse_sim3(K, queries, z, th, pre_mp=None):
    l = _ref_lib()
    q = np.ascontiguousarray(queries, FUSE_QUERY); z = np.ascontiguousarray(z, np.float32)
    v = K.view()
    out = np.full(max(len(q), 1), -1, np.int32)
    pm = None if pre_mp is None else np.ascontiguousarray(pre_mp, np.uint8)
    l.ref_fuse_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    n = l.ref_fuse_sim3(C.byref(v), q.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), len(q), th,
                        pm.ctypes.data_as(C.c_void_p) if pm is not None else None, out.ctypes.data_as(C.c_void_p))
    return n, out[:len(q)]


def _sim3_call(fn, K, queries, th, ratio, matched):
    q = np.ascontiguousarray(queries, LAST_QUERY)
    v = K.view()
    assign = np.full(max(K.n, 1), -1, np.int32)
    ml = None if matched is None else np.ascontiguousarray(matched, np.uint8)
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    n = fn(C.byref(v), q.ctypes.data_as(C.c_void_p), len(q), th, ratio, ml.ctypes.data_as(C.c_void_p) if ml is not None else None, assign.ctypes.data_as(C.c_void_p))
    return n, assign[:K.n]


def search_by_projection_sim3(K, queries, th, ratio=1.0, matched=None):
    return _sim3_call(_setup().orc_search_by_projection_sim3, K, queries, th, ratio, matched)


def ref_search_by_projection_sim3(K, queries, th, ratio=1.0, matched=None):
    """th is an int in the reference's signature (src/ORBmatcher.cc:509)"""
    return _sim3_call(_ref_lib().ref_search_by_projection_sim3, K, queries, th, ratio, matched)


# ---- ORBmatcher::SearchBySim3: two projection searches (the Fuse-Sim3 search with TH_HIGH) + mutual agreement --------------

def sim3_agreement(bi12, bd12, bi21, bd21, valid1, valid2):
    """the host part of SearchBySim3 (src/ORBmatcher.cc:1590-1769): vnMatch1/2 from the two searches (accepted if <= TH_HIGH),
    then the mutual-consistency check.  valid* = map point present (and not already matched)."""
    vn1 = np.where(valid1 & (bd12 <= 100), bi12, -1)
    vn2 = np.where(valid2 & (bd21 <= 100), bi21, -1)
    m12 = np.full(len(vn1), -1, np.int32)
    for i1, i2 in enumerate(vn1):
        if i2 >= 0 and vn2[i2] == i1:
            m12[i1] = i2
    return int((m12 >= 0).sum()), m12


def search_by_sim3(K1, K2, q12, q21, has1, has2, th, fuse=fuse_sim3):
    _, bi12, bd12 = fuse(K2, q12, th)
    _, bi21, bd21 = fuse(K1, q21, th)
    return sim3_agreement(bi12, bd12, bi21, bd21, np.asarray(has1, bool), np.asarray(has2, bool))


def ref_search_by_sim3(K1, K2, q12, q21, has1, has2, th):
    l = _ref_lib()
    a = np.ascontiguousarray(q12, FUSE_QUERY); b = np.ascontiguousarray(q21, FUSE_QUERY)
    h1 = np.ascontiguousarray(has1, np.uint8); h2 = np.ascontiguousarray(has2, np.uint8)
    v1, v2 = K1.view(), K2.view()
    m = np.full(max(K1.n, 1), -1, np.int32)
    l.ref_search_by_sim3.argtypes = [C.c_void_p] * 6 + [C.c_float, C.c_void_p]
    n = l.ref_search_by_sim3(C.byref(v1), C.byref(v2), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), h1.ctypes.data_as(C.c_void_p),
                             h2.ctypes.data_as(C.c_void_p), th, m.ctypes.data_as(C.c_void_p))
    return n, m[:K1.n]


def _bow_kf_call(fn, K1, K2, fv1, fv2, has1, has2, nn_ratio, check_ori):
    v1, v2 = K1.view(), K2.view()
    s1, s2 = featvec_struct(fv1), featvec_struct(fv2)
    h1 = np.ascontiguousarray(has1, np.uint8); h2 = np.ascontiguousarray(has2, np.uint8)
    m = np.full(max(K1.n, 1), -1, np.int32)
    fn.argtypes = [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_void_p]
    n = fn(C.byref(v1), C.byref(v2), C.byref(s1), C.byref(s2), h1.ctypes.data_as(C.c_void_p), h2.ctypes.data_as(C.c_void_p), nn_ratio, int(check_ori),
           m.ctypes.data_as(C.c_void_p))
    return n, m[:K1.n]


def search_by_bow_kf(K1, K2, fv1, fv2, has1, has2, nn_ratio=0.8, check_ori=True):
    return _bow_kf_call(_setup().orc_search_by_bow_kf, K1, K2, fv1, fv2, has1, has2, nn_ratio, check_ori)


def ref_search_by_bow_kf(K1, K2, fv1, fv2, has1, has2, nn_ratio=0.8, check_ori=True):
    return _bow_kf_call(_ref_lib().ref_search_by_bow_kf, K1, K2, fv1, fv2, has1, has2, nn_ratio, check_ori)


def _init_call(fn, F1, F2, prev_matched, window_size, nn_ratio, check_ori):
    v1, v2 = F1.view(), F2.view()
    prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    assert len(prev) == F1.n
    m = np.full(max(F1.n, 1), -1, np.int32)
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = fn(C.byref(v1), C.byref(v2), prev.ctypes.data_as(C.c_void_p), int(window_size), nn_ratio, int(check_ori), m.ctypes.data_as(C.c_void_p))
    return n, m[:F1.n], prev


def search_for_initialization(F1, F2, prev_matched, window_size=100, nn_ratio=0.9, check_ori=True):
    """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:732-852); Tracking::MonocularInitialization builds the matcher with (0.9, true) and
    calls it with windowSize 100 (src/Tracking.cc).  Returns (nmatches, vnMatches12, updated vbPrevMatched)."""
    return _init_call(_setup().orc_search_for_initialization, F1, F2, prev_matched, window_size, nn_ratio, check_ori)


def ref_search_for_initialization(F1, F2, prev_matched, window_size=100, nn_ratio=0.9, check_ori=True):
    return _init_call(_ref_lib().ref_search_for_initialization, F1, F2, prev_matched, window_size, nn_ratio, check_ori)


def distinctive_descriptor(desc):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:428-455) for one map point, in numpy: pairwise Hamming distances,
    median = sorted(row)[int(0.5*(N-1))], first descriptor with the smallest median.  Pinned: tests/test_oracle_vs_reference_match.py
    compares it with the reference's own function body (sliced out of MapPoint.cc at build time)."""
    d = np.asarray(desc, np.uint8).reshape(-1, 32)
    n = len(d)
    if n == 0:
        return -1
    D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2).astype(np.int64)
    med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
    return int(np.argmin(med))


# ---- Frame::ComputeStereoMatches through the reference's own text (oracle/_ref/libstereo_ref.so) -------------------------------------
_stereo = None


def stereo_ref_available():
    from . import ref_build
    return ref_build.build_stereo() is not None


def ref_compute_stereo_matches(left, right, nfeatures, mb, mbf, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """runs the reference's ORBextractor.cc on both rectified images and the reference's Frame::ComputeStereoMatches body on the result.
    Returns (left keypoints [n,7] float32, mvuRight, mvDepth, number of right keypoints)."""
    global _stereo
    if _stereo is None:
        from . import ref_build
        so = ref_build.build_stereo()
        if so is None:
            raise RuntimeError("oracle/_ref/libstereo_ref.so missing and /root/reference not present")
        lib()
        _stereo = C.CDLL(so)
        _stereo.ref_compute_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                                       C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
    cap = 4 * nfeatures
    keys = np.zeros((cap, 7), np.float32); ur = np.zeros(cap, np.float32); dp = np.zeros(cap, np.float32)
    nr = C.c_int()
    n = _stereo.ref_compute_stereo_matches(left.ctypes.data_as(C.c_void_p), right.ctypes.data_as(C.c_void_p), left.shape[1], left.shape[0], left.strides[0],
                                           nfeatures, scale_factor, nlevels, ini_th, min_th, mb, mbf, keys.ctypes.data_as(C.c_void_p), cap,
                                           ur.ctypes.data_as(C.c_void_p), dp.ctypes.data_as(C.c_void_p), C.byref(nr))
    assert n >= 0
    return keys[:n], ur[:n], dp[:n], nr.value


def ref_distinctive_descriptor(desc):
    """the reference's own MapPoint::ComputeDistinctiveDescriptors on an [n,32] uint8 array -> the chosen descriptor (32 bytes)"""
    l = _ref_lib()
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    out = np.zeros(32, np.uint8)
    l.ref_distinctive_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    l.ref_distinctive_descriptor(d.ctypes.data_as(C.c_void_p), len(d), out.ctypes.data_as(C.c_void_p))
    return out


def ref_stereo_from_rgbd(keys, depth, bf):
    """the reference's own Frame::ComputeStereoFromRGBD: keys = KP_DTYPE array, depth [h,w] float32 -> (mvuRight, mvDepth)"""
    l = _ref_lib()
    k = np.zeros((len(keys), 7), np.float32)
    k[:, 0] = keys["x"]; k[:, 1] = keys["y"]; k[:, 2] = keys["size"]
    d = np.ascontiguousarray(depth, np.float32)
    ur = np.zeros(max(len(keys), 1), np.float32); dz = np.zeros(max(len(keys), 1), np.float32)
    l.ref_stereo_from_rgbd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    l.ref_stereo_from_rgbd(k.ctypes.data_as(C.c_void_p), len(keys), d.ctypes.data_as(C.c_void_p), d.shape[1], d.shape[0], bf, ur.ctypes.data_as(C.c_void_p),
                           dz.ctypes.data_as(C.c_void_p))
    return ur[:len(keys)], dz[:len(keys)]


# ---- Frame::isInFrustum + MapPoint::PredictScale + Pinhole::project (the query builder of SearchLocalPoints) --------------------------
MAP_POINT = np.dtype([("xw", "f4", 3), ("normal", "f4", 3), ("min_dist", "f4"), ("max_dist", "f4"), ("flags", "u4"), ("desc", "u1", 32)])
assert MAP_POINT.itemsize == 68


class FrustumC(C.Structure):
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("bf", C.c_float), ("viewing_cos_limit", C.c_float), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float)]


def make_frustum(Twc, K, bounds, bf, viewing_cos_limit=0.5, scale_factor=1.2, nlevels=8):
    """Twc [3,4] float32 -> the members Frame::UpdatePoseMatrices keeps (mRcw = Rwc^T, mtcw = -Rcw*twc, mOw = twc), computed in float64 and rounded"""
    T = np.asarray(Twc, np.float64).reshape(3, 4)
    Rcw = T[:, :3].T; tcw = -Rcw @ T[:, 3]
    f = FrustumC()
    f.Rcw[:] = [float(np.float32(x)) for x in Rcw.reshape(9)]; f.tcw[:] = [float(np.float32(x)) for x in tcw]; f.Ow[:] = [float(np.float32(x)) for x in T[:, 3]]
    f.fx, f.fy, f.cx, f.cy = K["fx"], K["fy"], K["cx"], K["cy"]
    f.bf, f.viewing_cos_limit, f.scale_factor, f.nlevels = bf, viewing_cos_limit, scale_factor, nlevels
    f.min_x, f.min_y, f.max_x, f.max_y = bounds
    return f


def _frustum_call(fn, fr, pts):
    p = np.ascontiguousarray(pts, MAP_POINT)
    q = np.zeros(max(len(p), 1), MP_QUERY); iv = np.zeros(max(len(p), 1), np.uint8)
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    n = fn(C.byref(fr), p.ctypes.data_as(C.c_void_p), len(p), q.ctypes.data_as(C.c_void_p), iv.ctypes.data_as(C.c_void_p))
    return n, q[:len(p)], iv[:len(p)]


def in_frustum(fr, pts):
    return _frustum_call(_setup().orc_in_frustum, fr, pts)


_frustum = None


def frustum_ref_available():
    from . import ref_build
    return ref_build.build_frustum() is not None


def ref_in_frustum(fr, pts):
    global _frustum
    if _frustum is None:
        from . import ref_build
        _frustum = C.CDLL(ref_build.build_frustum())
    return _frustum_call(_frustum.ref_in_frustum, fr, pts)


def scale_thresholds(scale_factor, nlevels):
    l = _setup()
    T = np.zeros(max(nlevels - 1, 1), np.float32)
    l.orc_scale_thresholds.argtypes = [C.c_float, C.c_int, C.c_void_p]
    l.orc_scale_thresholds(scale_factor, nlevels, T.ctypes.data_as(C.c_void_p))
    return T[:nlevels - 1]
