"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/tsdf_oracle.cpp (brute-force CPU restatement of
Chisel's depth-scan integration).  Same parameter struct as the product (plvs_b200._lib.TsdfParams)."""
import ctypes as C
import numpy as np

from .orb import lib
from plvs_b200 import _lib as _abi


def default_params(**kw):
    """ChiselServerParams() defaults (Thirdparty/chisel_server/src/ChiselServer.cpp:44-69) in the flat parameter record, overridden by keywords --
    built here, on the oracle side, so the CPU arm of bench.py needs nothing from libplvs_b200.so (the product's copy: plvs_tsdf_default_params)"""
    p = _abi.TsdfParams()
    p.voxel_resolution = 0.015
    p.trunc_quad, p.trunc_linear, p.trunc_const, p.trunc_scale = 0.0019, -0.00152, 0.001504, 6.0
    p.weight, p.use_carving, p.carving_dist, p.use_color = 1.0, 1, 0.05, 1
    p.near_plane, p.far_plane, p.max_blocks = 0.05, 5.0, 65536
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class _Prefixed:
    """view of a ctypes library that prepends a prefix to every symbol (orc_ = restatement, ref_ = compiled reference)"""
    def __init__(self, l, prefix):
        object.__setattr__(self, "_l", l); object.__setattr__(self, "_p", prefix)

    def __getattr__(self, name):
        assert name.startswith("orc_")
        return getattr(self._l, self._p + name[4:])


class Map:
    def _library(self):
        return lib()

    def __init__(self, params, threads=1):
        self._l = self._library()
        self._l.orc_tsdf_create.restype = C.c_void_p
        self._l.orc_tsdf_create.argtypes = [C.c_void_p, C.c_int]
        self._l.orc_tsdf_destroy.argtypes = [C.c_void_p]
        self._l.orc_tsdf_set_camera.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        self._l.orc_tsdf_integrate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self._l.orc_tsdf_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._l.orc_tsdf_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        self._l.orc_tsdf_reset.argtypes = [C.c_void_p]
        self.params = params
        self._h = self._l.orc_tsdf_create(C.byref(params), threads)

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.orc_tsdf_destroy(self._h)
            self._h = None

    def set_camera(self, fx, fy, cx, cy, w, h):
        self._l.orc_tsdf_set_camera(self._h, fx, fy, cx, cy, w, h)

    def integrate(self, depth, Twc, bgr=None):
        depth = np.ascontiguousarray(depth, np.float32)
        T = np.ascontiguousarray(Twc, np.float32).reshape(12)
        h, w = depth.shape
        mode = 1 if bgr is not None else 0
        if bgr is not None:
            bgr = np.ascontiguousarray(bgr, np.uint8)
        rc = self._l.orc_tsdf_integrate(self._h, depth.ctypes.data_as(C.c_void_p), w, h,
                                        bgr.ctypes.data_as(C.c_void_p) if bgr is not None else None,
                                        bgr.shape[2] if bgr is not None else 0, T.ctypes.data_as(C.c_void_p), mode)
        assert rc == 0, rc

    def integrate_cloud(self, xyz, rgb, Twc, depth=None):
        """Chisel::IntegratePointCloudWidthDepth: xyz [n,3] camera-frame points, rgb [n,3] floats in [0,1] (or None)."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        T = np.ascontiguousarray(Twc, np.float32).reshape(12)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.float32)
        d = None if depth is None else np.ascontiguousarray(depth, np.float32)
        self._l.orc_tsdf_integrate_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        rc = self._l.orc_tsdf_integrate_cloud(self._h, xyz.ctypes.data_as(C.c_void_p), rgb.ctypes.data_as(C.c_void_p) if rgb is not None else None,
                                              len(xyz), d.ctypes.data_as(C.c_void_p) if d is not None else None,
                                              d.shape[1] if d is not None else 0, d.shape[0] if d is not None else 0, T.ctypes.data_as(C.c_void_p))
        assert rc == 0

    def stats(self):
        s = np.zeros(5, np.int32)
        self._l.orc_tsdf_stats(self._h, s.ctypes.data_as(C.c_void_p))
        return dict(n_blocks=int(s[0]), n_range=int(s[1]), n_updated=int(s[2]), n_new=int(s[3]), n_collected=int(s[4]))

    def download(self):
        n = self._l.orc_tsdf_download(self._h, None, None, None, None, 0)
        keys = np.zeros((n, 3), np.int32); sdf = np.zeros((n, 4096), np.float32); w = np.zeros((n, 4096), np.float32)
        rgba = np.zeros((n, 4096, 4), np.uint8)
        if n:
            self._l.orc_tsdf_download(self._h, keys.ctypes.data_as(C.c_void_p), sdf.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                      rgba.ctypes.data_as(C.c_void_p), n)
        return keys, sdf, w, rgba


def _mesh_call(fn, h):
    """(keys [m,3], counts [m], verts [v,3], normals [v,3], colors [v,3]) of the non-empty chunk meshes in key order"""
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    tot = C.c_long()
    nm = fn(h, None, None, 0, None, None, None, 0, C.byref(tot))
    nv = tot.value
    keys = np.zeros((nm, 3), np.int32); counts = np.zeros(nm, np.int32)
    V = np.zeros((nv, 3), np.float32); N = np.zeros((nv, 3), np.float32); Cc = np.zeros((nv, 3), np.float32)
    if nm:
        fn(h, keys.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), nm, V.ctypes.data_as(C.c_void_p), N.ctypes.data_as(C.c_void_p),
           Cc.ctypes.data_as(C.c_void_p), nv, C.byref(tot))
    return keys, counts, V, N, Cc


def _integrate_cloud_kf(self, xyz, rgb, Twc, depth=None, kfids=None, kfid=0):
    """Chisel::IntegratePointCloudWidthDepth with the cloud's keyframe ids: kfids [n] uint32, or one id for every point"""
    xyz = np.ascontiguousarray(xyz, np.float32)
    T = np.ascontiguousarray(Twc, np.float32).reshape(12)
    rgb = None if rgb is None else np.ascontiguousarray(rgb, np.float32)
    d = None if depth is None else np.ascontiguousarray(depth, np.float32)
    kf = None if kfids is None else np.ascontiguousarray(kfids, np.uint32)
    fn = self._l.orc_tsdf_integrate_cloud_kf
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = fn(self._h, p(xyz), p(rgb), p(kf), int(kfid), len(xyz), p(d), d.shape[1] if d is not None else 0, d.shape[0] if d is not None else 0, p(T))
    assert rc == 0


def _download_kfid(self):
    n = self._l.orc_tsdf_download(self._h, None, None, None, None, 0)
    out = np.zeros((n, 4096), np.uint32)
    self._l.orc_tsdf_download_kfid.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    if n:
        self._l.orc_tsdf_download_kfid(self._h, out.ctypes.data_as(C.c_void_p), n)
    return out


def _mesh_kfids(self, n_verts):
    """Mesh::kfids of the meshes extract_mesh() / meshes() return, per vertex"""
    out = np.zeros(max(n_verts, 1), np.uint32)
    self._l.orc_tsdf_extract_mesh_kfids.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    self._l.orc_tsdf_extract_mesh_kfids(self._h, out.ctypes.data_as(C.c_void_p), n_verts)
    return out[:n_verts]


def _deform(self, kfids, Rt, order=None):
    """ChunkManager::Deform: kfids [n] uint32, Rt [n,3,4] float32 (R | t).  `order`: chunk keys [m,3] in the visiting order (the restatement
    only; the compiled reference walks its own unordered_map)"""
    kf = np.ascontiguousarray(kfids, np.uint32); T = np.ascontiguousarray(Rt, np.float32).reshape(len(kf), 12)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    if isinstance(self, RefMap):
        fn = self._l.orc_tsdf_deform
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        with _quiet():
            rc = fn(self._h, p(kf), p(T), len(kf))
    else:
        o = None if order is None else np.ascontiguousarray(order, np.int32).reshape(-1, 3)
        fn = self._l.orc_tsdf_deform
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        rc = fn(self._h, p(kf), p(T), len(kf), p(o), 0 if o is None else len(o))
    assert rc == 0


def _integrate_world_cloud(self, xyz, rgb, normals, Twc, kfids=None, kfid=0):
    """Chisel::IntegrateWorldPointCloudWithNormals: points (cloud frame), colours in [0,1] or None, normals, Twc; per-point keyframe ids or one id"""
    xyz = np.ascontiguousarray(xyz, np.float32); nrm = np.ascontiguousarray(normals, np.float32)
    rgb = None if rgb is None else np.ascontiguousarray(rgb, np.float32)
    kf = None if kfids is None else np.ascontiguousarray(kfids, np.uint32)
    T = np.ascontiguousarray(Twc, np.float32).reshape(12)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    fn = self._l.orc_tsdf_integrate_world_cloud
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
    with _quiet():
        rc = fn(self._h, p(xyz), p(rgb), p(nrm), p(kf), int(kfid), len(xyz), p(T))
    assert rc == 0


import contextlib, os, sys


@contextlib.contextmanager
def _quiet():
    """the reference prints progress lines to stdout"""
    sys.stdout.flush()
    saved = os.dup(1); null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1); os.close(null); os.close(saved)


Map.deform = _deform
Map.integrate_world_cloud = _integrate_world_cloud
Map.integrate_cloud_kf = _integrate_cloud_kf
Map.download_kfid = _download_kfid
Map.mesh_kfids = _mesh_kfids


def _extract_mesh(self):
    """ChunkManager::RecomputeMesh for every chunk of the current map (GenerateMesh + ColorizeMesh + ComputeNormalsFromGradients)."""
    return _mesh_call(self._l.orc_tsdf_extract_mesh, self._h)


Map.extract_mesh = _extract_mesh


def depth_u16_to_f32(d16, factor):
    """`mImDepth.convertTo(mImDepth, CV_32F, mDepthMapFactor)` (src/Tracking.cc:1812-1813): OpenCV's scaled 16u -> 32f conversion is
    `(float)src * (float)alpha` with one rounding (beta = 0).  Pinned against cv2's scaled 16u->32f arithmetic in tests/test_oracle_match_tsdf.py
    (every u16 value); cv::Mat::convertTo itself has no Python binding, so this row is otherwise "parity unpinned"."""
    return np.asarray(d16, np.uint16).astype(np.float32) * np.float32(factor)


def ref_available():
    """oracle/_ref/libchisel_ref.so = the reference's own open_chisel sources (oracle/ref_build.py)."""
    from . import ref_build
    return ref_build.build() is not None


class RefMap(Map):
    """Same interface, backed by the REFERENCE's open_chisel compiled from /root/reference (against the Eigen stand-in).
    stats() only fills n_blocks and n_range (the reference does not count updates)."""
    _lib = None

    def _library(self):
        if RefMap._lib is None:
            from . import ref_build
            so = ref_build.build()
            if so is None:
                raise RuntimeError("oracle/_ref/libchisel_ref.so missing and /root/reference not present")
            RefMap._lib = C.CDLL(so)
        return _Prefixed(RefMap._lib, "ref_")

    def chunk_order(self):
        """keys [n,3] of the chunk map in its own iteration order (what ChunkManager::Deform walks)"""
        fn = RefMap._lib.ref_tsdf_chunk_order
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        n = fn(self._h, None, 0)
        keys = np.zeros((n, 3), np.int32)
        if n:
            fn(self._h, keys.ctypes.data_as(C.c_void_p), n)
        return keys

    def update_meshes(self, all_chunks=False):
        """Chisel::UpdateMeshes (the chunks flagged since the last call), or every chunk"""
        RefMap._lib.ref_tsdf_update_meshes.argtypes = [C.c_void_p, C.c_int]
        RefMap._lib.ref_tsdf_update_meshes(self._h, int(all_chunks))

    def meshes(self):
        """ChunkManager::GetAllMeshes as they are now (non-empty ones, key order)"""
        return _mesh_call(RefMap._lib.ref_tsdf_mesh_download, self._h)

    def extract_mesh(self):
        self.update_meshes(True)
        return self.meshes()

    def save_ply(self, path):
        """Chisel::SaveAllMeshesToPLY of the current meshes"""
        RefMap._lib.ref_tsdf_save_ply.argtypes = [C.c_void_p, C.c_char_p]
        RefMap._lib.ref_tsdf_save_ply(self._h, str(path).encode())
