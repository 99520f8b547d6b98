"""Host-side mirror of chisel_server::ChiselServer's depth-scan integration as PLVS drives it through
PointCloudMapChisel (reference: Thirdparty/chisel_server/include/chisel_server/ChiselServer.h:81-322,
src/PointCloudMapChisel.cc:46-225), on top of the C ABI of libplvs_b200.so."""
import ctypes as C
import numpy as np

from . import _lib

SCAN, SCAN_COLOR = 0, 1


def default_params(**kw):
    """ChiselServerParams() defaults (ChiselServer.cpp:44-69) overridden the way PointCloudMapChisel does."""
    lib = _lib.load()
    p = _lib.TsdfParams()
    lib.plvs_tsdf_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class ChiselServer:
    def __init__(self, params=None, device=0, **kw):
        self._lib = _lib.load()
        self.params = params or default_params(**kw)
        self._h = C.c_void_p()
        _lib.check(self._lib.plvs_tsdf_create(C.byref(self.params), device, C.byref(self._h)), "plvs_tsdf_create")
        self._pose = None
        self._depth = None
        self._color = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.plvs_tsdf_destroy(self._h)
            self._h = None

    __del__ = close

    def Reset(self):
        _lib.check(self._lib.plvs_tsdf_reset(self._h), "plvs_tsdf_reset")

    def SetDepthCameraInfo(self, fx, fy, cx, cy, width, height):
        _lib.check(self._lib.plvs_tsdf_set_camera(self._h, C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy), width, height),
                   "plvs_tsdf_set_camera")

    SetColorCameraInfo = SetDepthCameraInfo      # one camera model (IntegrateDepthScanColorWithOneCameraModelBGR)

    def SetDepthPose(self, Twc):
        self._pose = np.ascontiguousarray(Twc, np.float32).reshape(12)

    SetColorPose = SetDepthPose

    def SetDepthImageMemorySharing(self, depth):
        self._depth = np.ascontiguousarray(depth, np.float32)

    def SetColorImageMemorySharing(self, bgr):
        self._color = np.ascontiguousarray(bgr, np.uint8)

    def IntegrateLastDepthImage(self, updateMesh=False):
        if self._pose is None or self._depth is None:
            raise _lib.PlvsError("ChiselServer - PROBLEM in integrating depth scan (no pose / depth)")
        use_color = bool(self.params.use_color) and self._color is not None
        h, w = self._depth.shape
        bgr = self._color if use_color else None
        rc = self._lib.plvs_tsdf_integrate_depth(self._h, self._depth.ctypes.data_as(C.c_void_p), w, h,
                                                 bgr.ctypes.data_as(C.c_void_p) if use_color else None,
                                                 bgr.strides[0] if use_color else 0, bgr.shape[2] if use_color else 0,
                                                 self._pose.ctypes.data_as(C.c_void_p), SCAN_COLOR if use_color else SCAN, 0)
        _lib.check(rc, "plvs_tsdf_integrate_depth")

    def integrate(self, depth, Twc, bgr=None):
        self.SetDepthPose(Twc)
        self.SetDepthImageMemorySharing(depth)
        self._color = None if bgr is None else np.ascontiguousarray(bgr, np.uint8)
        self.IntegrateLastDepthImage(False)

    def integrate_u16(self, depth_u16, depth_factor, Twc, bgr=None):
        """IntegrateLastDepthImage on a raw 16-bit depth map: `convertTo(CV_32F, mDepthMapFactor)` (src/Tracking.cc:1812-1813) runs on the device."""
        d = np.asarray(depth_u16)
        if d.dtype != np.uint16 or d.ndim != 2 or d.strides[1] != 2 or d.strides[0] < 2 * d.shape[1]:      # padded rows are passed through as they are
            d = np.ascontiguousarray(d, np.uint16)
        T = np.ascontiguousarray(Twc, np.float32).reshape(12)
        c = None if bgr is None else np.ascontiguousarray(bgr, np.uint8)
        rc = self._lib.plvs_tsdf_integrate_depth_u16(self._h, d.ctypes.data_as(C.c_void_p), d.shape[1], d.shape[0], d.strides[0], depth_factor,
                                                     c.ctypes.data_as(C.c_void_p) if c is not None else None, c.strides[0] if c is not None else 0, c.shape[2] if c is not None else 0,
                                                     T.ctypes.data_as(C.c_void_p), SCAN_COLOR if c is not None else SCAN)
        _lib.check(rc, "plvs_tsdf_integrate_depth_u16")

    def integrate_cloud(self, xyz, rgb, Twc, depth=None):
        """SetPointCloud + (SetDepthImageMemorySharing) + IntegrateLastPointCloud(false): Chisel::IntegratePointCloudWidthDepth."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        T = np.ascontiguousarray(Twc, np.float32).reshape(12)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.float32)
        d = None if depth is None else np.ascontiguousarray(depth, np.float32)
        rc = self._lib.plvs_tsdf_integrate_cloud(self._h, xyz.ctypes.data_as(C.c_void_p), rgb.ctypes.data_as(C.c_void_p) if rgb is not None else None,
                                                 len(xyz), d.ctypes.data_as(C.c_void_p) if d is not None else None,
                                                 d.shape[1] if d is not None else 0, d.shape[0] if d is not None else 0, T.ctypes.data_as(C.c_void_p))
        _lib.check(rc, "plvs_tsdf_integrate_cloud")

    def integrate_cloud_kf(self, xyz, rgb, Twc, depth=None, kfids=None, kfid=0):
        """integrate_cloud with the cloud's keyframe ids (PointCloud::GetKfids): kfids [n] uint32, or one id for every point"""
        xyz = np.ascontiguousarray(xyz, np.float32)
        T = np.ascontiguousarray(Twc, np.float32).reshape(12)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.float32)
        d = None if depth is None else np.ascontiguousarray(depth, np.float32)
        kf = None if kfids is None else np.ascontiguousarray(kfids, np.uint32)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        rc = self._lib.plvs_tsdf_integrate_cloud_kf(self._h, p(xyz), p(rgb), p(kf), int(kfid), len(xyz), p(d), d.shape[1] if d is not None else 0,
                                                    d.shape[0] if d is not None else 0, p(T))
        _lib.check(rc, "plvs_tsdf_integrate_cloud_kf")

    def Deform(self, kfids, Rt, chunk_order=None):
        """ChiselServer::Deform(MapKfidRt&): kfids [n] uint32 -> Rt [n,3,4] (R | t).  chunk_order: optional chunk ids [m,3] to visit first (the
        reference walks a std::unordered_map; the default, key order, is deterministic)"""
        kf = np.ascontiguousarray(kfids, np.uint32); T = np.ascontiguousarray(Rt, np.float32).reshape(len(kf), 12)
        o = None if chunk_order is None else np.ascontiguousarray(chunk_order, np.int32).reshape(-1, 3)
        rc = self._lib.plvs_tsdf_deform(self._h, kf.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p), len(kf),
                                        o.ctypes.data_as(C.c_void_p) if o is not None else None, 0 if o is None else len(o))
        _lib.check(rc, "plvs_tsdf_deform")

    def IntegrateWorldPointCloud(self, xyz, rgb, normals, Twc, kfids=None, kfid=0):
        """ChiselServer::IntegrateWorldPointCloud (what PointCloudMapChisel::LoadMap feeds a saved map into): points + normals in the frame of Twc,
        colours in [0,1] or None, per-point keyframe ids or one id"""
        xyz = np.ascontiguousarray(xyz, np.float32); nrm = np.ascontiguousarray(normals, np.float32)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.float32)
        kf = None if kfids is None else np.ascontiguousarray(kfids, np.uint32)
        T = np.ascontiguousarray(Twc, np.float32).reshape(12)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        rc = self._lib.plvs_tsdf_integrate_world_cloud(self._h, p(xyz), p(rgb), p(nrm), p(kf), int(kfid), len(xyz), p(T))
        _lib.check(rc, "plvs_tsdf_integrate_world_cloud")

    def download_kfid(self):
        """DistVoxel::GetKfid of every voxel, blocks sorted like download()"""
        n = C.c_int()
        _lib.check(self._lib.plvs_tsdf_download_kfid(self._h, None, 0, C.byref(n)), "plvs_tsdf_download_kfid")
        n = n.value
        keys = np.zeros((n, 3), np.int32); out = np.zeros((n, 4096), np.uint32); m = C.c_int()
        if n:
            _lib.check(self._lib.plvs_tsdf_download_blocks(self._h, keys.ctypes.data_as(C.c_void_p), None, None, None, n, C.byref(m)), "plvs_tsdf_download_blocks")
            _lib.check(self._lib.plvs_tsdf_download_kfid(self._h, out.ctypes.data_as(C.c_void_p), n, C.byref(m)), "plvs_tsdf_download_kfid")
        return out[np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))]

    def mesh_kfids(self):
        """Mesh::kfids of the meshes GetMeshes() returns, per vertex"""
        nm, nv = getattr(self, "_mesh_sizes", None) or self.UpdateMesh()
        out = np.zeros(max(nv, 1), np.uint32)
        _lib.check(self._lib.plvs_tsdf_get_mesh_kfids(self._h, out.ctypes.data_as(C.c_void_p), nv, 0), "plvs_tsdf_get_mesh_kfids")
        return out[:nv]

    def stats(self):
        s = _lib.TsdfStats()
        _lib.check(self._lib.plvs_tsdf_last_stats(self._h, C.byref(s)), "plvs_tsdf_last_stats")
        return {f: getattr(s, f) for f, _ in s._fields_}

    def UpdateMesh(self):
        """ChiselServer::UpdateMesh (Chisel::UpdateMeshes): marching cubes + colours + gradient normals for every chunk, left on the device.
        -> (number of non-empty chunk meshes, number of vertices)"""
        nm, nv = C.c_int(), C.c_longlong()
        _lib.check(self._lib.plvs_tsdf_update_meshes(self._h, C.byref(nm), C.byref(nv)), "plvs_tsdf_update_meshes")
        self._mesh_sizes = (nm.value, nv.value)
        return self._mesh_sizes

    def GetMeshes(self):
        """ChunkManager::GetAllMeshes of the last UpdateMesh, chunk-key order -> keys[m,3], counts[m], vertices[v,3], normals[v,3], colours[v,3]"""
        nm, nv = getattr(self, "_mesh_sizes", None) or self.UpdateMesh()
        keys = np.zeros((nm, 3), np.int32); counts = np.zeros(nm, np.int32)
        V = np.zeros((nv, 3), np.float32); N = np.zeros((nv, 3), np.float32); Cc = np.zeros((nv, 3), np.float32)
        _lib.check(self._lib.plvs_tsdf_get_meshes(self._h, keys.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), nm, V.ctypes.data_as(C.c_void_p),
                                                  N.ctypes.data_as(C.c_void_p), Cc.ctypes.data_as(C.c_void_p), nv, 0), "plvs_tsdf_get_meshes")
        return keys, counts, V, N, Cc

    def GetPointCloud(self):
        """ChiselServer::GetPointCloud (ChiselServer.cpp:872-1075) for a coloured map: one point per mesh vertex, r/g/b = colour * 255 truncated
        to a byte, plus the normals.  Chunk order is (x,y,z) key order (the reference walks an unordered_map)."""
        _, _, V, N, Cc = self.GetMeshes()
        return V, (Cc * np.float32(255)).astype(np.uint8), N

    def SaveMesh(self, filename):
        """ChiselServer::SaveMesh -> Chisel::SaveAllMeshesToPLY: the meshes of the last UpdateMesh as the reference's ASCII PLY (chunks in key order)"""
        _, _, V, _, Cc = self.GetMeshes()
        return save_ply(filename, V, Cc if self.params.use_color else None)

    def download(self):
        """-> keys[n,3] (sorted lexicographically), sdf[n,4096], weight[n,4096], rgba[n,4096,4]"""
        n = C.c_int()
        _lib.check(self._lib.plvs_tsdf_download_blocks(self._h, None, None, None, None, 0, C.byref(n)), "plvs_tsdf_download_blocks")
        n = n.value
        keys = np.zeros((n, 3), np.int32); sdf = np.zeros((n, 4096), np.float32); w = np.zeros((n, 4096), np.float32)
        rgba = np.zeros((n, 4096, 4), np.uint8)
        m = C.c_int()
        if n:
            _lib.check(self._lib.plvs_tsdf_download_blocks(self._h, keys.ctypes.data_as(C.c_void_p), sdf.ctypes.data_as(C.c_void_p),
                                                           w.ctypes.data_as(C.c_void_p), rgba.ctypes.data_as(C.c_void_p), n, C.byref(m)),
                       "plvs_tsdf_download_blocks")
        order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
        return keys[order], sdf[order], w[order], rgba[order]


def save_ply(filename, verts, colors=None):
    """SaveMeshPLYASCII (Thirdparty/open_chisel/src/io/PLY.cpp:29-86) for concatenated mesh vertices; host I/O through the C ABI"""
    V = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    Cc = None if colors is None else np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
    rc = _lib.load().plvs_mesh_save_ply(str(filename).encode(), V.ctypes.data_as(C.c_void_p), Cc.ctypes.data_as(C.c_void_p) if Cc is not None else None, len(V))
    _lib.check(rc, "plvs_mesh_save_ply")
    return True


def save_map_ply(filename, xyz, bgra, normals, label, kfid, is_mesh=True, binary=True):
    """PointCloudMap<PointT>::WritePLY (src/PointCloudMap.cc:325-437): the map file PointCloudMapChisel::SaveMap writes.  bgra: n x 4 bytes in PCL's order."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3); n = len(xyz)
    bgra = np.ascontiguousarray(bgra, np.uint8).reshape(n, 4); normals = np.ascontiguousarray(normals, np.float32).reshape(n, 3)
    label = np.ascontiguousarray(label, np.uint32).reshape(n); kfid = np.ascontiguousarray(kfid, np.uint32).reshape(n)
    _lib.check(_lib.load().plvs_map_save_ply(str(filename).encode(), xyz.ctypes.data, bgra.ctypes.data, normals.ctypes.data, label.ctypes.data, kfid.ctypes.data, n,
                                             int(is_mesh), int(binary)), "plvs_map_save_ply")


def load_map_ply(filename):
    """the reading half of PointCloudMap<PointT>::LoadMap: -> dict(xyz, rgb (as named in the file), normals, label, kfid, fields)"""
    lib = _lib.load()
    n, fields = C.c_longlong(), C.c_int()
    rc = lib.plvs_map_load_ply(str(filename).encode(), None, None, None, None, None, 0, C.byref(n), C.byref(fields))
    if rc not in (0, -4):
        _lib.check(rc, "plvs_map_load_ply")
    m = max(int(n.value), 1)
    out = dict(xyz=np.zeros((m, 3), np.float32), rgb=np.zeros((m, 3), np.uint8), normals=np.zeros((m, 3), np.float32), label=np.zeros(m, np.uint32), kfid=np.zeros(m, np.uint32))
    _lib.check(lib.plvs_map_load_ply(str(filename).encode(), out["xyz"].ctypes.data, out["rgb"].ctypes.data, out["normals"].ctypes.data, out["label"].ctypes.data,
                                     out["kfid"].ctypes.data, m, C.byref(n), C.byref(fields)), "plvs_map_load_ply")
    out = {k: v[:n.value] for k, v in out.items()}
    out["fields"] = fields.value
    return out
