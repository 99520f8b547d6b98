"""ctypes binding of libplvs_b200.so (the C ABI of include/plvs_b200.h).

The library is built in-tree by plvs_b200/csrc/build.py.  There is no CPU fallback: if the
shared object is missing or no CUDA device is present the product fails loudly.
"""
import ctypes as C
import os
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("PLVS_B200_LIB", str(pathlib.Path(__file__).resolve().parent / "libplvs_b200.so")))


class PlvsError(RuntimeError):
    pass


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class OrbStats(C.Structure):
    _fields_ = [("pyramid_pixels", C.c_int64), ("candidates", C.c_int64), ("keypoints", C.c_int64),
                ("kernel_launches", C.c_int32), ("host_wait_candidates_ms", C.c_float), ("host_distribute_ms", C.c_float),
                ("host_wait_describe_ms", C.c_float), ("host_assemble_ms", C.c_float)]


class OrbDeviceView(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys", C.c_void_p), ("desc", C.c_void_p), ("cache_key", C.c_uint64),
                ("grid_cell_start", C.c_void_p), ("grid_sorted", C.c_void_p)]


MAX_LEVELS = 16


class FrameView(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys", C.c_void_p), ("desc", C.c_void_p), ("uright", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("grid_inv_w", C.c_float), ("grid_inv_h", C.c_float),
                ("scale_factors", C.c_float * MAX_LEVELS), ("level_sigma2", C.c_float * MAX_LEVELS),
                ("nlevels", C.c_int32), ("bf", C.c_float), ("on_device", C.c_int32), ("cache_key", C.c_uint64),
                ("grid_cell_start", C.c_void_p), ("grid_sorted", C.c_void_p)]


class PyramidView(C.Structure):
    _fields_ = [("data", C.c_void_p * MAX_LEVELS), ("w", C.c_int32 * MAX_LEVELS), ("h", C.c_int32 * MAX_LEVELS),
                ("pitch", C.c_int32 * MAX_LEVELS), ("nlevels", C.c_int32)]


class FeatVec(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_ids", C.c_void_p), ("offsets", C.c_void_p), ("features", C.c_void_p)]


class TsdfParams(C.Structure):
    _fields_ = [("voxel_resolution", C.c_float), ("trunc_quad", C.c_float), ("trunc_linear", C.c_float),
                ("trunc_const", C.c_float), ("trunc_scale", C.c_float), ("weight", C.c_float),
                ("use_carving", C.c_int32), ("carving_dist", C.c_float), ("use_color", C.c_int32),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("max_blocks", C.c_int32)]


class TsdfStats(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("n_range", C.c_int32), ("n_candidates", C.c_int32), ("n_updated", C.c_int32),
                ("n_new", C.c_int32), ("n_collected", C.c_int32), ("kernel_launches", C.c_int32), ("pool_exhausted", C.c_int32),
                ("total_updated", C.c_int64), ("total_candidates", C.c_int64), ("total_integrations", C.c_int64)]


class PipelineFrame(C.Structure):
    _fields_ = [("valid", C.c_int32), ("n_ql", C.c_int32), ("n_qm", C.c_int32), ("ql", C.c_void_p), ("qm", C.c_void_p), ("uright", C.c_void_p),
                ("fv_cur", FeatVec), ("fv_last", FeatVec), ("has_cur", C.c_void_p), ("has_last", C.c_void_p),
                ("F12", C.c_float * 9), ("ep", C.c_float * 2), ("last", FrameView)]


class PipelineJob(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("batch", C.c_int32), ("n_steps", C.c_int32),
                ("first_frame", C.c_int32), ("cap", C.c_int32), ("inputs_on_device", C.c_int32),
                ("gray", C.c_void_p), ("depth", C.c_void_p), ("bgr", C.c_void_p), ("poses", C.c_void_p), ("frames", C.c_void_p),
                ("view_template", FrameView), ("th_last", C.c_float), ("th_map", C.c_float), ("nnratio_map", C.c_float),
                ("flush_buf", C.c_void_p), ("flush_bytes", C.c_size_t)]


class PipelineStats(C.Structure):
    _fields_ = [("keypoints", C.c_int64), ("matches", C.c_int64), ("wall_s", C.c_double), ("busy_extract_s", C.c_double),
                ("busy_track_s", C.c_double), ("busy_tri_s", C.c_double), ("busy_map_s", C.c_double)]


_lib = None


def load():
    """Load the shared object (raises PlvsError with the build hint if it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PlvsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the product has no CPU fallback)")
    _lib = declare(C.CDLL(str(LIB_PATH)))
    return _lib


def declare(lib):
    """argument / result types of the C ABI (include/plvs_b200.h) on a loaded library object"""
    lib.plvs_version.restype = C.c_char_p
    lib.plvs_last_error.restype = C.c_char_p
    lib.plvs_orb_destroy.restype = None
    lib.plvs_orb_destroy.argtypes = [C.c_void_p]
    lib.plvs_orb_create.argtypes = [C.POINTER(OrbParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.plvs_orb_extract_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                           C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.plvs_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.plvs_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    lib.plvs_orb_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.plvs_orb_download_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.plvs_orb_device_result.argtypes = [C.c_void_p, C.c_int, C.POINTER(OrbDeviceView)]
    lib.plvs_orb_set_frame_grid.argtypes = [C.c_void_p, C.c_void_p]
    lib.plvs_match_last_phase_cycles.argtypes = [C.c_void_p, C.c_void_p]
    lib.plvs_line_knn2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.plvs_orb_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.plvs_orb_last_stats.argtypes = [C.c_void_p, C.POINTER(OrbStats)]
    lib.plvs_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    lib.plvs_host_free.argtypes = [C.c_void_p]
    lib.plvs_hamming256.argtypes = [C.c_void_p, C.c_void_p]
    lib.plvs_match_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.plvs_match_projection_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                                              C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_match_projection_last.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_match_triangulation.argtypes = [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_match_projection_reloc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_orb_extract_batch_color.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.plvs_orb_stereo_from_rgbd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.plvs_match_projection_map_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_tsdf_integrate_depth_u16.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.plvs_orb_undistort.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.plvs_match_in_frustum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_match_fuse_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_match_projection_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_match_bow.argtypes = [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_match_initialization.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_match_bow_kf.argtypes = [C.c_void_p] * 7 + [C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_distinctive_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.plvs_match_fuse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_orb_pyramid_view.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(PyramidView)]
    lib.plvs_stereo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_tsdf_default_params.restype = None
    lib.plvs_tsdf_default_params.argtypes = [C.POINTER(TsdfParams)]
    lib.plvs_tsdf_create.argtypes = [C.POINTER(TsdfParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.plvs_tsdf_reset.argtypes = [C.c_void_p]
    lib.plvs_tsdf_set_camera.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.plvs_tsdf_integrate_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    lib.plvs_tsdf_integrate_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.plvs_tsdf_last_stats.argtypes = [C.c_void_p, C.POINTER(TsdfStats)]
    lib.plvs_voc_create.argtypes = [C.c_int] * 6 + [C.c_void_p] * 5
    lib.plvs_voc_load_text.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    lib.plvs_voc_destroy.argtypes = [C.c_void_p]; lib.plvs_voc_destroy.restype = None
    lib.plvs_voc_size.argtypes = [C.c_void_p]
    lib.plvs_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_int)] + [C.c_void_p] * 3 + \
        [C.POINTER(C.c_int), C.c_void_p]
    lib.plvs_bow_vector.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    lib.plvs_tsdf_integrate_cloud_kf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.plvs_tsdf_download_kfid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.plvs_tsdf_get_mesh_kfids.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]
    lib.plvs_mesh_save_ply.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_longlong]
    lib.plvs_map_save_ply.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int]
    lib.plvs_map_load_ply.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    lib.plvs_tsdf_update_meshes.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
    lib.plvs_tsdf_get_meshes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]
    lib.plvs_tsdf_download_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.plvs_tsdf_export_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.plvs_tsdf_merge_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.plvs_tsdf_export_packed_rgba.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.plvs_tsdf_merge_packed_rgba.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.plvs_tsdf_deform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.plvs_tsdf_integrate_world_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
    lib.plvs_enable_peer_access.argtypes = [C.c_int, C.c_int]
    lib.plvs_io_bytes.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]
    lib.plvs_pipeline_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for name in ("plvs_match_destroy", "plvs_tsdf_destroy"):
        if hasattr(lib, name):
            getattr(lib, name).restype = None
            getattr(lib, name).argtypes = [C.c_void_p]
    return lib


def check(rc, what=""):
    if rc != 0:
        raise PlvsError(f"{what} failed with code {rc}: {load().plvs_last_error().decode()}")


def exported_symbols():
    """Names declared in include/plvs_b200.h (used by the CPU-side ABI test)."""
    import re
    hdr = (_HERE.parent / "include" / "plvs_b200.h").read_text()
    return sorted(set(re.findall(r"\b(plvs_[a-z0-9_]+)\s*\(", hdr)))
