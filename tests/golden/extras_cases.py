"""Inputs of tests/golden/extras_ref.npz (shared by the generator and the tests): seeded synthetic data only."""
import numpy as np

from plvs_b200 import synth, scenario, tsdf as T
from oracle import orb as O

INIT_PARAMS = [(100, 0.9, True), (30, 0.7, True), (100, 0.9, False)]
MESH_CASES = [("scan_colour", dict(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=1), (0, 1, 2, 6), True),
              ("scan_plain", dict(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=0), (0, 1), False),
              ("coarse_colour", dict(voxel_resolution=0.5, use_carving=1, near_plane=0.1, far_plane=6.0, max_blocks=4096, use_color=1), (0, 3), True)]
BOW_CASES = [(10, 3, 2, 0, 0, 0.0), (6, 4, 4, 1, 1, 0.1), (5, 3, 0, 2, 3, 0.2)]


def init_frames():
    K = synth.intrinsics(640, 480)
    tab = O.Tables(2000)
    out = []
    for f in (10, 11, 15):
        kp, desc, _, _ = O.extract_port(synth.gray_frame(f), 2000)
        out.append(scenario.make_frame(kp, desc, synth.depth_frame(f), K, tab.scale))
    return out


def mesh_map(cls, kw, frames, color, **ctor):
    """a map of class `cls` (oracle, compiled reference or the product's ChiselServer) after the case's scans"""
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    p = T.default_params(**kw)
    m = cls(p, **ctor)
    (m.SetDepthCameraInfo if hasattr(m, "SetDepthCameraInfo") else m.set_camera)(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    for f in frames:
        m.integrate(synth.depth_frame(f, w, h), synth.pose(f), synth.bgr_frame(f, w, h) if color else None)
    return m


def bow_descriptors():
    return np.concatenate([O.extract_port(synth.gray_frame(f), 1500)[1] for f in (3, 4)])
