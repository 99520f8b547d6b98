"""Caller-side data for the matcher entry points, built from two extracted frames of the synthetic
stream (SURVEY.md §8d "Matcher queries"): what Tracking/LocalMapping would hand to ORBmatcher.

frame k-1 keypoints with valid depth are back-projected to 3-D and act as map points (descriptor =
their ORB descriptor, predicted level = their octave); they are projected into frame k with the
known pose.  All arithmetic is float32 numpy, done ONCE and fed identically to the CUDA path and to
the oracle, so parity of the searches does not depend on it."""
import numpy as np

from .matcher import MP_QUERY, LAST_QUERY, Q_OBS_POSITIVE, Frame
from . import synth


def uright_from_depth(keys, depth, bf):
    """Frame::ComputeStereoFromRGBD (src/Frame.cc:2251-2279) without distortion: uRight = x - bf/d."""
    u = keys["x"].astype(np.int32); v = keys["y"].astype(np.int32)
    h, w = depth.shape
    d = depth[np.clip(v, 0, h - 1), np.clip(u, 0, w - 1)]
    ur = np.full(len(keys), -1, np.float32)
    ok = d > 0
    ur[ok] = keys["x"][ok] - np.float32(bf) / d[ok]
    return ur, np.where(ok, d, np.float32(-1)).astype(np.float32)


def backproject(keys, z, K, Twc):
    x = (keys["x"] - np.float32(K["cx"])) * z / np.float32(K["fx"])
    y = (keys["y"] - np.float32(K["cy"])) * z / np.float32(K["fy"])
    Pc = np.stack([x, y, z], 1).astype(np.float32)
    R, t = Twc[:, :3], Twc[:, 3]
    return (Pc @ R.T + t).astype(np.float32)


def project(Pw, K, Twc):
    R, t = Twc[:, :3], Twc[:, 3]
    Pc = ((Pw - t) @ R).astype(np.float32)
    invz = (np.float32(1.0) / Pc[:, 2]).astype(np.float32)
    u = (np.float32(K["fx"]) * Pc[:, 0] * invz + np.float32(K["cx"])).astype(np.float32)
    v = (np.float32(K["fy"]) * Pc[:, 1] * invz + np.float32(K["cy"])).astype(np.float32)
    return u, v, invz, Pc


def make_frame(keys, desc, depth, K, scale_factors):
    ur, z = uright_from_depth(keys, depth, K["bf"])
    f = Frame(keys, desc, K["w"], K["h"], scale_factors, uright=ur, bf=K["bf"])
    f.depth_at_kp = z
    return f


def map_queries(last, cur, K, Twc_last, Twc_cur, nlevels=8, seed=0):
    """Queries for SearchByProjection(F, vpMapPoints): map points seen in `last`, in view in `cur`."""
    rng = np.random.default_rng(seed)
    ok = last.depth_at_kp > 0
    idx = np.nonzero(ok)[0]
    Pw = backproject(last.keys[idx], last.depth_at_kp[idx], K, Twc_last)
    u, v, invz, Pc = project(Pw, K, Twc_cur)
    inview = (Pc[:, 2] > 0) & (u >= cur.min_x) & (u <= cur.max_x) & (v >= cur.min_y) & (v <= cur.max_y)
    idx, u, v, invz, Pc = idx[inview], u[inview], v[inview], invz[inview], Pc[inview]
    q = np.zeros(len(idx), MP_QUERY)
    q["proj_x"], q["proj_y"] = u, v
    q["proj_xr"] = u - np.float32(K["bf"]) * invz
    q["track_depth"] = np.linalg.norm(Pc, axis=1).astype(np.float32)
    q["view_cos"] = np.where(rng.random(len(idx)) < 0.5, np.float32(0.9995), np.float32(0.95))
    q["level"] = np.clip(last.keys["octave"][idx] + rng.integers(-1, 2, len(idx)), 0, nlevels - 1)
    q["flags"] = Q_OBS_POSITIVE
    q["desc"] = last.desc[idx]
    return q, idx


def last_queries(last, cur, K, Twc_last, Twc_cur):
    """Queries for SearchByProjection(Cur, Last): every last-frame feature with a (depth-backed) map point."""
    ok = last.depth_at_kp > 0
    idx = np.nonzero(ok)[0]
    Pw = backproject(last.keys[idx], last.depth_at_kp[idx], K, Twc_last)
    u, v, invz, Pc = project(Pw, K, Twc_cur)
    q = np.zeros(len(idx), LAST_QUERY)
    q["u"], q["v"], q["invz"] = u, v, invz
    q["last_octave"] = last.keys["octave"][idx]
    q["angle"] = last.keys["angle"][idx]
    q["flags"] = Q_OBS_POSITIVE
    q["desc"] = last.desc[idx]
    return q, idx


def node_ids(desc, nodes=1024):
    """Stand-in for DBoW2 level-4 node ids (declared substitution)."""
    return ((desc[:, 0].astype(np.int64) ^ (desc[:, 1].astype(np.int64) << 8)) % nodes).astype(np.int64)


def fundamental(K, Twc1, Twc2):
    """F12 as Pinhole::epipolarConstrain builds it (src/CameraModels/Pinhole.cpp:127-131) and the epipole of
    camera 1 in image 2 (src/ORBmatcher.cc:1009-1011); float32."""
    R1, t1 = Twc1[:, :3].astype(np.float64), Twc1[:, 3].astype(np.float64)
    R2, t2 = Twc2[:, :3].astype(np.float64), Twc2[:, 3].astype(np.float64)
    R12 = R1.T @ R2
    t12 = R1.T @ (t2 - t1)
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    Km = np.array([[K["fx"], 0, K["cx"]], [0, K["fy"], K["cy"]], [0, 0, 1]], np.float64)
    F12 = np.linalg.inv(Km.T) @ tx @ R12 @ np.linalg.inv(Km)
    C2 = R2.T @ (t1 - t2)
    ep = np.array([K["fx"] * C2[0] / C2[2] + K["cx"], K["fy"] * C2[1] / C2[2] + K["cy"]])
    return F12.astype(np.float32), ep.astype(np.float32)


def cloud_from_depth(depth, bgr, K, step=2, min_depth=0.1, max_depth=5.0):
    """Caller-side point cloud in the camera frame, the way PointCloudMapping builds it for the Chisel back-end
    (src/PointCloudMapping.cc:838-846,929-1010: every `step`-th pixel with a valid depth; colours in [0,1], r,g,b)."""
    h, w = depth.shape
    v, u = np.mgrid[0:h:step, 0:w:step]
    z = depth[v, u].astype(np.float32)
    ok = (z > np.float32(min_depth)) & (z < np.float32(max_depth))
    u, v, z = u[ok].astype(np.float32), v[ok].astype(np.float32), z[ok]
    x = (u - np.float32(K["cx"])) * z / np.float32(K["fx"])
    y = (v - np.float32(K["cy"])) * z / np.float32(K["fy"])
    xyz = np.stack([x, y, z], 1).astype(np.float32)
    rgb = None
    if bgr is not None:
        c = bgr[v.astype(int), u.astype(int)].astype(np.float32) * (np.float32(1.0) / np.float32(255.0))     # byteToFloat, Conversions.h:78
        rgb = np.ascontiguousarray(c[:, ::-1])
    return xyz, rgb
