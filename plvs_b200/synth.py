"""Synthetic RGB-D stream generator (SURVEY.md §8d "Synthetic inputs").

numpy PCG64, seed = 1234 + 1000*stream + frame.  Gray image = sigma-1.5 blurred uniform noise (contrast x1.8) +
random filled rectangles/discs (FAST corners) warped by a slow per-frame drift;
depth = analytic scene (floor + back wall at 3.5 m + sphere r=0.6 m at 2 m) ray-cast
from pose k + Kinect-style noise, 2 % invalid (0) pixels; poses on a circle of radius
0.3 m with 0.5 deg/frame yaw.  Pure numpy so it needs neither cv2 nor a GPU.
"""
import math
import numpy as np

TUM1 = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, bf=40.0)


def intrinsics(w, h):
    sx, sy = w / 640.0, h / 480.0
    return dict(fx=TUM1["fx"] * sx, fy=TUM1["fy"] * sy, cx=TUM1["cx"] * sx, cy=TUM1["cy"] * sy,
                bf=TUM1["bf"] * sx, w=w, h=h)


def _box_blur_noise(rng, h, w):
    n = rng.integers(0, 256, size=(h + 8, w + 8)).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-4, 5) / 1.5) ** 2); k /= k.sum()
    n = np.apply_along_axis(lambda r: np.convolve(r, k, mode="valid"), 1, n)
    n = np.apply_along_axis(lambda c: np.convolve(c, k, mode="valid"), 0, n)
    return n  # h x w


_SCENE_CACHE = {}


def _scene(stream, w, h):
    key = (stream, w, h)
    if key not in _SCENE_CACHE:
        rng = np.random.default_rng(99 + stream)
        H, W = h + 64, w + 64
        base = (_box_blur_noise(rng, H, W) - 127.5) * 1.8 + 110.0   # sigma-1.5 texture, ~8k FAST-20 corners @VGA
        yy, xx = np.mgrid[0:H, 0:W]
        nshape = int(40 * (w * h) / (640 * 480))
        for _ in range(nshape):
            g = float(rng.integers(0, 256))
            if rng.random() < 0.6:
                x0, y0 = int(rng.integers(0, W - 8)), int(rng.integers(0, H - 8))
                ww, hh = int(rng.integers(8, 90)), int(rng.integers(8, 90))
                base[y0:y0 + hh, x0:x0 + ww] = g
            else:
                cx, cy, r = rng.integers(0, W), rng.integers(0, H), rng.integers(5, 40)
                base[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = g
        _SCENE_CACHE[key] = base
    return _SCENE_CACHE[key]


def gray_frame(frame, w=640, h=480, stream=0):
    """u8 gray image of frame `frame`: the static scene shifted by a slow drift + fresh noise."""
    rng = np.random.default_rng(1234 + 1000 * stream + frame)
    base = _scene(stream, w, h)
    dx = 32 + int(round(24 * math.sin(0.05 * frame)))
    dy = 32 + int(round(16 * math.cos(0.035 * frame)))
    img = base[dy:dy + h, dx:dx + w] + rng.normal(0, 2.0, size=(h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def pose(frame):
    """Twc (3x4 float32): circle radius 0.3 m in the x-z plane, yaw 0.5 deg/frame."""
    yaw = math.radians(0.5 * frame)
    c, s = math.cos(yaw), math.sin(yaw)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64)
    ang = 0.02 * frame
    t = np.array([0.3 * math.cos(ang) - 0.3, 0.0, 0.3 * math.sin(ang)], np.float64)
    return np.concatenate([R, t[:, None]], 1).astype(np.float32)


def depth_frame(frame, w=640, h=480, stream=0, noise=True):
    """float32 depth (metres) of the analytic scene seen from pose(frame)."""
    K = intrinsics(w, h)
    Twc = pose(frame).astype(np.float64)
    R, t = Twc[:, :3], Twc[:, 3]
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d = np.stack([(u - K["cx"]) / K["fx"], (v - K["cy"]) / K["fy"], np.ones_like(u)], -1)  # z = 1 rays (camera)
    dw = d @ R.T
    o = t
    best = np.full((h, w), np.inf)
    # back wall z = 3.5 (world), floor y = +1.2 (y points down)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (3.5 - o[2]) / dw[..., 2]; s[(s <= 0) | ~np.isfinite(s)] = np.inf; best = np.minimum(best, s)
        s = (1.2 - o[1]) / dw[..., 1]; s[(s <= 0) | ~np.isfinite(s)] = np.inf; best = np.minimum(best, s)
        # sphere centre (0.1,0.3,2.0) r 0.6
        cc = np.array([0.1, 0.3, 2.0]) - o
        a = (dw * dw).sum(-1); b = -2 * (dw @ cc); c0 = cc @ cc - 0.36
        disc = b * b - 4 * a * c0
        s = (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a); s[(disc < 0) | (s <= 0)] = np.inf
        best = np.minimum(best, s)
    z = best  # since camera rays have z=1, parameter s equals camera depth
    z[~np.isfinite(z)] = 0.0
    rng = np.random.default_rng(777 + 1234 + 1000 * stream + frame)
    if noise:
        sig = 0.0012 + 0.0019 * (z - 0.4) ** 2
        z = z + rng.normal(0, 1, size=z.shape) * sig * (z > 0)
    z[rng.random(size=z.shape) < 0.02] = 0.0
    return z.astype(np.float32)


def bgr_frame(frame, w=640, h=480, stream=0):
    g = gray_frame(frame, w, h, stream)
    return np.stack([g, np.roll(g, 3, 1), 255 - g], -1).copy()
