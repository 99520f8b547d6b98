"""Synthetic RGB-D stream generator (SURVEY.md §8d "Synthetic inputs").

numpy PCG64, seed = 1234 + 1000*stream + frame.  An analytic scene (floor + back wall at 3.5 m + sphere r=0.6 m at
2 m) carries a procedural texture (sigma-1.5 blurred uniform noise, contrast x1.8, + random filled rectangles/discs
for FAST corners); every frame RENDERS that scene from pose k (gray) and ray-casts it (depth + Kinect-style noise,
2 % invalid pixels), so images, depth and poses are geometrically consistent and consecutive frames really match.
Poses: circle of radius 0.3 m, 0.5 deg/frame yaw.  Pure numpy so it needs neither cv2 nor a GPU.
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


_TEX_CACHE = {}
TEX = 2048            # texture side (texels)
TEXEL = 0.005         # metres per texel on the planes


def _texture(stream):
    """procedural surface texture: sigma-1.5 blurred noise (contrast x1.8) + random rectangles / discs"""
    if stream not in _TEX_CACHE:
        rng = np.random.default_rng(99 + stream)
        base = (_box_blur_noise(rng, TEX, TEX) - 127.5) * 1.8 + 110.0
        yy, xx = np.mgrid[0:TEX, 0:TEX]
        for _ in range(500):
            g = float(rng.integers(0, 256))
            if rng.random() < 0.6:
                x0, y0 = int(rng.integers(0, TEX - 8)), int(rng.integers(0, TEX - 8))
                ww, hh = int(rng.integers(8, 90)), int(rng.integers(8, 90))
                base[y0:y0 + hh, x0:x0 + ww] = g
            else:
                cx, cy, r = int(rng.integers(0, TEX)), int(rng.integers(0, TEX)), int(rng.integers(5, 40))
                y0, y1, x0, x1 = max(cy - r, 0), min(cy + r + 1, TEX), max(cx - r, 0), min(cx + r + 1, TEX)
                sub = base[y0:y1, x0:x1]
                sub[(xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= r * r] = g
        _TEX_CACHE[stream] = np.clip(base, 0, 255).astype(np.float32)
    return _TEX_CACHE[stream]


def pose(frame, eye=0.0):
    """Twc (3x4 float32): circle radius 0.3 m in the x-z plane, yaw 0.5 deg/frame.  `eye` shifts the camera
    along its own x axis (metres): the right eye of a rectified stereo rig is pose(frame, baseline)."""
    yaw = math.radians(0.5 * frame)
    c, s = math.cos(yaw), math.sin(yaw)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64)
    ang = 0.02 * frame
    t = np.array([0.3 * math.cos(ang) - 0.3, 0.0, 0.3 * math.sin(ang)], np.float64) + R[:, 0] * eye
    return np.concatenate([R, t[:, None]], 1).astype(np.float32)


_RAY_CACHE = {}


def _raycast(frame, w, h, eye=0.0):
    """analytic scene seen from pose(frame): back wall z=3.5, floor y=1.2 (y down), sphere r=0.6 at (0.1,0.3,2.0).
    Returns camera depth z (0 where nothing is hit), surface id (0 none, 1 wall, 2 floor, 3 sphere) and world points."""
    key = (frame, w, h, eye)
    if key in _RAY_CACHE:
        return _RAY_CACHE[key]
    K = intrinsics(w, h)
    Twc = pose(frame, eye).astype(np.float64)
    R, o = Twc[:, :3], Twc[:, 3]
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d = np.stack([(u - K["cx"]) / K["fx"], (v - K["cy"]) / K["fy"], np.ones_like(u)], -1)   # camera rays with z = 1
    dw = d @ R.T
    best = np.full((h, w), np.inf); surf = np.zeros((h, w), np.int8)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (3.5 - o[2]) / dw[..., 2]; s[(s <= 0) | ~np.isfinite(s)] = np.inf
        m = s < best; best[m] = s[m]; surf[m] = 1
        s = (1.2 - o[1]) / dw[..., 1]; s[(s <= 0) | ~np.isfinite(s)] = np.inf
        m = s < best; best[m] = s[m]; surf[m] = 2
        cc = np.array([0.1, 0.3, 2.0]) - o
        a = (dw * dw).sum(-1); b = -2 * (dw @ cc); c0 = cc @ cc - 0.36
        disc = b * b - 4 * a * c0
        s = (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a); s[(disc < 0) | (s <= 0)] = np.inf
        m = s < best; best[m] = s[m]; surf[m] = 3
    z = best.copy(); z[~np.isfinite(z)] = 0.0          # rays have z = 1, so the ray parameter is the camera depth
    P = o + dw * np.where(np.isfinite(best), best, 0.0)[..., None]
    if len(_RAY_CACHE) > 8:
        _RAY_CACHE.clear()
    _RAY_CACHE[key] = (z, surf, P)
    return z, surf, P


def gray_frame(frame, w=640, h=480, stream=0, eye=0.0):
    """u8 gray image of frame `frame`: the textured analytic scene rendered from pose(frame) (bilinear texture
    lookup) + N(0,2) sensor noise -- geometrically consistent with depth_frame() and pose()."""
    rng = np.random.default_rng(1234 + 1000 * stream + frame + (500 if eye else 0))
    tex = _texture(stream)
    z, surf, P = _raycast(frame, w, h, eye)
    # texture coordinates per surface (texels); the texel pitch is scaled with the image so VGA and 1080p see alike detail
    pitch = TEXEL * 640.0 / w
    tu = np.zeros((h, w)); tv = np.zeros((h, w))
    m = surf == 1; tu[m] = (P[..., 0][m] + 5.0) / pitch; tv[m] = (P[..., 1][m] + 4.0) / pitch
    m = surf == 2; tu[m] = (P[..., 0][m] + 5.0) / pitch; tv[m] = (P[..., 2][m] + 1.0) / pitch + 700
    m = surf == 3
    q = P[m] - np.array([0.1, 0.3, 2.0])
    tu[m] = (np.arctan2(q[:, 0], -q[:, 2]) + np.pi) * 0.6 / pitch + 300; tv[m] = (np.arcsin(np.clip(q[:, 1] / 0.6, -1, 1)) + 2.0) * 0.6 / pitch
    x0 = np.floor(tu).astype(np.int64); y0 = np.floor(tv).astype(np.int64)
    fx = (tu - x0).astype(np.float32); fy = (tv - y0).astype(np.float32)
    x0 %= TEX; y0 %= TEX; x1 = (x0 + 1) % TEX; y1 = (y0 + 1) % TEX
    img = (tex[y0, x0] * (1 - fx) + tex[y0, x1] * fx) * (1 - fy) + (tex[y1, x0] * (1 - fx) + tex[y1, x1] * fx) * fy
    img[surf == 0] = 30.0
    img = img + rng.normal(0, 2.0, size=(h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def render_gray_parallel(n, w, h, stream, workers):
    """gray_frame(f, w, h, stream) for f in range(n), rendered by `workers` child interpreters (each takes every workers-th frame and leaves its
    slice as a .npy file in a temporary directory).  Returns the list of frames, or None if that did not work (the caller then renders here)."""
    import os, subprocess, sys, tempfile, pathlib
    root = str(pathlib.Path(__file__).resolve().parent.parent)
    try:
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
            code = ("import sys, numpy as np; sys.path.insert(0, %r); from plvs_b200 import synth; k = int(sys.argv[1]); "
                    "np.save(%r + '/g%%d.npy' %% k, np.stack([synth.gray_frame(f, %d, %d, %d) for f in range(k, %d, %d)]))" % (root, tmp, w, h, stream, n, workers))
            procs = [subprocess.Popen([sys.executable, "-c", code, str(k)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for k in range(workers)]
            if any(pr.wait() != 0 for pr in procs):
                return None
            out = [None] * n
            for k in range(workers):
                part = np.load(os.path.join(tmp, "g%d.npy" % k))
                for j, f in enumerate(range(k, n, workers)):
                    out[f] = part[j]
            return out
    except Exception:
        return None


def depth_frame(frame, w=640, h=480, stream=0, noise=True):
    """float32 depth (metres) of the analytic scene seen from pose(frame): Kinect-style noise, 2 % invalid (0) pixels."""
    z = _raycast(frame, w, h)[0].copy()
    rng = np.random.default_rng(777 + 1234 + 1000 * stream + frame)
    if noise:
        sig = 0.0012 + 0.0019 * (z - 0.4) ** 2
        z = z + rng.normal(0, 1, size=z.shape) * sig * (z > 0)
    z[rng.random(size=z.shape) < 0.02] = 0.0
    return z.astype(np.float32)


def bgr_frame(frame, w=640, h=480, stream=0):
    g = gray_frame(frame, w, h, stream)
    return np.stack([g, np.roll(g, 3, 1), 255 - g], -1).copy()
