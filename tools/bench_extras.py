#!/usr/bin/env python3
"""Timing of the widened rows that bench.py's headline step does not contain: one JSON line per entry point with the wall time of the
synchronous C-ABI call (host buffers in, results out), the workload size and -- where the work is bandwidth-shaped -- the algorithmic bytes it moves.
Run on the GPU box (tools/gpu_next_round.sh); nothing here imports the oracle."""
import json
import sys
import time
import pathlib

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from plvs_b200 import synth, scenario, tsdf as T      # noqa: E402
from plvs_b200.orb import ORBextractor              # noqa: E402
from plvs_b200.matcher import ORBmatcher            # noqa: E402


def timed(fn, warm=3, reps=20):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(reps):
        a = time.perf_counter(); fn(); t.append((time.perf_counter() - a) * 1e3)
    return float(np.median(t)), float(np.min(t))


def emit(name, ms, best, **kw):
    print(json.dumps(dict(entry=name, ms_median=round(ms, 4), ms_min=round(best, 4), **kw)), flush=True)


def main():
    w, h = 640, 480
    K = synth.intrinsics(w, h)
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    ex5 = ORBextractor(5000, 1.2, 8, 20, 7)                 # the monocular initialiser's extractor (5 x nFeatures)
    frames, frames5 = [], []
    for f in (10, 11):
        img = synth.gray_frame(f)
        _, kp, desc = ex(img); frames.append(scenario.make_frame(kp, desc, synth.depth_frame(f), K, ex.GetScaleFactors()))
        _, kp, desc = ex5(img); frames5.append(scenario.make_frame(kp, desc, synth.depth_frame(f), K, ex5.GetScaleFactors()))

    # rank 1: SearchForInitialization
    m = ORBmatcher(0.9, True)
    f1, f2 = frames5
    prev = np.stack([f1.keys["x"], f1.keys["y"]], 1)
    ms, best = timed(lambda: m.SearchForInitialization(f1, f2, prev, 100))
    emit("plvs_match_initialization", ms, best, n1=int(f1.n), n2=int(f2.n), level0=int((f1.keys["octave"] == 0).sum()), matches=int(m.SearchForInitialization(f1, f2, prev, 100)[0]))

    # rank 2: undistort on the resident keypoints
    ms, best = timed(lambda: ex.UndistortKeyPoints((517.306408, 516.469215, 318.643040, 255.313989), np.array((0.262383, -0.953104, -0.005358, 0.002628, 1.163314), np.float32)))
    emit("plvs_orb_undistort", ms, best, n=int(frames[1].n))

    # rank 2: 16-bit depth in front of the TSDF, against the float path (same scan)
    p = T.default_params(voxel_resolution=0.01, use_carving=1, near_plane=0.1, far_plane=5.0, max_blocks=60000, use_color=1)
    d = synth.depth_frame(0); c = synth.bgr_frame(0)
    d16 = np.clip(d * 5000.0, 0, 65535).astype(np.uint16)
    for name, call in (("plvs_tsdf_integrate_depth(f32 host)", lambda g: g.integrate(d, synth.pose(0), c)),
                       ("plvs_tsdf_integrate_depth_u16", lambda g: g.integrate_u16(d16, 1.0 / 5000.0, synth.pose(0), c))):
        g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        ms, best = timed(lambda: (call(g), g.stats()))
        emit(name, ms, best, n_blocks=int(g.stats()["n_blocks"]))

    # rank 3: mesh read-out of a VGA / 1 cm map after 10 scans
    g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    for f in range(10):
        g.integrate(synth.depth_frame(f), synth.pose(f), synth.bgr_frame(f))
    nb = g.stats()["n_blocks"]
    ms, best = timed(lambda: g.UpdateMesh(), reps=10)
    nm, nv = g.UpdateMesh()
    algo = nb * 4096 * 8 + nv * 36
    emit("plvs_tsdf_update_meshes", ms, best, n_blocks=int(nb), meshes=int(nm), verts=int(nv), algorithmic_bytes=int(algo), algorithmic_GBps=round(algo / best / 1e6, 1))
    ms, best = timed(lambda: g.GetMeshes(), reps=5)
    emit("plvs_tsdf_get_meshes(D2H)", ms, best, bytes=int(nv * 36))

    # rank 4: bag-of-words transform with an ORBvoc-sized synthetic vocabulary (k = 10, L = 6: 1.1 M nodes)
    try:
        from plvs_b200.bow import ORBVocabulary
        k, L = 10, 6
        rng = np.random.default_rng(0)
        n_nodes = (k ** (L + 1) - 1) // (k - 1)
        parent = np.zeros(n_nodes, np.int32); parent[1:] = (np.arange(1, n_nodes) - 1) // k
        first_leaf = (k ** L - 1) // (k - 1)
        word = np.full(n_nodes, -1, np.int32); word[first_leaf:] = np.arange(n_nodes - first_leaf)
        desc = rng.integers(0, 256, (n_nodes, 32), dtype=np.uint8)
        weight = np.zeros(n_nodes, np.float64); weight[first_leaf:] = rng.uniform(0.5, 9.0, n_nodes - first_leaf)
        voc = ORBVocabulary(); voc.create(k, L, 0, 0, parent, word, desc, weight)
        fd = frames[1].desc
        ms, best = timed(lambda: voc.transform(fd, 4))
        emit("plvs_voc_transform", ms, best, n=int(len(fd)), k=k, L=L, hamming_per_feature=k * L, algorithmic_bytes=int(len(fd) * (32 + k * L * 32)))
    except Exception as e:          # noqa: BLE001
        emit("plvs_voc_transform", -1.0, -1.0, error=str(e))


if __name__ == "__main__":
    main()
