"""Development helper: wall-clock breakdown of the tracking-side calls (extract / 2x projection / triangulation) and of
the TSDF calls, run serially on one stream.  python tools/track_breakdown.py [steps]"""
import sys, time, pathlib
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from plvs_b200.pipeline import StreamData, HotPath
from plvs_b200.matcher import Frame

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = 8
data = StreamData(1 + (steps + 2) * B, 640, 480, stream=0, pinned=True)
hp = HotPath(data, 2000, 0.01, 5.0, max_blocks=49152, device=0, batch=B)
hp.prepare()
hp.tsdf.integrate(data.depth[0], data.poses[0], data.bgr[0])
acc = {}
def tick(name, t0):
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
d = data
for s in range(steps + 2):
    if s == 2:
        acc.clear()
    f0 = 1 + s * B
    t = time.perf_counter()
    mono, kps, descs = hp.ex.extract_batch(d.gray[f0:f0 + B])
    tick("extract_batch", t)
    sf, s2 = hp.ex.mvScaleFactor, hp.ex.mvLevelSigma2
    for b in range(B):
        f = f0 + b
        p = hp.prepared[f]
        t = time.perf_counter()
        dv = hp.ex.device_result(b)
        cur = Frame(None, None, d.w, d.h, sf, s2, bf=d.K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, 0, dv.cache_key))
        tick("frame_setup", t)
        t = time.perf_counter()
        n1, a1 = hp.m_track.SearchByProjectionLast(cur, p["ql"], 15.0)
        tick("proj_last", t)
        t = time.perf_counter()
        claimed = (a1 >= 0).astype(np.uint8)
        n2, a2 = hp.m_track.SearchByProjectionMap(cur, p["qm"], 3.0, claimed=claimed, nnratio=0.8)
        tick("proj_map", t)
        t = time.perf_counter()
        last = hp.frames[f - 1]
        n3, m12 = hp.m_tri.SearchForTriangulation(Frame(kps[b], descs[b], d.w, d.h, sf, s2, uright=hp.frames[f].uright, bf=d.K["bf"]), last,
                                                  p["fv1"], p["fv2"], p["has1"], p["has2"], p["F12"], p["ep"], False, False)
        tick("triangulation", t)
    t = time.perf_counter()
    for b in range(B):
        hp.tsdf.integrate(d.depth[f0 + b], d.poses[f0 + b], d.bgr[f0 + b])
    hp.tsdf.stats()
    tick("tsdf_8_scans", t)
tot = sum(acc.values())
print({k: round(v / steps * 1e3, 3) for k, v in acc.items()}, "ms/step; total %.3f ms/step" % (tot / steps * 1e3))
st = hp.ex.last_stats()
print("orb last_stats", st)
