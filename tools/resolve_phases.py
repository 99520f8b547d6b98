#!/usr/bin/env python3
"""Where the one-CTA claim resolution spends its time on the bench stream: per search, rounds / list walks / SM cycles per phase
(plvs_match_last_phase_cycles), for a few consecutive frames.  python tools/resolve_phases.py [nframes]"""
import ctypes as C, pathlib, sys
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from plvs_b200 import _lib                                        # noqa: E402
from plvs_b200.matcher import Frame                               # noqa: E402
from plvs_b200.pipeline import StreamData, HotPath                # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
d = StreamData(1 + n, 640, 480, stream=0, pinned=False)
hp = HotPath(d, nfeatures=2000, batch=1)
hp.prepare()
lib = _lib.load()
sf, s2 = hp.ex.mvScaleFactor, hp.ex.mvLevelSigma2
ph = (C.c_int32 * 4)()
print("frame search        nq  rounds walks   table  compare   re-eval  rounds+wrapup [SM cycles, thread 0]")
for f in range(1, 1 + n):
    p = hp.prepared[f]
    hp.ex(d.gray[f])
    dv = hp.ex.device_result(0)
    cur = Frame(None, None, d.w, d.h, sf, s2, uright=hp.frames[f].uright, bf=d.K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, 0, dv.cache_key, dv.grid_cell_start, dv.grid_sorted))
    for rep in range(2):
        n1, a1 = hp.m_track.SearchByProjectionLast(cur, p["ql"], 15.0)
    lib.plvs_match_last_phase_cycles(hp.m_track._h, ph)
    print(f"{f:5d} last     {len(p['ql']):6d} {hp.m_track.last_stats()[0]:6d} {lib.plvs_match_last_walks(hp.m_track._h):6d} {ph[0]:8d} {ph[1]:8d} {ph[2]:8d} {ph[3]:8d}")
    claimed = (a1 >= 0).astype(np.uint8)
    for rep in range(2):
        hp.m_track.SearchByProjectionMap(cur, p["qm"], 3.0, claimed=claimed, nnratio=0.8)
    lib.plvs_match_last_phase_cycles(hp.m_track._h, ph)
    print(f"{f:5d} map      {len(p['qm']):6d} {hp.m_track.last_stats()[0]:6d} {lib.plvs_match_last_walks(hp.m_track._h):6d} {ph[0]:8d} {ph[1]:8d} {ph[2]:8d} {ph[3]:8d}")
