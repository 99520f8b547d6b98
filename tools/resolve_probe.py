import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from plvs_b200 import synth, scenario
from plvs_b200.orb import ORBextractor
from plvs_b200.matcher import ORBmatcher
ex = ORBextractor(2000, 1.2, 8, 20, 7)
K = synth.intrinsics(640, 480)
fr=[]
for f in (10, 11):
    mono, kp, desc = ex(synth.gray_frame(f))
    fr.append(scenario.make_frame(kp, desc, synth.depth_frame(f), K, ex.GetScaleFactors()))
q,_ = scenario.last_queries(fr[0], fr[1], K, synth.pose(10), synth.pose(11))
qm,_ = scenario.map_queries(fr[0], fr[1], K, synth.pose(10), synth.pose(11))
m = ORBmatcher(0.9, True); m2 = ORBmatcher(0.8, True)
for i in range(5):
    t=time.perf_counter(); n,a = m.SearchByProjectionLast(fr[1], q, 15.0); t1=time.perf_counter()-t
    t=time.perf_counter(); n2,a2 = m2.SearchByProjectionMap(fr[1], qm, 3.0); t2=time.perf_counter()-t
print('last', n, t1*1e6, 'us; map', n2, t2*1e6, 'us')
