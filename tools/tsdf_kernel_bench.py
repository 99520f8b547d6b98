"""Development microbenchmark: the TSDF depth-scan kernels alone (CUDA events inside the library, profiling mask 4),
device-resident inputs, L2 flushed between scans.  Usage: python tools/tsdf_kernel_bench.py [frames] [w h voxel]"""
import ctypes as C
import sys, pathlib
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from plvs_b200 import _lib, synth, tsdf as T

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
voxel = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
lib = _lib.load()
K = synth.intrinsics(w, h)
p = T.default_params(voxel_resolution=voxel, use_carving=1, near_plane=0.1, far_plane=5.0, max_blocks=int(sys.argv[5]) if len(sys.argv) > 5 else 49152, use_color=1)
g = T.ChiselServer(p)
g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
dev = torch.device("cuda", 0)
depth = [torch.from_numpy(synth.depth_frame(f, w, h)).to(dev) for f in range(n)]
bgr = [torch.from_numpy(synth.bgr_frame(f, w, h)).to(dev) for f in range(n)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
warm = n // 3
for f in range(n):
    if f == warm:
        g.stats()
        lib.plvs_set_profiling(4)
        lib.plvs_tsdf_kernel_times(g._h, None, None, 1)
    flush.fill_(f & 255)
    torch.cuda.synchronize()
    pose = np.ascontiguousarray(synth.pose(f), np.float32).reshape(12)
    rc = lib.plvs_tsdf_integrate_depth(g._h, C.c_void_p(depth[f].data_ptr()), w, h, C.c_void_p(bgr[f].data_ptr()), w * 3, 3,
                                       pose.ctypes.data_as(C.c_void_p), T.SCAN_COLOR, 1)
    _lib.check(rc, "integrate")
st = g.stats()
ms = np.zeros(12, np.float32); cnt = np.zeros(12, np.int32)
lib.plvs_tsdf_kernel_times(g._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
lib.plvs_set_profiling(0)
names = ("depth_tiles", "classify", "integrate", "commit")
per = {nm: round(float(ms[i]) / max(int(cnt[i]), 1) * 1000, 1) for i, nm in enumerate(names)}
upd = st["total_updated"] / max(st["total_integrations"], 1)
print("us/scan", per, "n", int(cnt[2]), "updated/scan %.0f visited/scan %.0f blocks %d" % (upd, st["total_candidates"] / max(st["total_integrations"], 1), st["n_blocks"]),
      "integrate GB/s(alg) %.0f" % ((w * h * 7 + upd * 4096 * 24) / (per["integrate"] * 1e-6) / 1e9))
