#!/usr/bin/env python3
"""Stereo front end (BASELINE.json configs[4], the ORB part: EuRoC-sized 752x480 pairs, 1200 features per eye, ComputeStereoMatches) on 1 GPU (two
extractor handles) and split over 2 GPUs (left eye + stereo matching on GPU 0, right eye on GPU 1, pyramid / keypoints read over NVLink).
One JSON line per N.   python tools/bench_stereo_split.py [--frames 200]      (gpurun --gpus 2)"""
import argparse, json, pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
from plvs_b200 import synth, _lib
from plvs_b200.stereo import StereoFrontEnd
from plvs_b200.pipeline import PinnedArray

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=200); ap.add_argument("--warmup", type=int, default=20)
a = ap.parse_args()
w, h, nfeat, baseline = 752, 480, 1200, 0.11
K = synth.intrinsics(w, h)
n_img = 16
L = PinnedArray((n_img, h, w), np.uint8); R = PinnedArray((n_img, h, w), np.uint8)
for f in range(n_img):
    L.array[f] = synth.gray_frame(f, w, h); R.array[f] = synth.gray_frame(f, w, h, eye=baseline)
ngpu = _lib.load().plvs_device_count()
for devices in ((0, 0), (0, 1)):
    if devices[1] >= ngpu:
        print(json.dumps({"metric": "stereo frames/s (2x ORB extract + ComputeStereoMatches) 752x480", "n_gpus": 2, "unavailable": "one GPU on this box"}))
        continue
    fe = StereoFrontEnd(nfeat, w, h, baseline, K["fx"], devices=devices)
    kept = 0
    for f in range(a.warmup):
        fe(L.array[f % n_img], R.array[f % n_img])
    t0 = time.perf_counter()
    for f in range(a.frames):
        kept += fe(L.array[f % n_img], R.array[f % n_img])[6]
    dt = time.perf_counter() - t0
    print(json.dumps({"metric": "stereo frames/s (2x ORB extract + ComputeStereoMatches) 752x480", "value": a.frames / dt, "unit": "frames/s", "n_gpus": len(set(devices)),
                      "frames": a.frames, "ms_per_frame": 1e3 * dt / a.frames, "stereo_points_per_frame": kept / a.frames, "data": "synthetic",
                      "timing": "wall clock around synchronous calls (every call returns with its result on the host); pinned host images in, host keypoints out",
                      "config": {"workload": "synthetic 752x480 stereo pairs, ORB nFeatures=1200 per eye, Frame::ComputeStereoMatches; the libsgm disparity path of configs[4] is not built",
                                 "devices": list(devices)}}), flush=True)
