#!/usr/bin/env python3
"""bench.py -- frames/s of the PLVS per-frame hot path (ORB extract + Hamming match + Chisel TSDF) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of `frames_per_step` consecutive RGB-D frames of a synthetic stream (default: BASELINE.json configs[1], 640x480, ORB
nFeatures=2000 + Chisel TSDF 1 cm voxels; `--config c3` = configs[2], 1920x1080, nFeatures=4000, 5 mm); every frame is extracted, matched
against the previous frame (SearchByProjection Cur<-Last), against its local map points (SearchByProjection F<-map), triangulation-searched
against the previous frame and integrated into the TSDF.  One camera stream per rank (weak scaling, no collective on the data path).
Prints ONE JSON line (rank 0).
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time
import pathlib

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np

CONFIGS = {
    # BASELINE.json configs[1] / configs[2] (the ORB part; line features are outside SURVEY.md §8's scope table)
    "c2": dict(width=640, height=480, nfeatures=2000, voxel=0.01, far=5.0, batch=8, max_blocks=49152),
    "c3": dict(width=1920, height=1080, nfeatures=4000, voxel=0.005, far=5.0, batch=4, max_blocks=400000),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    for k, t in (("batch", int), ("width", int), ("height", int), ("nfeatures", int), ("voxel", float), ("far", float), ("max_blocks", int)):
        ap.add_argument("--" + k.replace("_", "-"), type=t, default=None)
    ap.add_argument("--repeats", type=int, default=5, help="timed passes per arm; the median is reported")
    ap.add_argument("--driver", default="native", choices=["native", "python"], help="stage threads in C++ (plvs_pipeline_run) or in Python")
    ap.add_argument("--dataset", default=None, help="directory of a TUM RGB-D sequence (rgb.txt, depth.txt, groundtruth.txt or an associations file): BASELINE.json configs[0] on real data; "
                                                    "without it the stream is synthetic")
    ap.add_argument("--rank-streams", default="same", choices=["same", "distinct"], help="N > 1: every rank the same synthetic stream (identical work per GPU) or one stream per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=2)
    a = ap.parse_args()
    for k, v in CONFIGS[a.config].items():
        if getattr(a, k) is None:
            setattr(a, k, v)
    return a


def metric_name(a):
    return f"frames/sec (extract+match+TSDF) {a.width}x{a.height} RGB-D"


def workload_config(a, extra=None):
    c = {"workload": f"{'TUM RGB-D sequence (--dataset)' if getattr(a, 'dataset', None) else 'synthetic'} {a.width}x{a.height} RGB-D stream, ORB nFeatures={a.nfeatures} (8 levels, 1.2, FAST 20/7) + "
                     f"SearchByProjection(Cur,Last) th=15 + SearchByProjection(F,map) th=3 + SearchForTriangulation + "
                     f"Chisel TSDF {a.voxel * 100:g} cm voxels (colour depth-scan, carving on, planes 0.1-{a.far:g} m), every frame integrated",
         "baseline_config": {"c2": "BASELINE.json configs[1]", "c3": "BASELINE.json configs[2] (ORB part)"}[a.config],
         "frames_per_step": a.batch, "streams_per_gpu": 1, "parallelism": f"1 camera stream per GPU x{a.gpus}",
         "rank_streams": ("every rank processes the same synthetic stream: identical work per GPU (weak scaling)" if getattr(a, "rank_streams", "same") == "same"
                          else "one synthetic stream per rank (different content, different work per GPU)")}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------------------
# CPU arm.  Nothing in this section touches libplvs_b200.so: parameters come from the oracle side, the frame / query helpers are numpy.
# ------------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def _quiet_stdout():
    """the reference's open_chisel prints per scan; bench.py must print exactly one JSON line"""
    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1); os.close(null); os.close(saved)


class CpuPath:
    """The reference's CPU path on one stream.  With oracle/_ref present: extraction = the reference's own src/ORBextractor.cc (liborb_ref.so),
    the three searches = its own src/ORBmatcher.cc (libmatch_ref.so), TSDF = its own open_chisel (libchisel_ref.so, single-threaded like
    Chisel.h:91) or the restated TSDF over `threads` OpenMP threads.  Without oracle/_ref everything is the restatement (kind "port")."""

    def __init__(self, a, threads):
        from oracle import orb as O, match as OM, tsdf as OT
        from plvs_b200 import synth
        self.a, self.O, self.OM, self.OT, self.synth = a, O, OM, OT, synth
        self.K = synth.intrinsics(a.width, a.height)
        self.seq = None
        if getattr(a, "dataset", None):          # the same frames the b200 arm reads (loaded on first use, enough of them for both arms' longest run)
            from plvs_b200.pipeline import StreamData
            self.seq = StreamData.from_tum(a.dataset, 1 + (min(a.warmup, 1) + a.steps) * a.batch, pinned=False)
            self.K = self.seq.K
        self.tab = O.Tables(a.nfeatures)
        self.have_ref = O.ref_available() and OM.ref_available() and OT.ref_available()
        self.ex = O.RefExtractor(a.nfeatures) if self.have_ref else None
        p = OT.default_params(voxel_resolution=a.voxel, use_carving=1, near_plane=0.1, far_plane=a.far, max_blocks=1 << 20, use_color=1)
        self.port_map = OT.Map(p, threads=threads); self.port_map.set_camera(self.K["fx"], self.K["fy"], self.K["cx"], self.K["cy"], a.width, a.height)
        self.ref_map = None
        if self.have_ref:
            self.ref_map = OT.RefMap(p); self.ref_map.set_camera(self.K["fx"], self.K["fy"], self.K["cx"], self.K["cy"], a.width, a.height)
        self.prev = None
        self.t = dict(extract_s=0.0, match_s=0.0, tsdf_ref_s=0.0, tsdf_port_s=0.0)
        self.n = dict(frames=0, tsdf_ref=0, tsdf_port=0)

    def frame(self, f, tsdf="port", timed=True):
        """one frame through extract + the three searches + one TSDF integration (`tsdf`: "ref", "port" or "both")"""
        from plvs_b200 import scenario
        from plvs_b200.matcher import featvec
        a, O, OM, synth = self.a, self.O, self.OM, self.synth
        if self.seq is not None:
            img, depth, bgr = self.seq.gray[f], self.seq.depth[f], self.seq.bgr[f]
            pose = lambda i: self.seq.poses[i]
        else:
            img = synth.gray_frame(f, a.width, a.height); depth = synth.depth_frame(f, a.width, a.height); bgr = synth.bgr_frame(f, a.width, a.height)
            pose = synth.pose
        t0 = time.perf_counter()
        if self.ex is not None:
            kp, desc, mono = self.ex(img)
        else:
            kp, desc, mono, _ = O.extract_port(img, a.nfeatures)
        t1 = time.perf_counter()
        cur = scenario.make_frame(kp, desc, depth, self.K, self.tab.scale); cur.level_sigma2 = self.tab.sigma2
        t_match = 0.0
        if self.prev is not None:
            prev = self.prev
            ql, _ = scenario.last_queries(prev, cur, self.K, pose(f - 1), pose(f))
            qm, _ = scenario.map_queries(prev, cur, self.K, pose(f - 1), pose(f), seed=f)
            fv1, fv2 = featvec(scenario.node_ids(cur.desc)), featvec(scenario.node_ids(prev.desc))
            F12, ep = scenario.fundamental(self.K, pose(f), pose(f - 1))
            z0, z1 = np.zeros(cur.n, np.uint8), np.zeros(prev.n, np.uint8)
            if self.have_ref:
                qc, z = OM.canonical_last_queries(ql)
                t2 = time.perf_counter()
                n1, a1 = OM.ref_search_by_projection_last(cur, qc, z, 15.0)
                OM.ref_search_by_projection_map(cur, qm, 3.0, 0.8, claimed=(a1 >= 0).astype(np.uint8))
                OM.ref_search_for_triangulation(cur, prev, fv1, fv2, z0, z1, F12, ep, check_ori=False)
            else:
                t2 = time.perf_counter()
                n1, a1 = OM.search_by_projection_last(cur, ql, 15.0)
                OM.search_by_projection_map(cur, qm, 3.0, 0.8, claimed=(a1 >= 0).astype(np.uint8))
                OM.search_for_triangulation(cur, prev, fv1, fv2, z0, z1, F12, ep, check_ori=False)
            t_match = time.perf_counter() - t2
        self.prev = cur
        t3 = time.perf_counter()
        t_ref = t_port = None
        if tsdf in ("port", "both"):
            self.port_map.integrate(depth, pose(f), bgr)
            t_port = time.perf_counter() - t3
        if tsdf in ("ref", "both") and self.ref_map is not None:
            t4 = time.perf_counter()
            with _quiet_stdout():
                self.ref_map.integrate(depth, pose(f), bgr)
            t_ref = time.perf_counter() - t4
        if timed:
            self.t["extract_s"] += t1 - t0; self.t["match_s"] += t_match; self.n["frames"] += 1
            if t_port is not None:
                self.t["tsdf_port_s"] += t_port; self.n["tsdf_port"] += 1
            if t_ref is not None:
                self.t["tsdf_ref_s"] += t_ref; self.n["tsdf_ref"] += 1
        return t_ref, t_port

    def per_frame(self):
        n = max(self.n["frames"], 1)
        out = dict(extract_s=self.t["extract_s"] / n, match_s=self.t["match_s"] / n)
        if self.n["tsdf_port"]:
            out["tsdf_port_s"] = self.t["tsdf_port_s"] / self.n["tsdf_port"]
        if self.n["tsdf_ref"]:
            out["tsdf_ref_s"] = self.t["tsdf_ref_s"] / self.n["tsdf_ref"]
        return out


def run_reference_arm(a):
    """The reference's own CPU implementation of the path on this box's host cores, K steps of `frames_per_step` frames like the b200 arm.
    Every frame of a step goes through extraction and the three searches; the TSDF -- 5-6 s per scan in the reference's brute-force
    open_chisel at C2 -- runs through the reference's own code for ONE frame of each step (the bounded sample) and through the restated TSDF
    spread over all host threads (OpenMP) for the other frames.  The step time is the measured wall time of exactly that work, so the line
    UNDERSTATES the reference's cost (value = frames_per_step / step time); cpu_baseline.sample spells the per-stage seconds out."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    threads = os.cpu_count() or 1
    cv2.setNumThreads(threads)
    cp = CpuPath(a, threads)
    B = a.batch
    cp.frame(0, tsdf="both", timed=False)                  # map seed, untimed (both maps)
    f = 1
    for s in range(min(a.warmup, 1)):                      # bounded warm-up: one step, restated TSDF only
        for b in range(B):
            cp.frame(f, tsdf="port", timed=False); f += 1
    step_s = []
    for s in range(a.steps):
        t0 = time.perf_counter()
        twice = 0.0
        for b in range(B):
            t_ref, t_port = cp.frame(f, tsdf="both" if (b == 0 and cp.have_ref) else "port"); f += 1
            if t_ref is not None and t_port is not None:
                twice += t_port        # this frame ran BOTH TSDFs (the restated map has to see every frame): only the reference's counts towards the step
        step_s.append(time.perf_counter() - t0 - twice)
    pf = cp.per_frame()
    mean_step = float(np.mean(step_s))
    fps = B / mean_step
    front_end = pf["extract_s"] + pf["match_s"]
    kind = "reference" if cp.have_ref else "port"
    note = ("every frame: extraction = the reference's own src/ORBextractor.cc (oracle/_ref/liborb_ref.so, OpenCV primitives restated in C and pinned to cv2), the three "
            "searches = its own src/ORBmatcher.cc (libmatch_ref.so); TSDF = its own open_chisel (libchisel_ref.so, single-threaded: Chisel.h:91 has its parallel_for "
            "commented out) on the first frame of each step, the restated TSDF over all host threads (OpenMP) on the other frames; rank 0 only") if cp.have_ref else \
           "CPU oracle (C restatement of the reference), TSDF chunks over OpenMP with all host threads; rank 0 only"
    all_ref_step = B * (front_end + pf.get("tsdf_ref_s", pf.get("tsdf_port_s", 0.0)))
    line = {"impl": "reference", "metric": metric_name(a), "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 * mean_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32+f32",
            "data": (f"TUM RGB-D sequence {cp.seq.stream}" if cp.seq is not None else "synthetic"), "config": workload_config(a, {"note": note}),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                             "sample": f"{a.steps} steps x {B} frames of the same stream; seconds/frame {json.dumps({k: round(v, 4) for k, v in pf.items()})}; "
                                       f"reference TSDF on 1 of {B} frames per step, restated TSDF ({threads} threads) on the rest",
                             "front_end_frames_per_s": 1.0 / front_end,
                             "all_frames_through_reference_tsdf_frames_per_s": B / all_ref_step,
                             "all_frames_through_restated_tsdf_frames_per_s": 1.0 / (front_end + pf.get("tsdf_port_s", pf.get("tsdf_ref_s", 0.0)))},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "wall_s": float(np.sum(step_s))}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period_ms=50, enabled=True):
        """index: one GPU index or a list of them (one process samples them all: under torchrun rank 0 watches every GPU of the job, the other
        ranks start nothing -- eight 20 Hz nvidia-smi loops on one host were part of what bent round 1's scaling curve)"""
        self.index, self.proc, self.lines, self.period_ms, self.enabled = index, None, [], period_ms, enabled

    def start(self):
        """one nvidia-smi process for the whole bench, started BEFORE any timed region: its start-up (NVML initialisation,
        a few hundred ms during which driver calls of this process can stall) must not fall into a timed window"""
        if not self.enabled:
            return
        try:
            idx = ",".join(str(i) for i in self.index) if isinstance(self.index, (list, tuple)) else str(self.index)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms), "-i", idx],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0:       # wait for the first sample = start-up is over
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        return len(self.lines)

    def window(self, i0):
        """clock statistics of the samples taken since mark() returned i0 (waits for one more sample: a timed region
        shorter than the sampling period still gets the reading taken right as it ends)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        n0 = len(self.lines)
        t0 = time.perf_counter()
        while len(self.lines) == n0 and time.perf_counter() - t0 < 0.3:
            time.sleep(0.01)
        sm, mx, reasons = [], [], set()
        for ln in self.lines[i0:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()
            self.proc = None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_table(lib, hp):
    import ctypes as C
    ktimes = {}
    ms = np.zeros(12, np.float32); cnt = np.zeros(12, np.int32)
    lib.plvs_tsdf_kernel_times(hp.tsdf._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
    for i, nme in enumerate(("tsdf.depth_tiles", "tsdf.classify", "tsdf.integrate", "tsdf.commit")):
        ktimes[nme] = (float(ms[i]), int(cnt[i]))
    om = np.zeros(12, np.float32); oc = np.zeros(12, np.int32)
    for ex in [hp.ex, hp.ex2] + list(getattr(hp, "ex_more", [])):
        if ex is not None:
            lib.plvs_orb_kernel_times(ex._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
            om += ms; oc += cnt
    for i, nme in enumerate(("orb.pyramid", "orb.fast_cells", "orb.compact", "orb.blur", "orb.orient_describe", "orb.distribute", "orb.frame_grid")):
        ktimes[nme] = (float(om[i]), int(oc[i]))
    mm = np.zeros(12, np.float32); mc = np.zeros(12, np.int32)
    for m in (hp.m_track, hp.m_map, hp.m_tri):
        lib.plvs_match_kernel_times(m._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
        mm += ms; mc += cnt
    for i, nme in enumerate(("match.grid", "match.candidates", "match.resolve", "match.triangulate")):
        ktimes[nme] = (float(mm[i]), int(mc[i]))
    return ktimes


def latency_block(hp, data, f_first, n):
    """per-call latency of the single-frame surfaces (include/ORBextractor.h:86, ORBmatcher.h:68-97, ChiselServer.h:190-286), one call at a
    time with nothing else on the GPU: what a sequential tracker sees per frame.  Host buffers in, host results out; median over n frames."""
    from plvs_b200.matcher import Frame
    d = data
    sf, s2 = hp.ex.mvScaleFactor, hp.ex.mvLevelSigma2
    t = {k: [] for k in ("extract_1_frame", "search_by_projection_last", "search_by_projection_map", "search_for_triangulation", "tsdf_scan")}
    hp.tsdf.stats()
    for f in range(f_first, f_first + n):
        p = hp.prepared[f]
        t0 = time.perf_counter()
        mono, kp, desc = hp.ex(d.gray[f])
        t1 = time.perf_counter()
        dv = hp.ex.device_result(0)
        cur = Frame(None, None, d.w, d.h, sf, s2, uright=hp.frames[f].uright, bf=d.K["bf"],
                    device_ptrs=(dv.n, dv.keys, dv.desc, 0, dv.cache_key, dv.grid_cell_start, dv.grid_sorted))
        t2 = time.perf_counter()
        n1, a1 = hp.m_track.SearchByProjectionLast(cur, p["ql"], 15.0)
        t3 = time.perf_counter()
        claimed = (a1 >= 0).astype(np.uint8)
        t4 = time.perf_counter()
        hp.m_track.SearchByProjectionMap(cur, p["qm"], 3.0, claimed=claimed, nnratio=0.8)
        t5 = time.perf_counter()
        k1 = Frame(kp, desc, d.w, d.h, sf, s2, uright=hp.frames[f].uright, bf=d.K["bf"])
        t6 = time.perf_counter()
        hp.m_tri.SearchForTriangulation(k1, hp.frames[f - 1], p["fv1"], p["fv2"], p["has1"], p["has2"], p["F12"], p["ep"], False, False)
        t7 = time.perf_counter()
        hp.tsdf.integrate(d.depth[f], d.poses[f], d.bgr[f]); hp.tsdf.stats()            # stats() waits for the scan
        t8 = time.perf_counter()
        for k, v in zip(t, (t1 - t0, t3 - t2, t5 - t4, t7 - t6, t8 - t7)):
            t[k].append(v * 1e3)
    out = {k + "_ms": round(float(np.median(v)), 4) for k, v in t.items()}
    out["frame_ms"] = round(float(np.median(np.sum([t[k] for k in t], 0))), 4)
    out["note"] = (f"median over {n} consecutive frames, one synchronous call at a time through the Python mirror of the single-frame surfaces "
                   "(host image in, host keypoints / assignments out); the pipeline value above overlaps these calls across stages and batches frames in the extractor")
    return out


def bind_to_gpu_socket(torch, local):
    """Host side of a rank, as a deployment would run it: the process (and with it the stage threads, the pinned staging buffers and the mapped result
    buffers it allocates afterwards) is bound to the CPUs that are local to its GPU's PCIe root (sysfs local_cpulist).  With eight ranks on a two-socket
    box, un-pinned ranks put half of their staging traffic on the inter-socket link.  PLVS_BENCH_NUMA=0 turns it off.  Returns a note for the JSON line."""
    if os.environ.get("PLVS_BENCH_NUMA", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return "off"
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = pathlib.Path(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read_text().strip()
        cpus = set()
        for part in txt.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 8:
            return f"not applied ({len(cpus)} local cpus)"
        os.sched_setaffinity(0, cpus)
        return f"bound to the {len(cpus)} cpus local to GPU {local} ({txt})"
    except Exception as e:
        return f"not applied ({type(e).__name__})"


def run_b200_arm(a):
    import torch
    sys.setswitchinterval(1e-4)          # python driver: 4 stage threads hand the GIL over quickly when a library call returns
    import torch.distributed as dist
    import ctypes as C
    from plvs_b200 import _lib, parallel
    from plvs_b200.pipeline import StreamData, HotPath

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_socket(torch, local)          # before the first pinned allocation: first touch puts the staging memory on the GPU's socket
    if world > 1:
        # NCCL writes its version banner (NCCL_DEBUG=VERSION/INFO) to stdout by default; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    B, K, W, R = a.batch, a.steps, a.warmup, max(1, a.repeats)
    n_lat = 0 if a.no_latency else 16
    nframes = 1 + (W + K) * B + n_lat            # frame 0 only seeds the map / the "last frame"
    # weak scaling: identical work per GPU -- every rank processes the same synthetic stream unless --rank-streams distinct gives each its own
    if a.dataset:
        data = StreamData.from_tum(a.dataset, nframes, pinned=True)
        if (data.w, data.h) != (a.width, a.height):
            raise SystemExit(f"--dataset: the sequence is {data.w}x{data.h}, the configuration {a.width}x{a.height}; pass --width / --height")
    else:
        data = StreamData(nframes, a.width, a.height, stream=parallel.stream_of_rank(rank, a.rank_streams), pinned=True)
    hp = HotPath(data, a.nfeatures, a.voxel, a.far, max_blocks=a.max_blocks, device=local, batch=B)
    hp.prepare()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    # N = 1: this GPU at 20 Hz; N > 1: rank 0 samples all GPUs of the job at 10 Hz (LOCAL_RANK i drives GPU i), the other ranks none
    sampler = ClockSampler(local if world == 1 else list(range(world)), 50 if world == 1 else 100, enabled=(rank == 0)); sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def stream(f0, nsteps, resident, timed):
        if a.driver == "native":
            return hp.run_stream_native(f0, nsteps, resident, flush=(flush.data_ptr(), flush.numel()) if timed else None)
        return hp.run_stream(f0, nsteps, resident, per_step=(lambda s: flush.fill_(s & 0xff)) if timed else None)

    def run(resident, timed_profile=0):
        hp.tsdf.Reset()
        hp.tsdf.integrate(data.depth[0], data.poses[0], data.bgr[0])           # seed (untimed)
        stream(1, W, resident, False)                                          # warm-up steps (untimed)
        lib.plvs_set_profiling(timed_profile)
        barrier()
        lib.plvs_io_bytes(None, None, 1)
        mark = sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        # K steps through the 4-stage pipeline (extract | projection searches | triangulation | TSDF); returns when drained
        agg = stream(1 + W * B, K, resident, True)                             # L2 flush once per step, inside the timed region
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        h2d, d2h = C.c_longlong(), C.c_longlong()
        lib.plvs_io_bytes(C.byref(h2d), C.byref(d2h), 0)
        clocks = sampler.window(mark)
        # job time = MAX over ranks of each rank's own device time; the line also lists every rank's own time (imbalance between the ranks)
        job_ms, per_rank = parallel.job_time(torch.tensor([e0.elapsed_time(e1)], device=dev))
        lib.plvs_set_profiling(0)
        return dict(ms=job_ms, wall=wall, clocks=clocks, agg=agg, h2d=h2d.value, d2h=d2h.value, per_rank_ms=per_rank)

    def reset_timers():
        lib.plvs_tsdf_kernel_times(hp.tsdf._h, None, None, 1)
        for ex in [hp.ex, hp.ex2] + list(getattr(hp, "ex_more", [])):
            if ex is not None:
                lib.plvs_orb_kernel_times(ex._h, None, None, 1)
        for m in (hp.m_track, hp.m_map, hp.m_tri):
            lib.plvs_match_kernel_times(m._h, None, None, 1)

    # ---- value arm: inputs resident in HBM ----------------------------------------------------------------
    hp.upload_inputs()
    # order: the per-kernel profile pass runs first (every kernel carries events; its result is not a bench value, it also settles every
    # data-dependent buffer size and the clocks), then the timed passes in which only the roofline kernel (k_integrate) carries events
    reset_timers()
    run(resident=True, timed_profile=7)
    ktimes = kernel_table(lib, hp)
    reset_timers()
    val = [run(resident=True, timed_profile=4 | 8) for _ in range(R)]
    ms = np.zeros(12, np.float32); cnt = np.zeros(12, np.int32)
    lib.plvs_tsdf_kernel_times(hp.tsdf._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
    integ_ms, integ_n = float(ms[2]), int(cnt[2])              # k_integrate over all R timed passes (and their warm-ups)
    frames = K * B
    vals = sorted(world * frames / (r["ms"] / 1000.0) for r in val)
    med = sorted(val, key=lambda r: r["ms"])[len(val) // 2]
    value, t_ms = world * frames / (med["ms"] / 1000.0), med["ms"]
    launches_per_step = hp.launches_per_frame() * B

    # roofline of the dominant kernel: TSDF voxel update.  Algorithmic bytes per launch:
    # depth 4*W*H + colour 3*W*H + n_updated_blocks * 4096 voxels * 12 B (sdf f32 + weight f32 + rgba) * 2 (read+write)
    tst = hp.tsdf.stats()          # running totals since the Reset at the start of the last pass (seed + warm-up + timed scans)
    upd_per_launch = tst["total_updated"] / max(tst["total_integrations"], 1)
    vis_per_launch = tst["total_candidates"] / max(tst["total_integrations"], 1)
    alg_bytes = a.width * a.height * (4 + 3) + upd_per_launch * 4096 * 12 * 2
    peak, peak_src = peaks()
    roof = None
    if integ_n:
        achieved = alg_bytes / (integ_ms / integ_n / 1000.0) / 1e9
        # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (not measurable live)
        traffic, traffic_src = None, None
        for name in (f"r02_k_integrate_traffic_{a.config}.json", "r01_k_integrate_traffic.json" if a.config == "c2" else ""):
            tp = ROOT / "profiles" / name
            if name and tp.exists():
                try:
                    tj = json.loads(tp.read_text())
                    traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("source")
                    break
                except Exception:
                    pass
        roof = {"kernel": "k_integrate (TSDF voxel update)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": integ_ms / integ_n,
                "launches_timed": integ_n, "updated_blocks_per_launch": upd_per_launch}
    gpu_time_ms = sum(v[0] for v in ktimes.values())

    # ---- e2e arm: host (pinned) buffers through the C ABI, copies inside the timed region -------------------
    e2e = [run(resident=False) for _ in range(R)]
    e2e_vals = sorted(world * frames / (r["ms"] / 1000.0) for r in e2e)
    emed = sorted(e2e, key=lambda r: r["ms"])[len(e2e) // 2]
    e2e_value = world * frames / (emed["ms"] / 1000.0)

    # ---- per-call latency of the single-frame surfaces (untimed region, rank-local) -------------------------
    latency = None
    if n_lat:
        hp.tsdf.Reset(); hp.tsdf.integrate(data.depth[0], data.poses[0], data.bgr[0])
        latency_block(hp, data, 1, min(4, n_lat))                       # warm the single-frame shapes
        latency = latency_block(hp, data, 1 + (W + K) * B, n_lat)

    sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()      # every rank leaves the group here; rank 0 still has the (CPU-only) baseline to time
    if rank != 0:
        return
    agg = med["agg"]
    line = {"metric": metric_name(a), "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32+f32",
            "data": (f"TUM RGB-D sequence {data.stream}: first {nframes} associated frames, ground-truth poses" if a.dataset else "synthetic"),
            "config": workload_config(a, {"l2": "256 MiB device buffer rewritten at the start of every step (inside the timed region); every step reads new frames",
                                          "timing": f"torch.cuda.Event pair around K steps, barrier+synchronize both sides, max over ranks; median of {R} timed passes "
                                                    "(each: map reset, seed scan, W warm-up steps, K timed steps); library calls synchronise their own streams before returning",
                                          "threads": "host threads = the reference's thread roles, each with its own library handle/CUDA stream: frame construction (two threads on alternate batches with the native driver; ORB extraction, "
                                                     f"batches of {B} frames: a throughput construct -- the reference's operator() takes one frame, see `latency`), Tracking (2x "
                                                     "SearchByProjection per frame), LocalMapping (SearchForTriangulation per frame), PointCloudMapping (TSDF per frame); consecutive "
                                                     "steps overlap as a software pipeline, the timed region ends when all stages have drained; the searches' queries are prepared "
                                                     "before the timed region (caller-side work, SURVEY.md §8d), which a live tracker would do between the calls",
                                          "driver": "plvs_pipeline_run (C++ stage threads over the C ABI)" if a.driver == "native" else "Python stage threads over the ctypes mirror",
                                          "host_affinity": numa}),
            "value_passes": [round(v, 1) for v in vals],
            "roofline": roof,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(emed["h2d"] / K), "d2h_bytes_per_step": int(emed["d2h"] / K),
                    "ms_per_step": emed["ms"] / K, "passes": [round(v, 1) for v in e2e_vals],
                    "bytes": "counted by the library at every copy it issues and every result written into mapped host memory (plvs_io_bytes)"},
            "gpu_launches": int(round(launches_per_step * K)), "clocks": med["clocks"],
            "kernel_ms_per_step": {k: round(v[0] / K, 4) for k, v in ktimes.items()}, "gpu_busy_frac": gpu_time_ms / t_ms,
            "per_step": {"keypoints": agg.get("keypoints", 0) / K, "matches": agg.get("matches", 0) / K,
                         "tsdf_blocks_visited_per_scan": vis_per_launch, "tsdf_blocks_updated_per_scan": upd_per_launch,
                         "match_rounds_last_call": hp.match_rounds()},
            "stage_busy_ms_per_step": {k[5:-2]: round(v / K * 1e3, 3) for k, v in agg.items() if k.startswith("busy_")},
            "wall_s": [round(med["wall"], 4), round(emed["wall"], 4)],
            "per_rank_ms_per_step": {"value": [round(x / K, 4) for x in med["per_rank_ms"]], "e2e": [round(x / K, 4) for x in emed["per_rank_ms"]]}}
    if latency:
        line["latency"] = latency
    if not a.no_cpu_baseline and world == 1:
        import cv2
        cv2.setNumThreads(1)
        cp = CpuPath(a, 1)
        cp.frame(0, tsdf="both", timed=False)
        for f in range(1, a.cpu_frames + 1):
            cp.frame(f, tsdf="both" if f == 1 else "port")
        pf = cp.per_frame()
        tsdf_s = pf.get("tsdf_ref_s", pf["tsdf_port_s"])
        fps = 1.0 / (pf["extract_s"] + pf["match_s"] + tsdf_s)
        front = 1.0 / (pf["extract_s"] + pf["match_s"])
        port = 1.0 / (pf["extract_s"] + pf["match_s"] + pf["tsdf_port_s"])
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "reference" if cp.have_ref else "port",
                                "front_end_frames_per_s": front, "port_value": port,
                                "sample": f"{a.cpu_frames} frames of stream 0 (after a 1-frame map seed), faithful mode: 1 thread; extract + match = "
                                          + ("the reference's own ORBextractor.cc / ORBmatcher.cc (oracle/_ref), TSDF = its own open_chisel on 1 frame "
                                             "('port_value' / tsdf_port_s = the restated TSDF)" if cp.have_ref else "the C restatement") +
                                          f"; seconds/frame {json.dumps({k: round(v, 4) for k, v in pf.items()})}; host has {os.cpu_count()} cpus"}
        line["speedup_vs_cpu_baseline"] = {"e2e_vs_whole_path": e2e_value / fps, "e2e_vs_front_end_only": e2e_value / front, "e2e_vs_restated_tsdf_path": e2e_value / port,
                                           "note": "same run, same box; the front-end figure is what north_star's >=30x target is about"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)
