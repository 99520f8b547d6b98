#!/usr/bin/env python3
"""bench.py -- frames/s of the PLVS per-frame hot path (ORB extract + Hamming match + Chisel TSDF) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of `--batch` consecutive 640x480 RGB-D frames of a synthetic stream (BASELINE.json
configs[1]: ORB nFeatures=2000 + Chisel TSDF 1 cm voxels); every frame is extracted, matched against the
previous frame (SearchByProjection Cur<-Last), against its local map points (SearchByProjection F<-map),
triangulation-searched against the previous frame and integrated into the TSDF.  One camera stream per
rank (weak scaling, no collective on the data path).  Prints ONE JSON line (rank 0).
"""
import argparse
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

METRIC = "frames/sec (extract+match+TSDF) 640x480 RGB-D"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--nfeatures", type=int, default=2000)
    ap.add_argument("--voxel", type=float, default=0.01)
    ap.add_argument("--far", type=float, default=5.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=2)
    return ap.parse_args()


def workload_config(a, extra=None):
    c = {"workload": f"synthetic {a.width}x{a.height} RGB-D stream, ORB nFeatures={a.nfeatures} (8 levels, 1.2, FAST 20/7) + "
                     f"SearchByProjection(Cur,Last) th=15 + SearchByProjection(F,map) th=3 + SearchForTriangulation + "
                     f"Chisel TSDF {a.voxel * 100:g} cm voxels (colour depth-scan, carving on, planes 0.1-{a.far:g} m), every frame integrated",
         "frames_per_step": a.batch, "streams_per_gpu": 1, "parallelism": f"1 camera stream per GPU x{a.gpus}"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's CPU path) timed on this box's host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_run(a, frames, threads, generous, tsdf_ref_frames=0):
    """times the CPU path on frames [1, frames] of stream 0; returns (frames/s, per-stage seconds per frame).
    Extract = cv2 primitives + C++ restatement, match = C++ restatement.  TSDF = the restatement (OpenMP over chunks when
    `generous`), and additionally -- for the first `tsdf_ref_frames` timed frames, when oracle/_ref/libchisel_ref.so
    exists -- the REFERENCE's own open_chisel code (single-threaded, as Chisel.h:91 is), whose mean seconds/frame then
    replaces the restatement's in the total ("tsdf_s"; the restatement's time stays in "tsdf_port_s")."""
    import cv2
    from oracle import orb as O, match as OM, tsdf as OT
    from plvs_b200 import synth, scenario, tsdf as T
    from plvs_b200.matcher import featvec
    cv2.setNumThreads(threads if generous else 1)
    K = synth.intrinsics(a.width, a.height)
    tab = O.Tables(a.nfeatures)
    p = T.default_params(voxel_resolution=a.voxel, use_carving=1, near_plane=0.1, far_plane=a.far, max_blocks=1 << 20, use_color=1)
    omap = OT.Map(p, threads=threads if generous else 1)
    omap.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], a.width, a.height)
    rmap = None
    if tsdf_ref_frames > 0 and OT.ref_available():
        rmap = OT.RefMap(p)
        rmap.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], a.width, a.height)
    t_ext = t_match = t_tsdf = t_ref = 0.0
    n_ref = 0
    prev = None
    for f in range(frames + 1):
        img = synth.gray_frame(f, a.width, a.height); depth = synth.depth_frame(f, a.width, a.height); bgr = synth.bgr_frame(f, a.width, a.height)
        t0 = time.perf_counter()
        kp, desc, mono, _ = O.extract_cv2(img, a.nfeatures, angle_impl="c")
        t1 = time.perf_counter()
        cur = scenario.make_frame(kp, desc, depth, K, tab.scale); cur.level_sigma2 = tab.sigma2
        if prev is not None:
            ql, _ = scenario.last_queries(prev, cur, K, synth.pose(f - 1), synth.pose(f))
            qm, _ = scenario.map_queries(prev, cur, K, synth.pose(f - 1), synth.pose(f), seed=f)
            fv1, fv2 = featvec(scenario.node_ids(cur.desc)), featvec(scenario.node_ids(prev.desc))
            F12, ep = scenario.fundamental(K, synth.pose(f), synth.pose(f - 1))
            t2 = time.perf_counter()
            n1, a1 = OM.search_by_projection_last(cur, ql, 15.0)
            OM.search_by_projection_map(cur, qm, 3.0, 0.8, claimed=(a1 >= 0).astype(np.uint8))
            OM.search_for_triangulation(cur, prev, fv1, fv2, np.zeros(cur.n, np.uint8), np.zeros(prev.n, np.uint8), F12, ep, check_ori=False)
            t3 = time.perf_counter()
            omap.integrate(depth, synth.pose(f), bgr)
            t4 = time.perf_counter()
            t_ext += t1 - t0; t_match += t3 - t2; t_tsdf += t4 - t3
            if rmap is not None and n_ref < tsdf_ref_frames:
                with _quiet_stdout():
                    rmap.integrate(depth, synth.pose(f), bgr)
                t_ref += time.perf_counter() - t4; n_ref += 1
        else:
            omap.integrate(depth, synth.pose(f), bgr)        # map initialisation, untimed
            if rmap is not None:
                with _quiet_stdout():
                    rmap.integrate(depth, synth.pose(f), bgr)
        prev = cur
    stages = dict(extract_s=t_ext / frames, match_s=t_match / frames, tsdf_s=t_tsdf / frames)
    if n_ref:
        stages["tsdf_port_s"] = stages["tsdf_s"]
        stages["tsdf_s"] = t_ref / n_ref
        stages["tsdf_ref_frames"] = n_ref
    tot = stages["extract_s"] + stages["match_s"] + stages["tsdf_s"]
    return 1.0 / tot, stages


import contextlib


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


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import tsdf as OT
    threads = os.cpu_count() or 1
    have_ref = OT.ref_available()
    # each step = one frame of the same workload (a bounded sample: the CPU TSDF alone takes seconds per frame).  The
    # warm-up frames run the multi-threaded restatement only; the timed frames run extract/match through the restatement with
    # all host threads and the TSDF through the reference's own open_chisel code (oracle/_ref) for the first REF_FRAMES frames.
    REF_FRAMES = 2
    w = max(0, min(a.warmup, 1))
    if w:
        cpu_reference_run(a, w, threads, True)
    t0 = time.perf_counter()
    fps, stages = cpu_reference_run(a, a.steps, threads, True, tsdf_ref_frames=REF_FRAMES if have_ref else 0)
    wall = time.perf_counter() - t0
    port_fps = 1.0 / (stages["extract_s"] + stages["match_s"] + stages.get("tsdf_port_s", stages["tsdf_s"]))
    kind = "reference" if have_ref else "port"
    note = ("extract = cv2 primitives + C++ restatement (cv2.setNumThreads(n)), match = C++ restatement, TSDF = the reference's own open_chisel "
            "sources compiled into oracle/_ref (single-threaded: Chisel.h:91 has its parallel_for commented out); 'port_value' = same run with "
            "the restated TSDF spread over all host threads with OpenMP; rank 0 only") if have_ref else \
           "CPU oracle (cv2 primitives + C++ restatement of the reference), all host threads: cv2.setNumThreads(n), TSDF chunks over OpenMP; rank 0 only"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32+f32",
            "data": "synthetic", "config": workload_config(a, {"frames_per_step": 1, "note": note}),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "port_value": port_fps,
                             "sample": f"{a.steps} frames of the same stream (TSDF through oracle/_ref on the first {stages.get('tsdf_ref_frames', 0)}), "
                                       f"stage seconds/frame {json.dumps({k: round(v, 4) for k, v in stages.items()})}"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "wall_s": wall}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        """one nvidia-smi process for the whole bench, started BEFORE any timed region: its start-up (NVML initialisation,
        a few hundred ms during which driver calls of this process can stall) must not fall into a timed window"""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
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


def run_b200_arm(a):
    import torch
    sys.setswitchinterval(1e-4)          # 4 stage threads: hand the GIL over quickly when a library call returns
    import torch.distributed as dist
    from plvs_b200 import _lib
    from plvs_b200.pipeline import StreamData, HotPath

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL writes its version banner (NCCL_DEBUG=VERSION/INFO) to stdout by default; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    B, K, W = a.batch, a.steps, a.warmup
    nframes = 1 + (W + K) * B            # frame 0 only seeds the map / the "last frame"
    data = StreamData(nframes, a.width, a.height, stream=rank, pinned=True)
    hp = HotPath(data, a.nfeatures, a.voxel, a.far, max_blocks=49152, device=local, batch=B)
    hp.prepare()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    sampler = ClockSampler(local); sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run(resident, timed_profile=0):
        hp.tsdf.Reset()
        hp.tsdf.integrate(data.depth[0], data.poses[0], data.bgr[0])           # seed (untimed)
        hp.run_stream(1, W, resident)                                          # warm-up steps (untimed)
        lib.plvs_set_profiling(timed_profile)
        lib.plvs_tsdf_kernel_times(hp.tsdf._h, None, None, 1)
        for ex in (hp.ex, hp.ex2):
            if ex is not None:
                lib.plvs_orb_kernel_times(ex._h, None, None, 1)
        for m in (hp.m_track, hp.m_map, hp.m_tri):
            lib.plvs_match_kernel_times(m._h, None, None, 1)
        barrier()
        mark = sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        # K steps through the 4-stage pipeline (extract | projection searches | triangulation | TSDF); returns when drained
        agg = hp.run_stream(1 + W * B, K, resident, per_step=lambda s: flush.fill_(s & 0xff))   # L2 flush once per step
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.window(mark)
        t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        lib.plvs_set_profiling(0)
        return float(t_ms.item()), wall, clocks, agg

    import ctypes as C
    # ---- value arm: inputs resident in HBM ----------------------------------------------------------------
    hp.upload_inputs()
    # timed run: only the roofline kernel (k_integrate) carries events (2 per frame); the per-kernel table comes from a
    # second, untimed pass with all events on
    # order: the per-kernel profile pass runs first (its result is not a bench value; it also settles every data-dependent
    # buffer size and the clocks), then the timed passes
    run(resident=True, timed_profile=7)
    ktimes = {}
    ms = np.zeros(12, np.float32); cnt = np.zeros(12, np.int32)
    lib.plvs_tsdf_kernel_times(hp.tsdf._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
    for i, nme in enumerate(("tsdf.depth_tiles", "tsdf.classify", "tsdf.integrate", "tsdf.commit")):
        ktimes[nme] = (float(ms[i]), int(cnt[i]))
    om = np.zeros(12, np.float32); oc = np.zeros(12, np.int32)
    for ex in (hp.ex, hp.ex2):
        if ex is not None:
            lib.plvs_orb_kernel_times(ex._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
            om += ms; oc += cnt
    for i, nme in enumerate(("orb.pyramid", "orb.fast_cells", "orb.compact", "orb.blur", "orb.orient_describe", "orb.distribute")):
        ktimes[nme] = (float(om[i]), int(oc[i]))
    mm = np.zeros(12, np.float32); mc = np.zeros(12, np.int32)
    for m in (hp.m_track, hp.m_map, hp.m_tri):
        lib.plvs_match_kernel_times(m._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
        mm += ms; mc += cnt
    for i, nme in enumerate(("match.grid", "match.candidates", "match.resolve", "match.triangulate")):
        ktimes[nme] = (float(mm[i]), int(mc[i]))
    t_ms, wall, clocks, agg = run(resident=True, timed_profile=4 | 8)
    ms = np.zeros(12, np.float32); cnt = np.zeros(12, np.int32)
    lib.plvs_tsdf_kernel_times(hp.tsdf._h, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 0)
    integ_live = (float(ms[2]), int(cnt[2]))
    frames = K * B
    value = world * frames / (t_ms / 1000.0)
    launches_per_step = hp.launches_per_frame() * B

    # roofline of the dominant kernel: TSDF voxel update.  Algorithmic bytes per launch:
    # depth 4*W*H + colour 3*W*H + n_updated_blocks * 4096 voxels * 12 B (sdf f32 + weight f32 + rgba) * 2 (read+write)
    integ_ms, integ_n = integ_live
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
        tp = ROOT / "profiles" / "r01_k_integrate_traffic.json"
        if tp.exists():
            try:
                tj = json.loads(tp.read_text())
                traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("source")
            except Exception:
                pass
        roof = {"kernel": "k_integrate (TSDF voxel update)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": integ_ms / integ_n,
                "updated_blocks_per_launch": upd_per_launch}
    gpu_time_ms = sum(v[0] for v in ktimes.values())

    # ---- e2e arm: host (pinned) buffers through the public classes, copies inside the timed region ---------
    t2_ms, wall2, clocks2, agg2 = run(resident=False)
    e2e_value = world * frames / (t2_ms / 1000.0)
    # every host->device byte of a step: the images (gray u8 + depth f32 + bgr u8x3) and what the searches stage per frame
    h2d = B * data.input_bytes_per_frame() + int(sum(hp.search_h2d_bytes(f) for f in range(1 + W * B, 1 + (W + K) * B)) / K)
    d2h = int(B * (agg2.get("keypoints", 0) / max(frames, 1)) * 60 + B * 3 * a.nfeatures * 4)

    sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()      # every rank leaves the group here; rank 0 still has the (CPU-only) baseline to time
    if rank != 0:
        return
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32+f32", "data": "synthetic",
            "config": workload_config(a, {"l2": "256 MiB device buffer rewritten between steps (inside the timed region); every step reads new frames",
                                          "timing": "torch.cuda.Event pair around K steps, barrier+synchronize both sides, max over ranks; "
                                                    "library calls synchronise their own streams before returning",
                                          "threads": "4 host threads = the reference's thread roles, each with its own library handle/CUDA stream: frame construction (ORB extraction, "
                                                     "batches of 8 frames), Tracking (2x SearchByProjection per frame), LocalMapping (SearchForTriangulation per frame), "
                                                     "PointCloudMapping (TSDF per frame); consecutive steps overlap as a software pipeline, the timed region ends when all "
                                                     "stages have drained"}),
            "roofline": roof, "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": t2_ms / K},
            "gpu_launches": int(round(launches_per_step * K)), "clocks": clocks,
            "kernel_ms_per_step": {k: round(v[0] / K, 4) for k, v in ktimes.items()}, "gpu_busy_frac": gpu_time_ms / t_ms,
            "per_step": {"keypoints": agg.get("keypoints", 0) / K, "matches": agg.get("matches", 0) / K,
                         "tsdf_blocks_visited_per_scan": vis_per_launch, "tsdf_blocks_updated_per_scan": upd_per_launch,
                         "match_rounds_last_call": hp.match_rounds()},
            "stage_busy_ms_per_step": {k[5:-2]: round(v / K * 1e3, 3) for k, v in agg.items() if k.startswith("busy_")}, "wall_s": [wall, wall2]}
    if not a.no_cpu_baseline:
        from oracle import tsdf as OT
        have_ref = OT.ref_available()
        fps, stages = cpu_reference_run(a, a.cpu_frames, 1, False, tsdf_ref_frames=1 if have_ref else 0)
        port_fps = 1.0 / (stages["extract_s"] + stages["match_s"] + stages.get("tsdf_port_s", stages["tsdf_s"]))
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "reference" if have_ref else "port", "port_value": port_fps,
                                "sample": f"{a.cpu_frames} frames of stream 0 (after a 1-frame map seed), faithful mode: 1 thread "
                                          f"(cv2.setNumThreads(1), single-threaded chunk loop like Chisel.h:91); extract = real OpenCV primitives + "
                                          f"restated octree/descriptor code, match = restatement, TSDF = "
                                          + ("the reference's own open_chisel sources (oracle/_ref, 1 frame; 'port_value'/'tsdf_port_s' = the restated TSDF)"
                                             if have_ref else "restatement") +
                                          f"; seconds/frame {json.dumps({k: round(v, 4) for k, v in stages.items()})}; host has {os.cpu_count()} cpus"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)
