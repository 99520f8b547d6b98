"""Stream runner used by bench.py / tests: drives the three reference-facing surfaces (ORBextractor,
ORBmatcher, ChiselServer) over a synthetic RGB-D stream the way PLVS's threads do -- a *tracking* thread
(extract + the two SearchByProjection calls + SearchForTriangulation against the previous frame) and a
*dense-mapping* thread (TSDF integration), which the reference also runs concurrently
(src/System.cc:317-398: Tracking in the caller's thread, PointCloudMapping::Run in its own).

Caller-side work that is NOT part of the hot path (building map-point / last-frame queries from poses and
depth) is precomputed once by `prepare()` so the timed region contains only the hot path."""
import ctypes as C
import os
import threading
import time
import numpy as np

from . import _lib, synth, scenario, tsdf as T
from .matcher import ORBmatcher, Frame, featvec, featvec_struct
from .orb import ORBextractor, KP_DTYPE


class PinnedArray:
    """numpy view of cudaHostAlloc'ed memory (pinned, so H2D copies are true DMA)."""

    def __init__(self, shape, dtype):
        self.lib = _lib.load()
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = C.c_void_p()
        _lib.check(self.lib.plvs_host_alloc(C.byref(self.ptr), max(nbytes, 1)), "plvs_host_alloc")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.plvs_host_free(self.ptr)
            self.ptr = None


class StreamData:
    """Synthetic inputs of one camera stream: gray u8, depth f32, bgr u8 and poses for `n` frames."""

    def __init__(self, n, w, h, stream=0, pinned=True, workers=None):
        self.n, self.w, self.h, self.stream = n, w, h, stream
        self.K = synth.intrinsics(w, h)
        mk = (lambda s, d: PinnedArray(s, d)) if pinned else None
        self._pins = []
        def alloc(shape, dtype):
            if pinned:
                p = PinnedArray(shape, dtype); self._pins.append(p); return p.array
            return np.empty(shape, dtype)
        self.gray = alloc((n, h, w), np.uint8)
        self.depth = alloc((n, h, w), np.float32)
        self.bgr = alloc((n, h, w, 3), np.uint8)
        self.poses = np.zeros((n, 3, 4), np.float32)
        # rendering a frame is pure numpy (about 0.2 s at VGA, 1.7 s at 1080p): long or large streams are rendered by worker processes
        grays = None
        if workers is None:
            ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))       # ranks of a torchrun job share the host's cores
            workers = min(32, max(1, (os.cpu_count() or 1) // ranks)) if n * w * h >= 64 * 640 * 480 else 0
        if workers > 1:
            grays = synth.render_gray_parallel(n, w, h, stream, min(workers, n))       # None if the worker processes could not be used
        for f in range(n):
            self.gray[f] = grays[f] if grays is not None else synth.gray_frame(f, w, h, stream)
            self.depth[f] = synth.depth_frame(f, w, h, stream)
            self.bgr[f] = np.stack([self.gray[f], np.roll(self.gray[f], 3, 1), 255 - self.gray[f]], -1)
            self.poses[f] = synth.pose(f)

    def input_bytes_per_frame(self):
        return self.w * self.h * (1 + 4 + 3)

    @classmethod
    def from_tum(cls, root, n, pinned=True, rgb_flag=True, depth_factor=5000.0, K=None, max_dt=0.02):
        """The first `n` frames of a TUM RGB-D benchmark sequence directory -- BASELINE.json configs[0], what Examples_old/RGB-D/rgbd_tum.cc reads
        (LoadImages :199-230: an associations file `t_rgb rgb/x.png t_depth depth/x.png`; without one, rgb.txt and depth.txt are paired by nearest timestamp,
        as the benchmark's associate.py does).  Gray as Tracking::GrabImageRGBD makes it (src/Tracking.cc:1797-1810: cvtColor of the BGR data cv::imread returns
        with COLOR_RGB2GRAY when Camera.RGB is 1, as in TUM1.yaml), depth = raw 16-bit x 1 / DepthMapFactor (:1812-1813), colour = the image as read,
        poses = groundtruth.txt (camera-to-world, nearest timestamp) -- the TSDF stage needs a pose and there is no tracker here to supply one."""
        import pathlib
        import cv2
        root = pathlib.Path(root)

        def table(name):
            rows = []
            for ln in (root / name).read_text().splitlines():
                ln = ln.strip()
                if ln and not ln.startswith("#"):
                    rows.append(ln.replace(",", " ").split())
            return rows

        pairs = []
        for cand in ("associations.txt", "associate.txt", "association.txt"):
            if (root / cand).exists():
                pairs = [(float(r[0]), r[1], r[3]) for r in table(cand) if len(r) >= 4]
                break
        if not pairs:
            rgb = [(float(r[0]), r[1]) for r in table("rgb.txt")]; dep = [(float(r[0]), r[1]) for r in table("depth.txt")]
            td = np.array([t for t, _ in dep]); used = set()
            for t, f in rgb:
                j = int(np.argmin(np.abs(td - t)))
                if abs(td[j] - t) <= max_dt and j not in used:
                    used.add(j); pairs.append((t, f, dep[j][1]))
        if len(pairs) < n:
            raise ValueError(f"{root}: {len(pairs)} associated RGB-D pairs, {n} needed")
        gt = np.array([[float(x) for x in r[:8]] for r in table("groundtruth.txt")]) if (root / "groundtruth.txt").exists() else None
        first = cv2.imread(str(root / pairs[0][1]), cv2.IMREAD_COLOR)
        if first is None:
            raise ValueError(f"cannot read {root / pairs[0][1]}")
        h, w = first.shape[:2]
        self = cls.__new__(cls)
        self.n, self.w, self.h, self.stream = n, w, h, str(root.name)
        self.K = dict(K) if K else synth.intrinsics(w, h)              # TUM1.yaml's pinhole parameters scaled to the image size
        self._pins = []

        def alloc(shape, dtype):
            if pinned:
                p = PinnedArray(shape, dtype); self._pins.append(p); return p.array
            return np.empty(shape, dtype)
        self.gray = alloc((n, h, w), np.uint8); self.depth = alloc((n, h, w), np.float32); self.bgr = alloc((n, h, w, 3), np.uint8)
        self.poses = np.zeros((n, 3, 4), np.float32)
        for i, (t, frgb, fdep) in enumerate(pairs[:n]):
            img = cv2.imread(str(root / frgb), cv2.IMREAD_COLOR); d16 = cv2.imread(str(root / fdep), cv2.IMREAD_UNCHANGED)
            if img is None or d16 is None or img.shape[:2] != (h, w) or d16.shape[:2] != (h, w):
                raise ValueError(f"{root}: frame {i} ({frgb}, {fdep}) unreadable or of a different size")
            self.bgr[i] = img
            self.gray[i] = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY if rgb_flag else cv2.COLOR_BGR2GRAY)
            self.depth[i] = d16.astype(np.float32) * np.float32(1.0 / depth_factor)
            T = np.eye(4, dtype=np.float64)
            if gt is not None and len(gt):
                g = gt[int(np.argmin(np.abs(gt[:, 0] - t)))]
                qx, qy, qz, qw = g[4:8] / np.linalg.norm(g[4:8])
                T[:3, :3] = [[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                             [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                             [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]]
                T[:3, 3] = g[1:4]
            self.poses[i] = T[:3].astype(np.float32)
        return self


class HotPath:
    """extract + match + TSDF for one stream on one GPU."""

    def __init__(self, data, nfeatures=2000, voxel=0.01, far=5.0, max_blocks=32768, device=0, batch=8, use_color=True):
        self.d, self.batch, self.device = data, batch, device
        self.ex = ORBextractor(nfeatures, 1.2, 8, 20, 7, device=device)
        self.ex.set_frame_grid(data.w, data.h)      # frame construction ends with AssignFeaturesToGrid (src/Frame.cc:598)
        self.ex2 = None           # second extractor workspace, created by run_stream (batch k+1 extracts while batch k is matched)
        self.ex_more = []         # further extractor handles of the native driver (two frame-construction threads)
        self.m_track = ORBmatcher(0.9, True, device=device)      # TrackWithMotionModel (src/Tracking.cc:3593)
        self.m_map = ORBmatcher(0.8, True, device=device)        # SearchLocalPoints (src/Tracking.cc:4477)
        self.m_tri = ORBmatcher(0.6, False, device=device)       # CreateNewMapFeatures (src/LocalMapping.cc:537)
        p = T.default_params(voxel_resolution=voxel, use_carving=1, near_plane=0.1, far_plane=far, max_blocks=max_blocks, use_color=int(use_color))
        self.tsdf = T.ChiselServer(p, device=device)
        K = data.K
        self.tsdf.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], data.w, data.h)
        self.use_color = use_color
        self.prepared = None
        self._pins = []
        self.dev = None       # device-resident copies of the inputs (the `value` arm)

    # ---- untimed: caller-side query construction -------------------------------------------------------
    def prepare(self):
        d, K = self.d, self.d.K
        sf, s2 = self.ex.GetScaleFactors(), self.ex.GetScaleSigmaSquares()
        frames = []
        for f0 in range(0, d.n, self.batch):
            mono, kps, descs = self.ex.extract_batch(d.gray[f0:f0 + self.batch])
            for b in range(len(kps)):
                fr = scenario.make_frame(kps[b].copy(), descs[b].copy(), d.depth[f0 + b], K, sf)
                fr.level_sigma2 = s2
                frames.append(fr)
        prep = []
        for f in range(d.n):
            if f == 0:
                prep.append(None); continue
            last, cur = frames[f - 1], frames[f]
            ql, _ = scenario.last_queries(last, cur, K, d.poses[f - 1], d.poses[f])
            qm, _ = scenario.map_queries(last, cur, K, d.poses[f - 1], d.poses[f], seed=f)
            fv_last, fv_cur = featvec(scenario.node_ids(last.desc)), featvec(scenario.node_ids(cur.desc))
            F12, ep = scenario.fundamental(K, d.poses[f], d.poses[f - 1])
            prep.append(dict(ql=ql, qm=qm, fv1=fv_cur, fv2=fv_last, F12=F12, ep=ep,
                             has1=np.zeros(cur.n, np.uint8), has2=np.zeros(last.n, np.uint8)))
        # the caller keeps its query records in pinned memory (one slab per kind), so the per-call H2D copy is a plain DMA
        for key in ("ql", "qm"):
            items = [p[key] for p in prep if p is not None]
            if not items:
                continue
            slab = PinnedArray((sum(len(a) for a in items),), items[0].dtype)
            self._pins.append(slab)
            o = 0
            for p in prep:
                if p is None:
                    continue
                a = p[key]
                slab.array[o:o + len(a)] = a
                p[key] = slab.array[o:o + len(a)]
                o += len(a)
        self.frames, self.prepared = frames, prep
        self.tsdf.Reset()

    def upload_inputs(self):
        """device-resident arm: inputs live in HBM before the timed region (torch owns the allocations)."""
        import torch
        dev = torch.device("cuda", self.device)
        self.dev = dict(gray=torch.from_numpy(self.d.gray).to(dev), depth=torch.from_numpy(self.d.depth).to(dev),
                        bgr=torch.from_numpy(self.d.bgr).to(dev))
        torch.cuda.synchronize(dev)

    # ---- timed: the hot path ---------------------------------------------------------------------------
    def _track_batch(self, f0, nb, resident, out):
        d = self.d
        if resident:
            g = self.dev["gray"]
            mono, kps, descs = self.ex.extract_batch((g[f0].data_ptr(), nb, d.h, d.w, d.w, d.w * d.h))
        else:
            mono, kps, descs = self.ex.extract_batch(d.gray[f0:f0 + nb])
        sf, s2 = self.ex.mvScaleFactor, self.ex.mvLevelSigma2
        nm = 0
        for b in range(nb):
            f = f0 + b
            p = self.prepared[f]
            if p is None:
                continue
            # descriptors/keypoints of the current frame stay on the device for the searches
            dv = self.ex.device_result(b)
            cur_ref = self.frames[f]
            cur = Frame(None, None, d.w, d.h, sf, s2, uright=self.frames[f].uright, bf=d.K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, 0, dv.cache_key, dv.grid_cell_start, dv.grid_sorted))
            n1, a1 = self.m_track.SearchByProjectionLast(cur, p["ql"], 15.0)
            claimed = (a1 >= 0).astype(np.uint8)
            # same tracking-thread workspace => the feature grid of the frame is built once for both searches
            n2, a2 = self.m_track.SearchByProjectionMap(cur, p["qm"], 3.0, claimed=claimed, nnratio=0.8)
            last = self.frames[f - 1]
            n3, m12 = self.m_tri.SearchForTriangulation(Frame(kps[b], descs[b], d.w, d.h, sf, s2, uright=cur_ref.uright, bf=d.K["bf"]), last,
                                                        p["fv1"], p["fv2"], p["has1"], p["has2"], p["F12"], p["ep"], False, False)
            nm += n1 + n2 + n3
        out["matches"] = out.get("matches", 0) + nm
        out["keypoints"] = out.get("keypoints", 0) + sum(len(k) for k in kps)

    def _map_batch(self, f0, nb, resident, out):
        d = self.d
        for b in range(nb):
            f = f0 + b
            if resident:
                self._integrate_device(f)
            else:
                self.tsdf.integrate(d.depth[f], d.poses[f], d.bgr[f] if self.use_color else None)
        # no per-frame statistics here: asking for them would wait for the scan (the library keeps running totals on the device)

    def _integrate_device(self, f):
        d = self.d
        lib, h = self.tsdf._lib, self.tsdf._h
        pose = np.ascontiguousarray(d.poses[f], np.float32).reshape(12)
        bgr = self.dev["bgr"][f] if self.use_color else None
        rc = lib.plvs_tsdf_integrate_depth(h, C.c_void_p(self.dev["depth"][f].data_ptr()), d.w, d.h,
                                           C.c_void_p(bgr.data_ptr()) if bgr is not None else None, d.w * 3, 3 if bgr is not None else 0,
                                           pose.ctypes.data_as(C.c_void_p), T.SCAN_COLOR if bgr is not None else T.SCAN, 1)
        _lib.check(rc, "plvs_tsdf_integrate_depth")

    def step(self, f0, nb, resident=False, concurrent=True):
        """one step = one batch of `nb` consecutive frames through extract+match (tracking thread) and TSDF
        (dense-mapping thread)."""
        out = {}
        if concurrent:
            err = []
            def mapper():
                try:
                    self._map_batch(f0, nb, resident, out)
                except Exception as e:      # surface worker failures in the caller
                    err.append(e)
            t = threading.Thread(target=mapper)
            t.start()
            self._track_batch(f0, nb, resident, out)
            t.join()
            if err:
                raise err[0]
        else:
            self._track_batch(f0, nb, resident, out)
            self._map_batch(f0, nb, resident, out)
        return out

    # ---- timed: the same hot path as a 4-stage software pipeline over consecutive steps --------------------
    def run_stream(self, f0, nsteps, resident=False, per_step=None):
        """`nsteps` steps of `batch` frames starting at frame f0, with the reference's thread structure made explicit:
        frame construction / ORB extraction, Tracking's two SearchByProjection calls, LocalMapping's
        SearchForTriangulation and PointCloudMapping's TSDF integration run in four threads, each on its own library
        handle (= its own CUDA stream), connected by queues.  Every call is still synchronous for its caller; the
        overlap is between calls of different stages (batch k+1 extracts while batch k is matched, ...).
        Returns the summed statistics once everything has drained."""
        import queue
        d, nb = self.d, self.batch
        if self.ex2 is None:
            self.ex2 = ORBextractor(self.ex.nfeatures, 1.2, 8, 20, 7, device=self.device)
            self.ex2.set_frame_grid(self.d.w, self.d.h)
        exs = (self.ex, self.ex2)
        free = [threading.Semaphore(1), threading.Semaphore(1)]      # extractor workspace i may be overwritten
        q_track, q_tri = queue.Queue(), queue.Queue()
        out, err = {}, []
        lock = threading.Lock()
        sf, s2 = self.ex.mvScaleFactor, self.ex.mvLevelSigma2

        def add(k, v):
            with lock:
                out[k] = out.get(k, 0) + v

        def guard(fn):
            def run():
                try:
                    fn()
                except Exception as e:          # surface worker failures in the caller
                    err.append(e)
                    q_track.put(None); q_tri.put(None)
                    for sem in free:
                        sem.release()
            return run

        def extract_stage():
            for s in range(nsteps):
                if err:
                    break
                i = s & 1
                free[i].acquire()
                t0 = time.perf_counter()
                if per_step is not None:
                    per_step(s)
                fs = f0 + s * nb
                if resident:
                    g = self.dev["gray"]
                    mono, kps, descs = exs[i].extract_batch((g[fs].data_ptr(), nb, d.h, d.w, d.w, d.w * d.h))
                else:
                    mono, kps, descs = exs[i].extract_batch(d.gray[fs:fs + nb])
                add("keypoints", sum(len(k) for k in kps))
                add("busy_extract_s", time.perf_counter() - t0)
                q_track.put((s, i, kps, descs))
            q_track.put(None)

        def track_stage():
            while True:
                it = q_track.get()
                if it is None:
                    break
                s, i, kps, descs = it
                q_tri.put((s, kps, descs))
                t0 = time.perf_counter()
                nm = 0
                for b in range(nb):
                    f = f0 + s * nb + b
                    p = self.prepared[f]
                    if p is None:
                        continue
                    dv = exs[i].device_result(b)
                    cur = Frame(None, None, d.w, d.h, sf, s2, uright=self.frames[f].uright, bf=d.K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, 0, dv.cache_key, dv.grid_cell_start, dv.grid_sorted))
                    n1, a1 = self.m_track.SearchByProjectionLast(cur, p["ql"], 15.0)
                    claimed = (a1 >= 0).astype(np.uint8)
                    n2, a2 = self.m_track.SearchByProjectionMap(cur, p["qm"], 3.0, claimed=claimed, nnratio=0.8)
                    nm += n1 + n2
                free[i].release()
                add("matches", nm)
                add("busy_track_s", time.perf_counter() - t0)
            q_tri.put(None)

        def tri_stage():
            while True:
                it = q_tri.get()
                if it is None:
                    break
                s, kps, descs = it
                t0 = time.perf_counter()
                nm = 0
                for b in range(nb):
                    f = f0 + s * nb + b
                    p = self.prepared[f]
                    if p is None:
                        continue
                    cur_ref, last = self.frames[f], self.frames[f - 1]
                    n3, m12 = self.m_tri.SearchForTriangulation(Frame(kps[b], descs[b], d.w, d.h, sf, s2, uright=cur_ref.uright, bf=d.K["bf"]), last,
                                                                p["fv1"], p["fv2"], p["has1"], p["has2"], p["F12"], p["ep"], False, False)
                    nm += n3
                add("matches", nm)
                add("busy_tri_s", time.perf_counter() - t0)

        def map_stage():
            t0 = time.perf_counter()
            for s in range(nsteps):
                if err:
                    break
                self._map_batch(f0 + s * nb, nb, resident, out)
            self.tsdf.stats()                      # waits for the last scan
            add("busy_map_s", time.perf_counter() - t0)

        ts = [threading.Thread(target=guard(fn)) for fn in (extract_stage, track_stage, tri_stage, map_stage)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        return out

    # ---- timed: the same pipeline on native threads (plvs_b200/csrc/pipeline.cu) -----------------------------------------------
    def _native_job(self):
        """flat per-frame records for plvs_pipeline_run: pointers into the arrays prepare() built (kept alive by self)"""
        if getattr(self, "_job_frames", None) is not None:
            return self._job_frames
        d = self.d
        recs = (_lib.PipelineFrame * d.n)()
        keep = []
        for f in range(d.n):
            p = self.prepared[f]
            r = recs[f]
            if p is None:
                r.valid = 0
                continue
            cur, last = self.frames[f], self.frames[f - 1]
            r.valid = 1
            r.n_ql, r.n_qm = len(p["ql"]), len(p["qm"])
            r.ql, r.qm = p["ql"].ctypes.data, p["qm"].ctypes.data
            r.uright = cur.uright.ctypes.data
            r.fv_cur, r.fv_last = featvec_struct(p["fv1"]), featvec_struct(p["fv2"])
            r.has_cur, r.has_last = p["has1"].ctypes.data, p["has2"].ctypes.data
            F12 = np.ascontiguousarray(p["F12"], np.float32).reshape(9); ep = np.ascontiguousarray(p["ep"], np.float32)
            for i in range(9):
                r.F12[i] = float(F12[i])
            r.ep[0], r.ep[1] = float(ep[0]), float(ep[1])
            r.last = last.view()
            keep.append((cur, last, p))
        self._job_keep, self._job_frames = keep, recs
        return recs

    def run_stream_native(self, f0, nsteps, resident=False, flush=None):
        """HotPath.run_stream with the four stage threads in C++ (plvs_pipeline_run): the same calls in the same order per stage, without the
        interpreter between them.  `flush` = (device pointer, bytes) of a buffer rewritten at the start of every step, or None."""
        d = self.d
        if self.ex2 is None:
            self.ex2 = ORBextractor(self.ex.nfeatures, 1.2, 8, 20, 7, device=self.device)
            self.ex2.set_frame_grid(self.d.w, self.d.h)
        recs = self._native_job()
        job = _lib.PipelineJob()
        job.device, job.width, job.height, job.batch, job.n_steps, job.first_frame = self.device, d.w, d.h, self.batch, nsteps, f0
        job.cap = self.ex._cap
        job.inputs_on_device = int(resident)
        if resident:
            job.gray, job.depth = self.dev["gray"].data_ptr(), self.dev["depth"].data_ptr()
            job.bgr = self.dev["bgr"].data_ptr() if self.use_color else None
        else:
            job.gray, job.depth = d.gray.ctypes.data, d.depth.ctypes.data
            job.bgr = d.bgr.ctypes.data if self.use_color else None
        poses = np.ascontiguousarray(d.poses, np.float32)
        job.poses = poses.ctypes.data
        job.frames = C.addressof(recs)
        sf, s2 = self.ex.mvScaleFactor, self.ex.mvLevelSigma2
        job.view_template = Frame(np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), d.w, d.h, sf, s2, bf=d.K["bf"]).view()
        job.th_last, job.th_map, job.nnratio_map = 15.0, 3.0, 0.8
        if flush is not None:
            job.flush_buf, job.flush_bytes = flush
        # extractor handles: 2 = one frame-construction thread (batch k+1 extracts while batch k is matched); 4 (default) = two threads on alternate batches
        n_ex = max(2, int(os.environ.get("PLVS_PIPELINE_EXTRACTORS", "4")) & ~1)
        while len(self.ex_more) < n_ex - 2:
            e = ORBextractor(self.ex.nfeatures, 1.2, 8, 20, 7, device=self.device)
            e.set_frame_grid(self.d.w, self.d.h)
            self.ex_more.append(e)
        exs = (C.c_void_p * n_ex)(self.ex._h.value, self.ex2._h.value, *[e._h.value for e in self.ex_more[:n_ex - 2]])
        st = _lib.PipelineStats()
        rc = self.ex._lib.plvs_pipeline_run(exs, n_ex, self.m_track._h, self.m_tri._h, self.tsdf._h, C.byref(job), C.byref(st))
        _lib.check(rc, "plvs_pipeline_run")
        return dict(keypoints=st.keypoints, matches=st.matches, busy_extract_s=st.busy_extract_s, busy_track_s=st.busy_track_s,
                    busy_tri_s=st.busy_tri_s, busy_map_s=st.busy_map_s, wall_s=st.wall_s)

    def match_rounds(self):
        lib = self.ex._lib
        out = []
        r, k = C.c_int(), C.c_int()
        lib.plvs_match_last_stats(self.m_track._h, C.byref(r), C.byref(k))
        out.append(r.value)
        return out

    def search_h2d_bytes(self, f):
        """host->device bytes the three searches of frame f copy besides the images: query records, the frame's mvuRight (twice),
        the pre-claim array, and for SearchForTriangulation both host-side frame views, their flattened FeatureVectors,
        map-point flags and F12"""
        p = self.prepared[f]
        if p is None:
            return 0
        cur, last = self.frames[f], self.frames[f - 1]
        b = p["ql"].nbytes + p["qm"].nbytes + 2 * cur.n * 4 + cur.n
        for fr, fv, has in ((cur, p["fv1"], p["has1"]), (last, p["fv2"], p["has2"])):
            b += fr.n * (28 + 32 + 4) + has.nbytes + sum(np.asarray(a).nbytes for a in fv)
        return b + 9 * 4

    def launches_per_frame(self):
        """kernel launches of the last batch / frame, as counted by the library"""
        lib = self.ex._lib
        n = self.ex.last_stats()["kernel_launches"] / max(self.batch, 1)
        r, k = C.c_int(), C.c_int()
        for m in (self.m_track, self.m_tri):
            lib.plvs_match_last_stats(m._h, C.byref(r), C.byref(k))
            n += k.value * (2 if m is self.m_track else 1)
        n += self.tsdf.stats()["kernel_launches"]
        return n
