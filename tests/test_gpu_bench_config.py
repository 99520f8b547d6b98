"""-m gpu: parity at the configurations bench.py measures (VERDICT r1 "weak #1"): the exact C2 geometry over consecutive scans of one map,
bench.py's unit of work (HotPath.step and the threaded HotPath.run_stream) stage by stage, and the C3 geometry at full depth range.

The bodies take their sizes as arguments so tests/test_emulated_kernels.py can replay them, reduced, on the CPU execution model; the GPU tests
run them at BASELINE.json's sizes.  Tolerance (north_star): chunk keys identical, sdf / weight within 1e-4 (the kernels follow the oracle's
operation order, so the arrays are compared for bit equality first and the tolerance only decides the verdict), colours and every integer
(keypoints, descriptors, assignments, counters) exactly equal."""
import os
import numpy as np
import pytest

from plvs_b200 import synth, tsdf as T
from oracle import orb as O, match as OM, tsdf as OT

TOL = 1e-4
THREADS = os.cpu_count() or 8


def _tsdf_equal(g, o, what, stats=True):
    gk, gs, gw, gc = g.download()
    ok, os_, ow, oc = o.download()
    assert gk.shape == ok.shape and np.array_equal(gk, ok), f"{what}: chunk key sets differ: {len(gk)} vs {len(ok)}"
    assert np.abs(gw - ow).max() <= TOL, what
    known = ow > 0
    assert np.abs(gs[known] - os_[known]).max() <= TOL, what
    assert np.array_equal(gs[~known], os_[~known]), what
    assert np.array_equal(gc, oc), what
    if stats:
        so, sg = o.stats(), g.stats()
        for f in ("n_blocks", "n_range", "n_updated", "n_new"):
            assert so[f] == sg[f], (what, f, so, sg)
    return len(gk), bool(np.array_equal(gs.view(np.uint32), os_.view(np.uint32)) and np.array_equal(gw.view(np.uint32), ow.view(np.uint32)))


def _impl_depth_scan_sequence(w, h, voxel, far, nscans, max_blocks, clip=None, device_input=False):
    """a24/a25 at bench geometry: `nscans` consecutive colour scans with carving on ONE map, compared after every scan (pool growth, the cull of
    k_classify_b over existing chunks, the 64x64-tile path for chunks around the camera, the dynamic work counter)"""
    K = synth.intrinsics(w, h)
    p = T.default_params(voxel_resolution=voxel, use_carving=1, near_plane=0.1, far_plane=far, max_blocks=max_blocks, use_color=1)
    g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    o = OT.Map(p, threads=THREADS); o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    exact = True
    for f in range(nscans):
        d = synth.depth_frame(f, w, h)
        if clip is not None:
            d[d > clip] = 0.0
        c = synth.bgr_frame(f, w, h)
        g.integrate(d, synth.pose(f), c); o.integrate(d, synth.pose(f), c)
        n, bits = _tsdf_equal(g, o, f"scan {f}")
        exact = exact and bits
    s = g.stats()
    return n, s, exact


def _impl_hot_path(w, h, nfeatures, voxel, far, batch, nsteps, max_blocks, threaded, native=True):
    """bench.py's step: batch extraction -> SearchByProjection(Cur,Last) + SearchByProjection(F,map) on the device-resident frame ->
    SearchForTriangulation -> colour depth-scan integration, compared with the oracle stage by stage (keypoints + descriptors per frame, the
    assignment arrays of every search, the TSDF map at the end of every step).  `threaded` additionally runs the same frames through the 4-thread
    pipeline (HotPath.run_stream, host buffers and device-resident inputs) and compares its totals and final map; `native` does the same through
    plvs_pipeline_run, the C++ stream driver bench.py times (`threaded="host"`: host buffers only -- the CPU model has no device tensors)."""
    from plvs_b200.pipeline import StreamData, HotPath
    from plvs_b200.matcher import Frame
    n = 1 + batch * nsteps
    d = StreamData(n, w, h, stream=0, pinned=False)
    hp = HotPath(d, nfeatures=nfeatures, voxel=voxel, far=far, max_blocks=max_blocks, batch=batch)
    hp.prepare()
    p = T.default_params(voxel_resolution=voxel, use_carving=1, near_plane=0.1, far_plane=far, max_blocks=max_blocks, use_color=1)
    o = OT.Map(p, threads=THREADS); o.set_camera(d.K["fx"], d.K["fy"], d.K["cx"], d.K["cy"], d.w, d.h)
    # stage 1: the frames prepare() extracted in batches == the oracle's extraction, frame by frame
    want_kp = 0
    for f in range(n):
        kp, desc, _, _ = O.extract_port(d.gray[f], nfeatures)
        assert np.array_equal(kp, hp.frames[f].keys) and np.array_equal(desc, hp.frames[f].desc), f"frame {f}: extraction differs"
        want_kp += len(kp)
    # stages 2+3 per frame, through the same calls HotPath._track_batch makes (device-resident current frame)
    sf, s2 = hp.ex.mvScaleFactor, hp.ex.mvLevelSigma2
    want_matches = 0
    hp.tsdf.Reset()
    hp.tsdf.integrate(d.depth[0], d.poses[0], d.bgr[0]); o.integrate(d.depth[0], d.poses[0], d.bgr[0])
    for s in range(nsteps):
        f0 = 1 + s * batch
        mono, kps, descs = hp.ex.extract_batch(d.gray[f0:f0 + batch])
        for b in range(batch):
            f = f0 + b
            q = hp.prepared[f]
            assert np.array_equal(kps[b], hp.frames[f].keys) and np.array_equal(descs[b], hp.frames[f].desc)
            dv = hp.ex.device_result(b)
            cur_ref, last = hp.frames[f], hp.frames[f - 1]
            cur = Frame(None, None, d.w, d.h, sf, s2, uright=cur_ref.uright, bf=d.K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, 0, dv.cache_key))
            n1, a1 = hp.m_track.SearchByProjectionLast(cur, q["ql"], 15.0)
            on1, oa1 = OM.search_by_projection_last(cur_ref, q["ql"], 15.0, False, False, True)
            assert n1 == on1 and np.array_equal(a1, oa1), f"frame {f}: SearchByProjection(Cur,Last) differs"
            claimed = (a1 >= 0).astype(np.uint8)
            n2, a2 = hp.m_track.SearchByProjectionMap(cur, q["qm"], 3.0, claimed=claimed, nnratio=0.8)
            on2, oa2 = OM.search_by_projection_map(cur_ref, q["qm"], 3.0, 0.8, claimed=claimed)
            assert n2 == on2 and np.array_equal(a2, oa2), f"frame {f}: SearchByProjection(F,map) differs"
            # the same two searches on the grid the extractor built at frame construction (what HotPath.step / plvs_pipeline_run hand over)
            assert dv.grid_cell_start and dv.grid_sorted
            curg = Frame(None, None, d.w, d.h, sf, s2, uright=cur_ref.uright, bf=d.K["bf"],
                         device_ptrs=(dv.n, dv.keys, dv.desc, 0, 0, dv.grid_cell_start, dv.grid_sorted))
            g1, ga1 = hp.m_map.SearchByProjectionLast(curg, q["ql"], 15.0)
            g2, ga2 = hp.m_map.SearchByProjectionMap(curg, q["qm"], 3.0, claimed=claimed, nnratio=0.8)
            assert (g1, g2) == (on1, on2) and np.array_equal(ga1, oa1) and np.array_equal(ga2, oa2), f"frame {f}: searches on the frame-construction grid differ"
            assert hp.m_map.last_stats()[1] == 2, "a search handed a grid must not build one"
            n3, m12 = hp.m_tri.SearchForTriangulation(Frame(kps[b], descs[b], d.w, d.h, sf, s2, uright=cur_ref.uright, bf=d.K["bf"]), last,
                                                      q["fv1"], q["fv2"], q["has1"], q["has2"], q["F12"], q["ep"], False, False)
            on3, om12 = OM.search_for_triangulation(cur_ref, last, q["fv1"], q["fv2"], q["has1"], q["has2"], q["F12"], q["ep"], False, False, False)
            assert n3 == on3 and np.array_equal(m12, om12), f"frame {f}: SearchForTriangulation differs"
            want_matches += n1 + n2 + n3
            hp.tsdf.integrate(d.depth[f], d.poses[f], d.bgr[f]); o.integrate(d.depth[f], d.poses[f], d.bgr[f])
        _tsdf_equal(hp.tsdf, o, f"step {s}")
    assert want_matches > 50 * nsteps * batch // 8
    # the unit bench.py calls, sequential form: totals and the final map
    hp.tsdf.Reset(); hp.tsdf.integrate(d.depth[0], d.poses[0], d.bgr[0])
    got = {}
    for s in range(nsteps):
        for k, v in hp.step(1 + s * batch, batch, resident=False, concurrent=False).items():
            got[k] = got.get(k, 0) + v
    assert got["keypoints"] == want_kp - len(hp.frames[0].keys) and got["matches"] == want_matches, (got, want_kp, want_matches)
    nblk, _ = _tsdf_equal(hp.tsdf, o, "HotPath.step", stats=False)
    # the native stream driver (plvs_pipeline_run: the stage threads in C++; PLVS_PIPELINE_SERIAL=1 runs its stages on one thread)
    if native:
        for resident in ((False,) if native == "host" else (False, True)):
            if resident:
                hp.upload_inputs()
            hp.tsdf.Reset(); hp.tsdf.integrate(d.depth[0], d.poses[0], d.bgr[0])
            agg = hp.run_stream_native(1, nsteps, resident)
            assert agg["keypoints"] == got["keypoints"] and agg["matches"] == want_matches, ("native", resident, agg, got)
            _tsdf_equal(hp.tsdf, o, f"plvs_pipeline_run resident={resident}", stats=False)
    if threaded:
        for resident in ((False,) if threaded == "host" else (False, True)):
            if resident:
                hp.upload_inputs()
            hp.tsdf.Reset(); hp.tsdf.integrate(d.depth[0], d.poses[0], d.bgr[0])
            agg = hp.run_stream(1, nsteps, resident)
            assert agg["keypoints"] == got["keypoints"] and agg["matches"] == want_matches, (resident, agg, got)
            _tsdf_equal(hp.tsdf, o, f"HotPath.run_stream resident={resident}", stats=False)
    return nblk, want_matches


@pytest.mark.gpu
def test_c2_bench_geometry_ten_scans(gpu):
    """BASELINE configs[1] exactly as bench.py integrates it: 640x480, 1 cm voxels, planes 0.1-5 m, colour + carving, 10 consecutive scans"""
    n, s, exact = _impl_depth_scan_sequence(640, 480, 0.01, 5.0, 10, 49152)
    assert n > 1500 and s["n_range"] > 40000 and s["n_updated"] > 1000 and s["n_candidates"] < 0.2 * s["n_range"], s
    assert exact, "within 1e-4 but not bit-identical (expected: the kernels follow the oracle's operation order)"


@pytest.mark.gpu
def test_c2_bench_geometry_device_input(gpu):
    """same geometry through the device-resident input path bench.py's `value` arm uses (plvs_tsdf_integrate_depth with device pointers)"""
    import torch
    from plvs_b200 import _lib
    import ctypes as C
    w, h = 640, 480
    K = synth.intrinsics(w, h)
    p = T.default_params(voxel_resolution=0.01, use_carving=1, near_plane=0.1, far_plane=5.0, max_blocks=49152, use_color=1)
    g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    o = OT.Map(p, threads=THREADS); o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    dev = torch.device("cuda", 0)
    for f in range(3):
        d, c = synth.depth_frame(f, w, h), synth.bgr_frame(f, w, h)
        td, tc = torch.from_numpy(d).to(dev), torch.from_numpy(c).to(dev)
        torch.cuda.synchronize(dev)
        pose = np.ascontiguousarray(synth.pose(f), np.float32).reshape(12)
        rc = g._lib.plvs_tsdf_integrate_depth(g._h, C.c_void_p(td.data_ptr()), w, h, C.c_void_p(tc.data_ptr()), w * 3, 3,
                                              pose.ctypes.data_as(C.c_void_p), T.SCAN_COLOR, 1)
        _lib.check(rc, "plvs_tsdf_integrate_depth")
        o.integrate(d, synth.pose(f), c)
        _tsdf_equal(g, o, f"device-input scan {f}")


@pytest.mark.gpu
def test_hot_path_step_at_bench_config(gpu):
    """bench.py's configuration (VGA, 2000 features, batch 8, 1 cm, 0.1-5 m): two steps = 17 frames, every stage against the oracle, then the
    threaded pipeline with host buffers and with device-resident inputs"""
    nblk, nm = _impl_hot_path(640, 480, 2000, 0.01, 5.0, 8, 2, 49152, threaded=True)
    assert nblk > 1500 and nm > 10000


@pytest.mark.gpu
def test_c3_full_range_two_scans(gpu):
    """BASELINE configs[2] geometry: 1920x1080, 5 mm voxels, planes 0.1-5 m (no depth clipping), two consecutive colour scans with carving"""
    n, s, exact = _impl_depth_scan_sequence(1920, 1080, 0.005, 5.0, 2, 200000)
    assert n > 5000 and s["n_range"] > 300000, s
    assert exact


@pytest.mark.gpu
def test_orb_1080p_against_every_arm(gpu):
    """1920x1080 / 4000 features: CUDA == C restatement == cv2-driven arm == the reference's own ORBextractor.cc (when oracle/_ref travelled)"""
    from plvs_b200.orb import ORBextractor
    img = synth.gray_frame(2, 1920, 1080)
    ex = ORBextractor(4000, 1.2, 8, 20, 7)
    mono, kp, desc = ex(img)
    okp, odesc, omono, _ = O.extract_port(img, 4000)
    assert mono == omono and np.array_equal(kp, okp) and np.array_equal(desc, odesc)
    ckp, cdesc, cmono, _ = O.extract_cv2(img, 4000, angle_impl="c")
    assert mono == cmono and np.array_equal(kp, ckp) and np.array_equal(desc, cdesc)
    if O.ref_available():
        rkp, rdesc, rmono = O.RefExtractor(4000)(img)
        assert mono == rmono and np.array_equal(kp, rkp) and np.array_equal(desc, rdesc)
