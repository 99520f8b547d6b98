"""-m gpu: the widened rows: UndistortKeyPoints, the resident SearchLocalPoints chain, 16-bit depth in front of the
TSDF, SearchForInitialization, mesh read-out, the DBoW2 transform, keyframe ids and the goldens recorded from the reference's own code.
They first ran on a B200 at the end of round 1 (all passed); they are ordinary strict tests now.  Each body still runs in a child interpreter
with a time limit: a kernel that hangs or host code that crashes ends that child (and is reported as this test's failure), not the session."""
import os
import pathlib
import subprocess
import sys
import cv2
import numpy as np
import pytest

from plvs_b200 import synth
from plvs_b200.orb import ORBextractor
from oracle import orb as O

pytestmark = pytest.mark.gpu


def _impl_undistort_keypoints_on_device():
    """Frame::UndistortKeyPoints (§8f rank 2): device == oracle == the real cv2.undistortPoints, bit for bit, TUM1 / EuRoC / rational models"""
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    mono, kp, desc = ex(synth.gray_frame(3))
    xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float32)
    for K4, dist in (((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
                     ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
                     ((500.0, 500.0, 320.0, 240.0), (0.1, -0.2, 0.001, -0.002, 0.05, 0.01, -0.02, 0.003))):
        un, dptr = ex.UndistortKeyPoints(K4, np.array(dist, np.float32))
        want = O.undistort_points(xy, K4, np.array(dist, np.float32))
        Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
        cvw = cv2.undistortPoints(xy.reshape(-1, 1, 2), Km, np.array(dist, np.float32), None, Km).reshape(-1, 2)
        assert np.array_equal(want.view(np.uint32), cvw.view(np.uint32))
        assert np.array_equal(un["x"].view(np.uint32), want[:, 0].view(np.uint32)) and np.array_equal(un["y"].view(np.uint32), want[:, 1].view(np.uint32))
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(un[f], kp[f])
        assert dptr
    un, _ = ex.UndistortKeyPoints((517.3, 516.5, 318.6, 255.3), np.zeros(5, np.float32))
    assert np.array_equal(un, kp)


def _impl_search_local_points_resident():
    """Tracking::SearchLocalPoints without a host round trip: plvs_match_in_frustum leaves the in-view queries on the device (compacted in order)
    and plvs_match_projection_map_resident searches with them; result == oracle isInFrustum -> host compaction -> oracle SearchByProjection"""
    from plvs_b200 import scenario
    from plvs_b200.matcher import ORBmatcher
    from oracle import match as OM
    w, h = 640, 480
    K = synth.intrinsics(w, h)
    tab = O.Tables(2000)
    fr = []
    for f in (10, 11):
        kp, desc, _, _ = O.extract_port(synth.gray_frame(f), 2000)
        x = scenario.make_frame(kp, desc, synth.depth_frame(f), K, tab.scale); fr.append(x)
    last, cur = fr
    Tl, Tc = synth.pose(10), synth.pose(11)
    ok = last.depth_at_kp > 0
    Pw = scenario.backproject(last.keys[ok], last.depth_at_kp[ok], K, Tl)
    n = len(Pw)
    pts = np.zeros(n, OM.MAP_POINT)
    pts["xw"] = Pw
    Ow = np.asarray(Tl, np.float64).reshape(3, 4)[:, 3]
    v = Pw.astype(np.float64) - Ow; d = np.linalg.norm(v, axis=1)
    pts["normal"] = (v / d[:, None]).astype(np.float32)
    lvl = last.keys["octave"][ok]
    pts["max_dist"] = (d * tab.scale[lvl]).astype(np.float32); pts["min_dist"] = (pts["max_dist"] / tab.scale[7]).astype(np.float32)
    rng = np.random.default_rng(1)
    pts["flags"] = (rng.random(n) < 0.9).astype(np.uint32); pts["desc"] = last.desc[ok]
    frm = OM.make_frustum(Tc, K, (0.0, 0.0, float(w), float(h)), K["bf"], 0.5, 1.2, 8)
    claimed = (rng.random(cur.n) < 0.2).astype(np.uint8)
    m = ORBmatcher(0.8, True)
    nin, q, iv = m.InFrustum(frm, pts)
    nm, assign = m.SearchByProjectionMapResident(cur, 3.0, claimed=claimed)
    on, oq, oiv = OM.in_frustum(frm, pts)
    assert nin == on and np.array_equal(iv, oiv) and nin > 500
    src = np.nonzero(oiv)[0]
    onm, oassign = OM.search_by_projection_map(cur, oq[src], 3.0, 0.8, claimed=claimed)
    want = np.where(oassign >= 0, src[np.maximum(oassign, 0)], -1)
    assert nm == onm and np.array_equal(assign, want) and nm > 200
    # the ordinary entry point with the same (downloaded) queries agrees as well
    nm2, assign2 = m.SearchByProjectionMap(cur, q[src], 3.0, claimed=claimed)
    assert nm2 == onm and np.array_equal(assign2, oassign)


def _impl_tsdf_from_raw_u16_depth():
    """§8f rank 2: `mImDepth.convertTo(CV_32F, mDepthMapFactor)` (src/Tracking.cc:1812-1813) on the device: the map built from the raw 16-bit
    image (TUM factor 5000, row padding) is the map built from the host-converted float image, bit for bit, with and without colour"""
    from plvs_b200 import tsdf as T
    from oracle import tsdf as OT
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    factor = np.float32(1.0) / np.float32(5000.0)                     # Tracking's `mDepthMapFactor = 1.0f / mDepthMapFactor`
    for color in (0, 1):
        p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=color)
        a, b = T.ChiselServer(p), T.ChiselServer(p)
        for g in (a, b):
            g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        for f in (0, 1, 2, 3, 4):
            padded = np.zeros((h, w + 6), np.uint16)
            padded[:, :w] = np.clip(synth.depth_frame(f, w, h) * 5000.0, 0, 65535).astype(np.uint16)
            d16 = padded[:, :w]                                       # a view with a row stride of 2*(w+6) bytes
            c = synth.bgr_frame(f, w, h) if color else None
            a.integrate_u16(d16, float(factor), synth.pose(f), c)
            b.integrate(OT.depth_u16_to_f32(d16, factor), synth.pose(f), c)
        ka, sa, wa, ca = a.download(); kb, sb, wb, cb = b.download()
        assert len(ka) > 50 and np.array_equal(ka, kb)
        assert np.array_equal(sa.view(np.uint32), sb.view(np.uint32)) and np.array_equal(wa.view(np.uint32), wb.view(np.uint32)) and np.array_equal(ca, cb)


def _impl_search_for_initialization():
    """ORBmatcher::SearchForInitialization (§8f rank 1, the last overload): device == oracle (itself pinned to the reference's compiled function) for
    the matches, the count and the updated vbPrevMatched, two calls in a row as Tracking::MonocularInitialization makes them; the mono
    initialisation extractor asks for 5x the features, so ~2000 level-0 keypoints compete"""
    from plvs_b200 import scenario
    from plvs_b200.matcher import ORBmatcher
    from oracle import match as OM
    K = synth.intrinsics(640, 480)
    tab = O.Tables(5000)
    fr = []
    for f in (10, 11, 14):
        kp, desc, _, _ = O.extract_port(synth.gray_frame(f), 5000)
        fr.append(scenario.make_frame(kp, desc, synth.depth_frame(f), K, tab.scale))
    f1, f2, f3 = fr
    for window, ratio, check in ((100, 0.9, True), (100, 0.9, False), (30, 0.7, True), (400, 1.0, True)):
        m = ORBmatcher(ratio, check)
        prev = np.stack([f1.keys["x"], f1.keys["y"]], 1)
        n, a, p = m.SearchForInitialization(f1, f2, prev, window)
        on, oa, op = OM.search_for_initialization(f1, f2, prev, window, ratio, check)
        assert n == on and np.array_equal(a, oa) and np.array_equal(p.view(np.uint32), op.view(np.uint32))
        assert n > 100
        n, a, p = m.SearchForInitialization(f1, f3, p, window)
        on, oa, op = OM.search_for_initialization(f1, f3, op, window, ratio, check)
        assert n == on and np.array_equal(a, oa) and np.array_equal(p.view(np.uint32), op.view(np.uint32))
    e = scenario.make_frame(f2.keys[:0], f2.desc[:0], synth.depth_frame(11), K, tab.scale)
    n, a, p = ORBmatcher(0.9, True).SearchForInitialization(f1, e, prev, 100)
    assert n == 0 and (a == -1).all() and np.array_equal(p, prev)


def _impl_mesh_read_out():
    """§8f rank 3: ChunkManager::RecomputeMesh for every chunk on the device (marching cubes in the reference's vertex order, colours through
    InterpolateColor as written, gradient normals) == the oracle, which is pinned bit-exactly to the compiled open_chisel
    (tests/test_oracle_vs_reference_mesh.py).  The voxel arrays of the two maps are identical in these sequences, so the meshes must be too."""
    from plvs_b200 import tsdf as T
    from oracle import tsdf as OT

    def pair(w, h, **kw):
        K = synth.intrinsics(w, h)
        p = T.default_params(**kw)
        g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        o = OT.Map(p, threads=8); o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        return g, o

    def same(g, o):
        nm, nv = g.UpdateMesh()
        gk, gc, gV, gN, gC = g.GetMeshes()
        ok, oc, oV, oN, oC = o.extract_mesh()
        assert nm == len(ok) and nv == len(oV)
        assert np.array_equal(gk, ok) and np.array_equal(gc, oc)
        for a, b in ((gV, oV), (gN, oN), (gC, oC)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        return nv

    for color, res in ((1, 0.04), (0, 0.04), (1, 0.5)):            # 0.5 m voxels: InterpolateColor's trilinear branch is live
        g, o = pair(160, 120, voxel_resolution=res, use_carving=1, near_plane=0.1, far_plane=4.0 if res < 0.1 else 6.0, max_blocks=8192, use_color=color)
        for f in (0, 1, 2, 6):
            d = synth.depth_frame(f, 160, 120); c = synth.bgr_frame(f, 160, 120) if color else None
            g.integrate(d, synth.pose(f), c); o.integrate(d, synth.pose(f), c)
            assert np.array_equal(g.download()[1].view(np.uint32), o.download()[1].view(np.uint32))
            nv = same(g, o)
        assert nv > (10000 if res < 0.1 else 100)
    xyz, rgb, nrm = g.GetPointCloud()
    assert len(xyz) == nv and rgb.dtype == np.uint8 and rgb.max() > 50
    g.Reset()
    assert g.UpdateMesh() == (0, 0) and len(g.GetMeshes()[2]) == 0


def _impl_bow_transform(tmp_path):
    """§8f rank 4: ORBVocabulary::transform (Frame::ComputeBoW) on the device == the oracle (pinned to the compiled DBoW2): words, weights, nodes,
    BowVector bit for bit, FeatureVector; then SearchByBoW fed with the device-resident FeatureVectors gives the same matches as with host ones"""
    from oracle import bow as OB, match as OM
    from plvs_b200.bow import ORBVocabulary
    from plvs_b200 import scenario
    from plvs_b200.matcher import ORBmatcher
    K = synth.intrinsics(640, 480)
    tab = O.Tables(2000)
    fr = []
    for f in (10, 11):
        kp, desc, _, _ = O.extract_port(synth.gray_frame(f), 2000)
        fr.append(scenario.make_frame(kp, desc, synth.depth_frame(f), K, tab.scale))
    for k, L, levelsup, zero in ((10, 4, 2, 0.0), (10, 3, 1, 0.2), (6, 4, 4, 0.1)):
        path = tmp_path / ("voc%d%d.txt" % (k, L))
        OB.write_vocabulary(path, k, L, seed=k + L, zero_weight_fraction=zero)
        ov = OB.Vocabulary(path)
        gv = ORBVocabulary()
        assert gv.loadFromTextFile(path) and gv.size() == ov.size() == k ** L
        outs = []
        for f in fr:
            g, o = gv.transform(f.desc, levelsup), ov.transform(f.desc, levelsup)
            for name in o:
                a, b = g[name], o[name]
                assert np.array_equal(a.view(np.uint64), b.view(np.uint64)) if a.dtype == np.float64 else np.array_equal(a, b), name
            outs.append(o)
    # the matcher consumes what the transform produced (node ids two levels above the leaves of the last vocabulary: root -> one node; use the first)
    path = tmp_path / "voc_match.txt"
    OB.write_vocabulary(path, 10, 3, seed=5)
    gv = ORBVocabulary(); assert gv.loadFromTextFile(path)
    ov = OB.Vocabulary(path)
    fk, ff = (ov.transform(f.desc, 1) for f in fr)
    fvK = (fk["fv_nodes"], fk["fv_offsets"], fk["fv_features"]); fvF = (ff["fv_nodes"], ff["fv_offsets"], ff["fv_features"])
    has = np.ones(fr[0].n, np.uint8)
    m = ORBmatcher(0.7, True)
    n1, m1 = m.SearchByBoW(fr[0], fr[1], fvK, fvF, has)
    n2, m2 = OM.search_by_bow(fr[0], fr[1], fvK, fvF, has, 0.7, True)
    assert n1 == n2 and np.array_equal(m1, m2) and n1 > 50


# ---- the tests proper: one isolated child per body -----------------------------------------------------------------------------------
ROOT = pathlib.Path(__file__).resolve().parent.parent
CHILD_TIMEOUT_S = 120
_timed_out = []              # once a child had to be killed, the remaining bodies are not started: the whole file then costs one time limit, not seven


def _isolated(name, *args):
    if _timed_out:
        pytest.fail("not started: %s did not finish within its time limit" % _timed_out[0])
    code = "import sys; sys.path.insert(0, %r); import pathlib; from tests import test_gpu_widened as t; t.%s(%s)" % (
        str(ROOT), name, ", ".join("pathlib.Path(%r)" % str(a) for a in args))
    try:
        r = subprocess.run([sys.executable, "-c", code], cwd=str(ROOT), capture_output=True, text=True, timeout=CHILD_TIMEOUT_S,
                           env=dict(os.environ, PYTHONFAULTHANDLER="1"))
    except subprocess.TimeoutExpired:
        _timed_out.append(name)
        pytest.fail("%s did not finish within %d s (child killed)" % (name, CHILD_TIMEOUT_S))
    assert r.returncode == 0, "%s: exit %d\n%s\n%s" % (name, r.returncode, r.stdout[-2000:], r.stderr[-4000:])


def test_undistort_keypoints_on_device(gpu):
    _isolated("_impl_undistort_keypoints_on_device")


def test_search_local_points_resident(gpu):
    _isolated("_impl_search_local_points_resident")


def test_tsdf_from_raw_u16_depth(gpu):
    _isolated("_impl_tsdf_from_raw_u16_depth")


def test_search_for_initialization(gpu):
    _isolated("_impl_search_for_initialization")


def test_mesh_read_out(gpu):
    _isolated("_impl_mesh_read_out")


def test_bow_transform(gpu, tmp_path):
    _isolated("_impl_bow_transform", tmp_path)


def _impl_reference_goldens(tmp_path):
    """the three widened rows against vectors recorded from the reference's own code (tests/golden/extras_ref.npz): no oracle in between"""
    import tests.test_extras_golden as G
    from plvs_b200 import tsdf as T
    from plvs_b200.bow import ORBVocabulary
    from plvs_b200.matcher import ORBmatcher

    def search(f1, f2, prev, window, ratio, check):
        return ORBmatcher(ratio, check).SearchForInitialization(f1, f2, prev, window)

    def meshes(g):
        g.UpdateMesh()
        return g.GetMeshes()

    def vocabulary(path):
        v = ORBVocabulary(); assert v.loadFromTextFile(path)
        return v
    G.check_initialization(search)
    G.check_meshes(lambda kw, frames, color: G.X.mesh_map(T.ChiselServer, kw, frames, color), meshes)
    G.check_bow(vocabulary, tmp_path)


def test_reference_goldens(gpu, tmp_path):
    _isolated("_impl_reference_goldens", tmp_path)


def _impl_keyframe_ids():
    """DistVoxel::kfid / Mesh::kfids through the point-cloud route: product == oracle (pinned to the compiled open_chisel), with one id per cloud and
    with per-point ids, carving resets included; the voxel arrays stay what plvs_tsdf_integrate_cloud gives"""
    from plvs_b200 import scenario, tsdf as T
    from oracle import tsdf as OT
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    p = T.default_params(voxel_resolution=0.04, use_carving=1, carving_dist=0.05, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=1)
    g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    o = OT.Map(p, threads=8); o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    rng = np.random.default_rng(7)
    for step, f in enumerate((0, 1, 3, 6)):
        d = synth.depth_frame(f, w, h) + np.float32(0.15 * step)
        c = synth.bgr_frame(f, w, h)
        xyz, rgb = scenario.cloud_from_depth(d, c, K, step=2)
        kfids = None if step % 2 == 0 else rng.integers(1, 50, len(xyz)).astype(np.uint32)
        g.integrate_cloud_kf(xyz, rgb, synth.pose(f), d, kfids=kfids, kfid=100 + f); o.integrate_cloud_kf(xyz, rgb, synth.pose(f), d, kfids=kfids, kfid=100 + f)
        gk, gs, gw, gc = g.download(); ok, os_, ow, oc = o.download()
        assert np.array_equal(gk, ok) and np.array_equal(gs.view(np.uint32), os_.view(np.uint32)) and np.array_equal(gw.view(np.uint32), ow.view(np.uint32))
        assert np.array_equal(g.download_kfid(), o.download_kfid())
        nm, nv = g.UpdateMesh()
        assert nv == len(o.extract_mesh()[2]) and np.array_equal(g.mesh_kfids(), o.mesh_kfids(nv))
    assert (o.download_kfid() > 0).sum() > 1000
    g.Reset()
    assert g.download_kfid().size == 0


def test_keyframe_ids(gpu):
    _isolated("_impl_keyframe_ids")


def _rt(rng, scale):
    a = rng.normal(size=3); a /= np.linalg.norm(a)
    th = scale * rng.uniform(0.2, 1.0)
    Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    return np.concatenate([R, (scale * rng.normal(size=3))[:, None]], 1).astype(np.float32)


def _impl_deform():
    """ChunkManager::Deform (§8f rank 3): product == oracle (pinned bit-exactly to the compiled open_chisel with the reference's own chunk order,
    tests/test_oracle_vs_reference_deform.py) for the default key order and for an explicit visiting order; voxels of unlisted keyframes vanish;
    collisions fold in order; the meshes of the last UpdateMesh move along; the map keeps working afterwards; a pool too small leaves it untouched"""
    from plvs_b200 import scenario, tsdf as T
    from oracle import tsdf as OT
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    rng = np.random.default_rng(3)
    for color, scale, explicit in ((1, 0.02, False), (1, 0.3, True), (0, 0.1, False)):
        p = T.default_params(voxel_resolution=0.04, use_carving=1, carving_dist=0.05, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=color)
        g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        o = OT.Map(p, threads=8); o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        for i, f in enumerate((0, 2, 5, 9)):
            d = synth.depth_frame(f, w, h)
            xyz, rgb = scenario.cloud_from_depth(d, synth.bgr_frame(f, w, h) if color else None, K, step=2)
            g.integrate_cloud_kf(xyz, rgb, synth.pose(f), d, kfid=101 + i); o.integrate_cloud_kf(xyz, rgb, synth.pose(f), d, kfid=101 + i)
        nm, nv = g.UpdateMesh()
        mk, mc, V0, N0, C0 = g.GetMeshes(); vk = g.mesh_kfids()
        kf = np.array([101, 102, 104], np.uint32)
        Rt = np.stack([_rt(rng, scale) for _ in kf])
        order = o.download()[0][::-1].copy() if explicit else None
        g.Deform(kf, Rt, order); o.deform(kf, Rt, order=order)
        gk, gs, gw, gc = g.download(); ok, os_, ow, oc = o.download()
        assert np.array_equal(gk, ok) and len(gk) > 30
        assert np.array_equal(gw.view(np.uint32), ow.view(np.uint32)) and np.array_equal(gs.view(np.uint32), os_.view(np.uint32)) and np.array_equal(gc, oc)
        assert np.array_equal(g.download_kfid(), o.download_kfid())
        assert g.stats()["n_blocks"] == len(gk)
        # the stored meshes moved with their keyframes (vertices of keyframe 103 stay)
        _, _, V1, N1, _ = g.GetMeshes()
        for j, k in enumerate(kf):
            sel = vk == k
            R, t = Rt[j][:, :3], Rt[j][:, 3]
            want = np.stack([R[a, 0] * V0[sel, 0] + (R[a, 1] * V0[sel, 1] + R[a, 2] * V0[sel, 2]) + t[a] for a in range(3)], 1).astype(np.float32)
            assert np.array_equal(V1[sel].view(np.uint32), want.view(np.uint32)) and sel.sum() > 10
        assert np.array_equal(V1[vk == 103], V0[vk == 103])
        # the deformed map keeps working
        d = synth.depth_frame(11, w, h)
        xyz, rgb = scenario.cloud_from_depth(d, synth.bgr_frame(11, w, h) if color else None, K, step=3)
        g.integrate_cloud_kf(xyz, rgb, synth.pose(11), d, kfid=105); o.integrate_cloud_kf(xyz, rgb, synth.pose(11), d, kfid=105)
        gk, gs, gw, gc = g.download(); ok, os_, ow, oc = o.download()
        assert np.array_equal(gk, ok) and np.array_equal(gw.view(np.uint32), ow.view(np.uint32)) and np.array_equal(gs.view(np.uint32), os_.view(np.uint32))
    # a pool that cannot hold both maps: refused, map unchanged
    p = T.default_params(voxel_resolution=0.04, use_carving=0, near_plane=0.1, far_plane=4.0, max_blocks=0, use_color=1)
    d = synth.depth_frame(0, w, h)
    xyz, rgb = scenario.cloud_from_depth(d, synth.bgr_frame(0, w, h), K, step=2)
    probe = T.ChiselServer(T.default_params(voxel_resolution=0.04, use_carving=0, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=1))
    probe.integrate_cloud_kf(xyz, rgb, synth.pose(0), None, kfid=1)
    nblk = probe.stats()["n_blocks"]
    p.max_blocks = nblk + 8
    small = T.ChiselServer(p)
    small.integrate_cloud_kf(xyz, rgb, synth.pose(0), None, kfid=1)
    before = small.download()
    with pytest.raises(Exception):
        small.Deform(np.array([1], np.uint32), _rt(rng, 0.5)[None])
    after = small.download()
    assert all(np.array_equal(a, b) for a, b in zip(before, after)) and small.stats()["n_blocks"] == nblk


def test_deform(gpu):
    _isolated("_impl_deform")


def _impl_world_cloud():
    """Chisel::IntegrateWorldPointCloudWithNormals (§8f rank 3, the map-loading route): product == oracle (pinned bit-exactly to the compiled open_chisel)"""
    from plvs_b200 import scenario, tsdf as T
    from oracle import tsdf as OT
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    rng = np.random.default_rng(9)
    for color in (1, 0):
        p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=color)
        g = T.ChiselServer(p); o = OT.Map(p, threads=8)
        for f in (0, 4):
            d = synth.depth_frame(f, w, h)
            xyz, rgb = scenario.cloud_from_depth(d, synth.bgr_frame(f, w, h), K, step=2)
            Pw = (xyz @ synth.pose(f)[:, :3].T + synth.pose(f)[:, 3]).astype(np.float32)
            nrm = rng.normal(size=Pw.shape).astype(np.float32) * np.float32(0.2) + np.array([0, 0, -1], np.float32)
            nrm[::97] = 0
            kf = rng.integers(1, 9, len(Pw)).astype(np.uint32)
            Twc = np.eye(4, dtype=np.float32)[:3] if f == 0 else synth.pose(1)
            g.IntegrateWorldPointCloud(Pw, rgb if color else None, nrm, Twc, kfids=kf); o.integrate_world_cloud(Pw, rgb if color else None, nrm, Twc, kfids=kf)
            gk, gs, gw, gc = g.download(); ok, os_, ow, oc = o.download()
            assert np.array_equal(gk, ok) and len(gk) > 30
            assert np.array_equal(gw.view(np.uint32), ow.view(np.uint32)) and np.array_equal(gs.view(np.uint32), os_.view(np.uint32)) and np.array_equal(gc, oc)
            assert np.array_equal(g.download_kfid(), o.download_kfid())
        g.IntegrateWorldPointCloud(np.zeros((0, 3), np.float32), None, np.zeros((0, 3), np.float32), np.eye(4, dtype=np.float32)[:3])      # empty cloud: nothing happens
        assert g.stats()["n_blocks"] == len(gk)


def test_world_cloud(gpu):
    _isolated("_impl_world_cloud")


def _impl_line_knn2():
    """§8f rank 4, line features: LineMatcher::ComputeDescriptorMatches -- the 2-NN of the vendored multi-index hashing (tie order included) and the
    ratio test -- against the restatement, against the golden recorded from the compiled reference, and against that library itself where it travelled"""
    import pathlib
    from plvs_b200.matcher import LineMatcher
    from oracle import linematch as L
    from tests.linematch_cases import cases
    lm = LineMatcher(0.78)
    ties = 0
    for q, t, mask in cases(40, seed=1, nq_max=300, nt_max=400):
        nv, qi, ti, di, vi = lm.ComputeDescriptorMatches(q, t, mask)
        a = L.knn2(q, t, mask, 0.78)
        assert np.array_equal(qi, a[0]) and np.array_equal(ti, a[1]) and np.array_equal(di, a[2]) and np.array_equal(vi, a[3]) and nv == a[4]
        ties += int(np.sum(di[:, 0] == di[:, 1]))
    assert ties > 100
    g = np.load(pathlib.Path(__file__).parent / "golden" / "linematch_ref.npz")
    for i, (q, t, mask) in enumerate(cases(int(g["n_cases"]), seed=int(g["seed"]))):
        nv, qi, ti, di, vi = lm.ComputeDescriptorMatches(q, t, mask)
        assert np.array_equal(qi, g[f"qi{i}"]) and np.array_equal(ti, g[f"ti{i}"]) and np.array_equal(di, g[f"di{i}"]) and np.array_equal(vi, g[f"vi{i}"])
    try:
        R = L.RefLineMatcher()
    except Exception:
        R = None
    if R is not None:                 # oracle/_ref travels to the GPU box
        for q, t, mask in cases(8, seed=9, nq_max=500, nt_max=600):
            nv, qi, ti, di, vi = lm.ComputeDescriptorMatches(q, t, mask)
            b = R.knn2(q, t, mask, 0.78)
            assert np.array_equal(qi, b[0]) and np.array_equal(ti, b[1]) and np.array_equal(di, b[2]) and np.array_equal(vi, b[3]) and nv == b[4]
    # error behaviour: empty inputs and fewer train descriptors than k are refused, nothing is written
    import ctypes as C
    from plvs_b200 import _lib
    rows, nvv = C.c_int(7), C.c_int(7)
    z = np.zeros((4, 32), np.uint8); o = np.zeros(16, np.int32); f = np.zeros(16, np.float32); u = np.zeros(8, np.uint8)
    for nq_, nt_ in ((0, 4), (4, 0), (4, 1)):
        rc = lm._m._lib.plvs_line_knn2(lm._m._h, z.ctypes.data, nq_, z.ctypes.data, nt_, None, 0.78, o.ctypes.data, o.ctypes.data, f.ctypes.data, u.ctypes.data, C.byref(rows), C.byref(nvv))
        assert rc == -1 and rows.value == 0          # PLVS_EINVAL


def test_line_knn2(gpu):
    _impl_line_knn2()


def _impl_map_file_round_trip(tmp_path):
    """§8f rank 3, "PLY save / load": a map point cloud written like PointCloudMapChisel::SaveMap writes it (plvs_map_save_ply == the reference's WritePLY byte for
    byte, tests/test_map_ply.py), read back like PointCloudMap::LoadMap reads it, InvertColors, IntegrateWorldPointCloud with the identity == the oracle fed the
    original points"""
    from plvs_b200 import scenario, tsdf as T
    from oracle import tsdf as OT
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    rng = np.random.default_rng(4)
    d = synth.depth_frame(2, w, h)
    xyz, rgb = scenario.cloud_from_depth(d, synth.bgr_frame(2, w, h), K, step=2)            # rgb: floats in [0, 1], r g b
    Pw = (xyz @ synth.pose(2)[:, :3].T + synth.pose(2)[:, 3]).astype(np.float32)
    nrm = (rng.normal(size=Pw.shape).astype(np.float32) * np.float32(0.2) + np.array([0, 0, -1], np.float32)).astype(np.float32)
    kf = rng.integers(1, 9, len(Pw)).astype(np.uint32)
    rgb8 = np.clip(np.round(rgb * 255.0), 0, 255).astype(np.uint8)
    bgra = np.concatenate([rgb8[:, ::-1], np.full((len(Pw), 1), 255, np.uint8)], 1)          # PCL's memory order
    n3 = len(Pw) // 3 * 3
    f = tmp_path / "volumetric_map_out_0.ply"
    T.save_map_ply(f, Pw[:n3], bgra[:n3], nrm[:n3], np.zeros(n3, np.uint32), kf[:n3], is_mesh=True, binary=True)
    m = T.load_map_ply(f)
    assert m["fields"] == 31 and len(m["xyz"]) == n3
    # the binary file stores (b, g, r) under the names red, green, blue; InvertColors swaps red and blue back: r g b again
    rgb_loaded = m["rgb"][:, ::-1].astype(np.float32) * np.float32(1.0 / 255.0)
    assert np.array_equal(m["rgb"][:, ::-1], rgb8[:n3])
    p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=1)
    g = T.ChiselServer(p); o = OT.Map(p, threads=8)
    I = np.eye(4, dtype=np.float32)[:3]
    g.IntegrateWorldPointCloud(m["xyz"], rgb_loaded, m["normals"], I, kfids=m["kfid"])
    o.integrate_world_cloud(Pw[:n3], rgb8[:n3].astype(np.float32) * np.float32(1.0 / 255.0), nrm[:n3], I, kfids=kf[:n3])
    gk, gs, gw, gc = g.download(); ok, os_, ow, oc = o.download()
    assert np.array_equal(gk, ok) and len(gk) > 30
    assert np.array_equal(gw.view(np.uint32), ow.view(np.uint32)) and np.array_equal(gs.view(np.uint32), os_.view(np.uint32)) and np.array_equal(gc, oc)
    assert np.array_equal(g.download_kfid(), o.download_kfid())


def test_map_file_round_trip(gpu, tmp_path):
    _impl_map_file_round_trip(tmp_path)
