"""CPU: the restatements of ChunkManager::Deform and Chisel::IntegrateWorldPointCloudWithNormals against the reference's
own open_chisel compiled into oracle/_ref/libchisel_ref.so -- bit for bit (keys, sdf, weight, colour, keyframe ids).  Deform's result depends on
the order in which the reference's std::unordered_map hands out the chunks; the restatement takes that order as an input and gets the compiled
reference's own (exported by the harness before the call)."""
import numpy as np
import pytest

from plvs_b200 import synth, scenario
from oracle import tsdf as OT

pytestmark = pytest.mark.skipif(not OT.ref_available(), reason="oracle/_ref/libchisel_ref.so not built (no /root/reference here)")


def _pair(**kw):
    p = OT.default_params(**kw)
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    o = OT.Map(p, threads=4); r = OT.RefMap(p)
    for m in (o, r):
        m.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    return o, r, K, w, h


def _same(o, r):
    ok, os_, ow, oc = o.download(); rk, rs, rw, rc = r.download()
    assert np.array_equal(ok, rk), (len(ok), len(rk))
    assert np.array_equal(ow.view(np.uint32), rw.view(np.uint32)) and np.array_equal(os_.view(np.uint32), rs.view(np.uint32)) and np.array_equal(oc, rc)
    assert np.array_equal(o.download_kfid(), r.download_kfid())
    return len(ok)


def _rt(rng, scale):
    """a small rigid correction: rotation about a random axis (Rodrigues, float32) and a translation"""
    a = rng.normal(size=3); a /= np.linalg.norm(a)
    th = scale * rng.uniform(0.2, 1.0)
    Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    return np.concatenate([R, (scale * rng.normal(size=3))[:, None]], 1).astype(np.float32)


# colour maps only: the reference's point-cloud routes (the only writers of keyframe ids) touch the colour voxels of every chunk and crash on a map without colours
@pytest.mark.parametrize("color,scale", [(1, 0.02), (1, 0.15), (1, 0.6)])
def test_deform_equals_reference(color, scale):
    o, r, K, w, h = _pair(voxel_resolution=0.04, use_carving=1, carving_dist=0.05, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=color)
    rng = np.random.default_rng(3)
    ids = [101, 102, 103, 104]
    for i, f in enumerate((0, 2, 5, 9)):             # four keyframes, each cloud stamped with its id (the point-cloud route writes kfids)
        d = synth.depth_frame(f, w, h); c = synth.bgr_frame(f, w, h)
        xyz, rgb = scenario.cloud_from_depth(d, c if color else None, K, step=2)
        for m in (o, r):
            m.integrate_cloud_kf(xyz, rgb, synth.pose(f), d, kfid=ids[i])
    _same(o, r)
    order = r.chunk_order()
    assert len(order) == len(o.download()[0])
    # keyframe 103 is not in the map (its voxels are discarded, ChunkManager.cpp:955-962); the others get different corrections, so voxels collide
    kf = np.array([101, 102, 104], np.uint32)
    Rt = np.stack([_rt(rng, scale) for _ in kf])
    o.deform(kf, Rt, order=order); r.deform(kf, Rt)
    n = _same(o, r)
    assert n > 30
    # the deformed map keeps working: another cloud on top
    d = synth.depth_frame(11, w, h)
    xyz, rgb = scenario.cloud_from_depth(d, synth.bgr_frame(11, w, h) if color else None, K, step=3)
    for m in (o, r):
        m.integrate_cloud_kf(xyz, rgb, synth.pose(11), d, kfid=105)
    _same(o, r)


def test_deform_order_matters_and_key_order_is_deterministic():
    """the same deformation with a different visiting order gives a different map where voxels collide -- the reason the order is an input"""
    o1, r, K, w, h = _pair(voxel_resolution=0.04, use_carving=0, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=1)
    o2 = OT.Map(o1.params, threads=4); o2.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    rng = np.random.default_rng(5)
    for i, f in enumerate((0, 3)):
        d = synth.depth_frame(f, w, h)
        xyz, rgb = scenario.cloud_from_depth(d, synth.bgr_frame(f, w, h), K, step=2)
        for m in (o1, o2):
            m.integrate_cloud_kf(xyz, rgb, synth.pose(f), d, kfid=7 + i)
    kf = np.array([7, 8], np.uint32); Rt = np.stack([_rt(rng, 0.2), _rt(rng, 0.2)])
    keys = o1.download()[0]
    o1.deform(kf, Rt); o2.deform(kf, Rt, order=keys[::-1])
    a, b = o1.download(), o2.download()
    assert np.array_equal(a[0], b[0])                          # the same cells are occupied ...
    assert not np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))     # ... but collisions folded in another order
    o3 = OT.Map(o1.params, threads=4)
    assert o3 is not None


@pytest.mark.parametrize("color", [1])
def test_world_cloud_with_normals_equals_reference(color):
    o, r, K, w, h = _pair(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=color)
    rng = np.random.default_rng(9)
    for f in (0, 4):
        d = synth.depth_frame(f, w, h)
        xyz, rgb = scenario.cloud_from_depth(d, synth.bgr_frame(f, w, h), K, step=2)
        Pw = (xyz @ synth.pose(f)[:, :3].T + synth.pose(f)[:, 3]).astype(np.float32)      # a saved map is in the world frame
        nrm = rng.normal(size=Pw.shape).astype(np.float32) * np.float32(0.2) + np.array([0, 0, -1], np.float32)
        nrm[::97] = 0                                                                      # degenerate normals: normalized() leaves them
        kf = rng.integers(1, 9, len(Pw)).astype(np.uint32)
        Twc = np.eye(4, dtype=np.float32)[:3]
        for m in (o, r):
            m.integrate_world_cloud(Pw, rgb if color else None, nrm, Twc, kfids=kf)
        n = _same(o, r)
    assert n > 30
    assert o.stats()["n_blocks"] == r.stats()["n_blocks"]
