"""Pins the mesh read-out oracle (oracle/tsdf_oracle.cpp: orc_tsdf_extract_mesh) to the REFERENCE's own open_chisel
(ChunkManager::RecomputeMesh = GenerateMesh + ColorizeMesh + ComputeNormalsFromGradients, MarchingCubes::MeshCube), compiled into
oracle/_ref/libchisel_ref.so: vertex order, positions, normals and colours bit for bit, and the reference's incremental flow
(Chisel::UpdateMeshes after every integration) against a full pass."""
import numpy as np
import pytest

from plvs_b200 import synth, tsdf as T
from oracle import tsdf as OT

pytestmark = pytest.mark.skipif(not OT.ref_available(), reason="oracle/_ref/libchisel_ref.so not built (/root/reference absent)")


def _same(a, b):
    ka, ca, Va, Na, Ca = a; kb, cb, Vb, Nb, Cb = b
    assert np.array_equal(ka, kb) and np.array_equal(ca, cb)
    for x, y in ((Va, Vb), (Na, Nb), (Ca, Cb)):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def _pair(w, h, **kw):
    K = synth.intrinsics(w, h)
    p = T.default_params(**kw)
    o = OT.Map(p, threads=8); r = OT.RefMap(p)
    for m in (o, r):
        m.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    return o, r


@pytest.mark.parametrize("color", [0, 1])
def test_mesh_after_depth_scans(color):
    o, r = _pair(160, 120, voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=color)
    for f in (0, 1, 2, 6):
        d = synth.depth_frame(f, 160, 120)
        c = synth.bgr_frame(f, 160, 120) if color else None
        o.integrate(d, synth.pose(f), c); r.integrate(d, synth.pose(f), c)
        r.update_meshes()                                   # the reference's flow: re-mesh what the scan touched
        inc = r.meshes()
        _same(o.extract_mesh(), inc)
    full = r.extract_mesh()                                 # every chunk again: nothing changes
    _same(inc, full)
    keys, counts, V, N, C = full
    assert len(keys) > 40 and counts.sum() == len(V) > 10000 and (counts % 3 == 0).all()
    assert np.allclose(np.linalg.norm(N, axis=1), 1.0, atol=1e-5)
    if color:
        assert C.max() > 0.2 and C.max() <= 1.0
    else:
        assert (C == 0).all()


def test_mesh_near_the_world_origin_and_retreating_surface():
    """a wall close to the world origin (InterpolateColor's look-ups by voxel index can land on existing chunks there) that then jumps back
    (carving resets voxels: meshes shrink or vanish)"""
    o, r = _pair(128, 96, voxel_resolution=0.05, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=1)
    Twc = np.eye(4, dtype=np.float32)[:3].copy()
    Twc[:, 3] = (-0.1, -0.05, -0.45)
    for depth_m in (0.6, 0.6, 0.6, 1.7, 1.7):
        d = np.full((96, 128), depth_m, np.float32)
        d += (np.arange(128, dtype=np.float32) * 0.002)[None, :]
        c = synth.bgr_frame(1, 128, 96)
        o.integrate(d, Twc, c); r.integrate(d, Twc, c)
        r.update_meshes()
        _same(o.extract_mesh(), r.meshes())
    assert len(o.extract_mesh()[2]) > 1000


def test_mesh_after_cloud_integration():
    from plvs_b200 import scenario
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    o, r = _pair(w, h, voxel_resolution=0.04, use_carving=1, carving_dist=0.05, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=1)
    for f in (0, 1, 3):
        d = synth.depth_frame(f, w, h); c = synth.bgr_frame(f, w, h)
        xyz, rgb = scenario.cloud_from_depth(d, c, K, step=2)
        o.integrate_cloud(xyz, rgb, synth.pose(f), d); r.integrate_cloud(xyz, rgb, synth.pose(f), d)
        r.update_meshes()
        _same(o.extract_mesh(), r.meshes())
    assert len(o.extract_mesh()[2]) > 5000


def test_empty_map():
    o, r = _pair(160, 120, voxel_resolution=0.04, max_blocks=64, use_color=1)
    _same(o.extract_mesh(), r.extract_mesh())
    assert len(o.extract_mesh()[0]) == 0


def test_mesh_with_coarse_voxels_takes_the_trilinear_colour_branch():
    """InterpolateColor looks its eight voxels up by INDEX used as a metric position; with 0.5 m voxels those positions fall inside the map,
    so the trilinear branch (not the nearest-voxel fallback) produces the colours -- restated as written"""
    o, r = _pair(160, 120, voxel_resolution=0.5, use_carving=1, near_plane=0.1, far_plane=6.0, max_blocks=4096, use_color=1)
    for f in (0, 3):
        d = synth.depth_frame(f, 160, 120); c = synth.bgr_frame(f, 160, 120)
        o.integrate(d, synth.pose(f), c); r.integrate(d, synth.pose(f), c)
        r.update_meshes()
        _same(o.extract_mesh(), r.meshes())
    keys, counts, V, N, C = o.extract_mesh()
    blocks = set(map(tuple, o.download()[0]))
    v0 = np.floor(V / np.float32(0.5)).astype(int)
    live = sum(all(tuple(np.floor((v + np.array(dv)) / 8.0).astype(int)) in blocks for dv in np.ndindex(2, 2, 2)) for v in v0)
    assert len(V) > 100 and live > 20


def _ply_parts(path):
    lines = open(path).read().split("\n")
    e = lines.index("end_header")
    nv = int([l for l in lines[:e] if l.startswith("element vertex")][0].split()[-1])
    return lines[:e + 1], lines[e + 1:e + 1 + nv], lines[e + 1 + nv:]


@pytest.mark.parametrize("color", [0, 1])
def test_ply_writer_against_the_reference_file(tmp_path, color):
    """plvs_mesh_save_ply (host I/O behind the C ABI) fed with the oracle's meshes == Chisel::SaveAllMeshesToPLY of the compiled reference: same header,
    same face list, the same vertex lines per chunk (the reference walks an unordered_map, so chunks come in another order: compared as sorted triangles)"""
    o, r = _pair(160, 120, voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=color)
    for f in (0, 1):
        d = synth.depth_frame(f, 160, 120); c = synth.bgr_frame(f, 160, 120) if color else None
        o.integrate(d, synth.pose(f), c); r.integrate(d, synth.pose(f), c)
    r.update_meshes()
    r.save_ply(tmp_path / "ref.ply")
    _, _, V, _, Cc = o.extract_mesh()
    T.save_ply(tmp_path / "own.ply", V, Cc if color else None)
    rh, rv, rf = _ply_parts(tmp_path / "ref.ply"); oh, ov, of = _ply_parts(tmp_path / "own.ply")
    assert oh == rh and of == rf and len(ov) == len(V) > 3000
    tri = lambda v: sorted(tuple(v[i:i + 3]) for i in range(0, len(v), 3))
    assert tri(ov) == tri(rv)


def test_keyframe_ids_of_voxels_and_mesh_vertices():
    """DistVoxel::kfid (set by the point-cloud route per integrated point, zeroed by Reset) and Mesh::kfids (the cube's first corner): oracle == compiled
    reference, with one id per cloud and with per-point ids"""
    from plvs_b200 import scenario
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    o, r = _pair(w, h, voxel_resolution=0.04, use_carving=1, carving_dist=0.05, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=1)
    rng = np.random.default_rng(7)
    for step, f in enumerate((0, 1, 3, 6)):
        d = synth.depth_frame(f, w, h) + np.float32(0.15 * step)         # the surface retreats: carving resets voxels (kfid back to 0)
        c = synth.bgr_frame(f, w, h)
        xyz, rgb = scenario.cloud_from_depth(d, c, K, step=2)
        kfids = None if step % 2 == 0 else rng.integers(1, 50, len(xyz)).astype(np.uint32)
        for m in (o, r):
            m.integrate_cloud_kf(xyz, rgb, synth.pose(f), d, kfids=kfids, kfid=100 + f)
        ko, kr = o.download_kfid(), r.download_kfid()
        assert np.array_equal(o.download()[0], r.download()[0]) and np.array_equal(ko, kr)
        assert (ko > 0).sum() > 1000
        r.update_meshes()
        mo, mr = o.extract_mesh(), r.meshes()
        _same(mo, mr)
        assert np.array_equal(o.mesh_kfids(len(mo[2])), r.mesh_kfids(len(mr[2])))
    wts = o.download()[2]
    assert ((ko > 0) <= (wts > 0)).all()                                 # a keyframe id only on observed voxels: Reset clears both
    assert len(np.unique(ko)) > 10
