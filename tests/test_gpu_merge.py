"""-m gpu: the optional voxel-block merge (not in the reference): export (key, w*sdf, w) of two maps that see
the same world from different poses, fold both into a third map, compare with the weighted sum computed on the host."""
import ctypes as C
import numpy as np
import pytest

from plvs_b200 import synth, tsdf as T

pytestmark = pytest.mark.gpu
VOX = 4096


def _export(g, color=False):
    import torch
    lib, h = g._lib, g._h
    n = C.c_int()
    assert lib.plvs_tsdf_export_packed(h, None, None, None, 0, C.byref(n)) == 0
    n = n.value
    dev = torch.device("cuda", 0)
    keys = torch.empty((n, 3), dtype=torch.int32, device=dev); ws = torch.empty((n, VOX), dtype=torch.float32, device=dev); w = torch.empty((n, VOX), dtype=torch.float32, device=dev)
    m = C.c_int()
    if color:
        rgba = torch.empty((n, VOX), dtype=torch.int32, device=dev)
        assert lib.plvs_tsdf_export_packed_rgba(h, C.c_void_p(keys.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(rgba.data_ptr()),
                                                n, C.byref(m)) == 0 and m.value == n
        return keys, ws, w, rgba
    assert lib.plvs_tsdf_export_packed(h, C.c_void_p(keys.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_void_p(w.data_ptr()), n, C.byref(m)) == 0 and m.value == n
    return keys, ws, w


def test_export_merge_two_maps(gpu):
    import torch
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=0)
    maps = []
    for frames in ((0, 1, 2), (2, 5, 9)):
        g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        for f in frames:
            g.integrate(synth.depth_frame(f, w, h), synth.pose(f))
        maps.append(g)
    host = [m.download() for m in maps]
    packs = [_export(m) for m in maps]
    for (keys, ws, wt), (hk, hs, hw, _) in zip(packs, host):           # the export is (key, w*sdf, w) of every live block
        order = np.lexsort((keys.cpu().numpy()[:, 2], keys.cpu().numpy()[:, 1], keys.cpu().numpy()[:, 0]))
        assert np.array_equal(keys.cpu().numpy()[order], hk)
        assert np.array_equal(wt.cpu().numpy()[order], hw)
        assert np.array_equal(ws.cpu().numpy()[order], np.where(hw > 0, hw * hs, 0).astype(np.float32))
    keys = torch.cat([packs[0][0], packs[1][0]]); ws = torch.cat([packs[0][1], packs[1][1]]); wt = torch.cat([packs[0][2], packs[1][2]])
    c = T.ChiselServer(p); c.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    rc = c._lib.plvs_tsdf_merge_packed(c._h, C.c_void_p(keys.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_void_p(wt.data_ptr()), len(keys))
    assert rc == 0, c._lib.plvs_last_error()
    ck, cs, cw, _ = c.download()
    from tests.merge_expect import fold, compare
    exp = fold([(hk, hs, hw) for (hk, hs, hw, _) in host])              # list order: map 0 first, then map 1
    assert compare(exp, ck, cs, cw) == 0
    both = set(map(tuple, host[0][0])) & set(map(tuple, host[1][0]))
    assert len(both) > 10                                              # the two maps really overlap


def test_export_merge_two_colour_maps(gpu):
    """the packed format carries the ColorVoxel state: two colour maps folded into a third, distances and colours against the host fold"""
    import torch
    from tests.merge_expect import fold, compare
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=1)
    maps = []
    for frames in ((0, 1, 2), (2, 5, 9)):
        g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        for f in frames:
            g.integrate(synth.depth_frame(f, w, h), synth.pose(f), synth.bgr_frame(f, w, h))
        maps.append(g)
    host = [m.download() for m in maps]
    packs = [_export(m, color=True) for m in maps]
    for (keys, ws, wt, rgba), (hk, hs, hw, hc) in zip(packs, host):
        kk = keys.cpu().numpy()
        order = np.lexsort((kk[:, 2], kk[:, 1], kk[:, 0]))
        assert np.array_equal(rgba.cpu().numpy().view(np.uint8).reshape(-1, VOX, 4)[order], hc)
    keys, ws, wt, rgba = (torch.cat([packs[0][i], packs[1][i]]) for i in range(4))
    c = T.ChiselServer(p); c.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    rc = c._lib.plvs_tsdf_merge_packed_rgba(c._h, C.c_void_p(keys.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_void_p(wt.data_ptr()), C.c_void_p(rgba.data_ptr()), len(keys))
    assert rc == 0, c._lib.plvs_last_error()
    ck, cs, cw, cc = c.download()
    assert compare(fold(host), ck, cs, cw, cc) == 0
    assert int((cc[..., 3] > 1).sum()) > 1000
