"""TSDF goldens recorded from the REFERENCE's own open_chisel sources (tests/golden/make_tsdf_golden.py):
 * not gpu: the oracle restatement reproduces them bit-exactly (runs everywhere, also without /root/reference);
 * gpu: the CUDA path through the C ABI reproduces them (keys identical, sdf/weight within 1e-4, colours identical),
        and, when oracle/_ref travelled to the box, agrees with the compiled reference run live."""
import importlib.util
import pathlib
import numpy as np
import pytest

from plvs_b200 import synth, tsdf as T
from oracle import tsdf as OT

G = pathlib.Path(__file__).parent / "golden"
spec = importlib.util.spec_from_file_location("make_tsdf_golden", G / "make_tsdf_golden.py")
mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)


def _golden(name):
    z = np.load(G / f"tsdf_ref_{name}.npz")
    return z["keys"], z["sdf"], z["weight"], z["rgba"]


@pytest.mark.parametrize("name", list(mk.CASES))
def test_oracle_reproduces_reference_golden(name):
    kw, steps = mk.CASES[name]
    p = T.default_params(max_blocks=4096, **kw)
    o = OT.Map(p, threads=4)
    K = synth.intrinsics(mk.W, mk.H)
    o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], mk.W, mk.H)
    keys, sdf, w, rgba = mk.run(o, kw, steps)
    gk, gs, gw, gc = _golden(name)
    assert np.array_equal(keys, gk) and np.array_equal(sdf, gs) and np.array_equal(w, gw)
    if kw["use_color"]:
        assert np.array_equal(rgba, gc)


class _Cuda:
    """adapter: plvs_b200.tsdf.ChiselServer with the oracle's method names"""
    def __init__(self, p, K, w, h):
        self.g = T.ChiselServer(p)
        self.g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    def integrate(self, d, P, c=None): self.g.integrate(d, P, c)
    def integrate_cloud(self, xyz, rgb, P, d=None): self.g.integrate_cloud(xyz, rgb, P, d)
    def download(self): return self.g.download()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(mk.CASES))
def test_cuda_reproduces_reference_golden(gpu, name):
    kw, steps = mk.CASES[name]
    p = T.default_params(max_blocks=4096, **kw)
    K = synth.intrinsics(mk.W, mk.H)
    keys, sdf, w, rgba = mk.run(_Cuda(p, K, mk.W, mk.H), kw, steps)
    gk, gs, gw, gc = _golden(name)
    assert np.array_equal(keys, gk)
    assert np.abs(w - gw).max() <= 1e-4
    known = gw > 0
    assert np.abs(sdf[known] - gs[known]).max() <= 1e-4 and np.array_equal(sdf[~known], gs[~known])
    if kw["use_color"]:
        assert np.array_equal(rgba, gc)


@pytest.mark.gpu
def test_cuda_vs_compiled_reference_live(gpu):
    if not OT.ref_available():
        pytest.skip("oracle/_ref/libchisel_ref.so did not travel to this box")
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    p = T.default_params(voxel_resolution=0.04, near_plane=0.1, far_plane=4.0, max_blocks=8192, use_color=1, use_carving=1)
    g = _Cuda(p, K, w, h); r = OT.RefMap(p); r.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    for f in (0, 1, 2, 3, 8):
        d, c = synth.depth_frame(f, w, h), synth.bgr_frame(f, w, h)
        g.integrate(d, synth.pose(f), c); r.integrate(d, synth.pose(f), c)
        gk, gs, gw, gc = g.download(); rk, rs, rw, rc = r.download()
        assert np.array_equal(gk, rk) and np.abs(gw - rw).max() <= 1e-4 and np.array_equal(gc, rc)
        assert np.abs(gs[rw > 0] - rs[rw > 0]).max() <= 1e-4
        assert g.g.stats()["n_range"] == r.stats()["n_range"]
