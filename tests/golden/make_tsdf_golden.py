"""Generates tests/golden/tsdf_ref_*.npz from the REFERENCE's own open_chisel (oracle/_ref/libchisel_ref.so, compiled
from /root/reference by oracle/ref_build.py against the Eigen stand-in).  Run in the build container:
    python tests/golden/make_tsdf_golden.py
The inputs are regenerated from plvs_b200.synth at test time (seeded), only the reference OUTPUTS are stored."""
import contextlib, os, pathlib, sys
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from plvs_b200 import synth, scenario, tsdf as T          # noqa: E402
from oracle import tsdf as OT                              # noqa: E402

W, H = 128, 96
CASES = {
    # name: (params, list of (route, frame))
    "scan_colour_carve": (dict(voxel_resolution=0.05, near_plane=0.1, far_plane=4.0, use_color=1, use_carving=1), [("scan", 0), ("scan", 1), ("scan", 6)]),
    "scan_plain": (dict(voxel_resolution=0.05, near_plane=0.1, far_plane=4.0, use_color=0, use_carving=1), [("scan", 0), ("scan", 3)]),
    "cloud_colour_carve": (dict(voxel_resolution=0.05, near_plane=0.1, far_plane=4.0, use_color=1, use_carving=1), [("cloud", 0), ("cloud", 2), ("scan", 4)]),
}


def inputs(route, f):
    d = synth.depth_frame(f, W, H)
    if f in (3, 6):
        d = np.maximum(d - np.float32(0.3), 0).astype(np.float32)       # a closer surface: carving / reset has work
    return d, synth.bgr_frame(f, W, H), synth.pose(f)


def run(m, params, steps):
    K = synth.intrinsics(W, H)
    for route, f in steps:
        d, c, P = inputs(route, f)
        colour = params["use_color"]
        if route == "scan":
            m.integrate(d, P, c if colour else None)
        else:
            xyz, rgb = scenario.cloud_from_depth(d, c, K, step=2)
            m.integrate_cloud(xyz, rgb, P, d)
    return m.download()


if __name__ == "__main__":
    for name, (kw, steps) in CASES.items():
        p = T.default_params(max_blocks=4096, **kw)
        r = OT.RefMap(p)
        K = synth.intrinsics(W, H)
        r.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], W, H)
        keys, sdf, w, rgba = run(r, kw, steps)
        out = pathlib.Path(__file__).with_name(f"tsdf_ref_{name}.npz")
        np.savez_compressed(out, keys=keys, sdf=sdf, weight=w, rgba=rgba, source="oracle/_ref/libchisel_ref.so (reference open_chisel sources + Eigen stand-in)")
        print(out.name, keys.shape, out.stat().st_size)
