"""CPU: StreamData.from_tum -- the loader behind `bench.py --dataset DIR` (BASELINE.json configs[0]: a TUM RGB-D sequence as Examples_old/RGB-D/rgbd_tum.cc reads
it).  No dataset is mounted here, so the test writes a small sequence in the benchmark's on-disk format (PNG colour, 16-bit PNG depth x 5000, rgb.txt /
depth.txt / groundtruth.txt, optionally an associations file) from the synthetic generator and reads it back."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from plvs_b200 import synth
from plvs_b200.pipeline import StreamData


def _write_sequence(root, n, w, h, with_associations):
    (root / "rgb").mkdir(); (root / "depth").mkdir()
    rgb_lines, dep_lines, gt_lines, assoc = ["# color images", "# timestamp filename"], ["# depth maps", "# timestamp filename"], ["# ground truth trajectory", "# timestamp tx ty tz qx qy qz qw"], []
    imgs, deps, poses = [], [], []
    for f in range(n):
        t_rgb, t_dep = 1305031102.175304 + f / 30.0, 1305031102.160407 + f / 30.0
        bgr = synth.bgr_frame(f, w, h); d = synth.depth_frame(f, w, h)
        d16 = np.clip(np.round(np.nan_to_num(d, nan=0.0) * 5000.0), 0, 65535).astype(np.uint16)
        cv2.imwrite(str(root / f"rgb/{t_rgb:.6f}.png"), bgr); cv2.imwrite(str(root / f"depth/{t_dep:.6f}.png"), d16)
        rgb_lines.append(f"{t_rgb:.6f} rgb/{t_rgb:.6f}.png"); dep_lines.append(f"{t_dep:.6f} depth/{t_dep:.6f}.png")
        assoc.append(f"{t_rgb:.6f} rgb/{t_rgb:.6f}.png {t_dep:.6f} depth/{t_dep:.6f}.png")
        T = synth.pose(f).astype(np.float64)
        R = T[:, :3]; qw = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
        qx, qy, qz = (R[2, 1] - R[1, 2]) / (4 * qw), (R[0, 2] - R[2, 0]) / (4 * qw), (R[1, 0] - R[0, 1]) / (4 * qw)
        gt_lines.append(f"{t_rgb + 0.001:.4f} {T[0, 3]:.6f} {T[1, 3]:.6f} {T[2, 3]:.6f} {qx:.8f} {qy:.8f} {qz:.8f} {qw:.8f}")
        imgs.append(bgr); deps.append(d16); poses.append(T)
    (root / "rgb.txt").write_text("\n".join(rgb_lines) + "\n"); (root / "depth.txt").write_text("\n".join(dep_lines) + "\n")
    (root / "groundtruth.txt").write_text("\n".join(gt_lines) + "\n")
    if with_associations:
        (root / "associations.txt").write_text("\n".join(assoc) + "\n")
    return imgs, deps, poses


@pytest.mark.parametrize("with_associations", [False, True])
def test_tum_sequence_round_trip(tmp_path, with_associations):
    n, w, h = 5, 160, 120
    imgs, deps, poses = _write_sequence(tmp_path, n, w, h, with_associations)
    d = StreamData.from_tum(tmp_path, 4, pinned=False)
    assert (d.n, d.w, d.h) == (4, w, h) and d.K["fx"] == synth.intrinsics(w, h)["fx"]
    for f in range(4):
        assert np.array_equal(d.bgr[f], imgs[f])
        assert np.array_equal(d.gray[f], cv2.cvtColor(imgs[f], cv2.COLOR_RGB2GRAY))          # Camera.RGB: 1 applied to imread's BGR data, as the reference does
        assert np.array_equal(d.depth[f], deps[f].astype(np.float32) * np.float32(1.0 / 5000.0))
        assert np.allclose(d.poses[f], poses[f], atol=2e-6)
    with pytest.raises(ValueError):
        StreamData.from_tum(tmp_path, 9, pinned=False)           # more frames than the sequence holds
