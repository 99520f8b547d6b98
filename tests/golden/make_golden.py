"""Generate the committed golden vectors (tests/golden/*.npz).

No golden vectors exist in the reference, so these come from the arm of the oracle that
executes the real OpenCV primitives (cv2, version recorded in the file) in the reference's call pattern
(oracle/orb.py: extract_cv2 with cv2.fastAtan2 for the orientation).  Run in the build container:
    python tests/golden/make_golden.py
"""
import pathlib, sys
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import cv2
from oracle import orb as O
from plvs_b200 import synth

out = pathlib.Path(__file__).resolve().parent
for name, (w, h, frame, nfeat) in {"orb_qvga_f0_500": (320, 240, 0, 500), "orb_vga_f3_1000": (640, 480, 3, 1000)}.items():
    img = synth.gray_frame(frame, w, h)
    kp, desc, mono, ncand = O.extract_cv2(img, nfeat, angle_impl="cv2")
    np.savez_compressed(out / f"{name}.npz", image=img, keypoints=kp, descriptors=desc, mono_index=mono, n_candidates=ncand,
                        cv2_version=cv2.__version__, params=np.array([nfeat, 8, 20, 7]), scale_factor=np.float32(1.2))
    print(name, len(kp), ncand)
# fastAtan2 known answers straight from OpenCV's scalar implementation
rng = np.random.default_rng(0)
yx = rng.integers(-200000, 200000, size=(20000, 2)).astype(np.float32)
yx[:8] = [[0, 0], [5, 0], [0, 5], [0, -5], [-5, 0], [1, 1], [-1, 1], [3, -4]]
ang = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
np.savez_compressed(out / "fast_atan2.npz", yx=yx, angle=ang, cv2_version=cv2.__version__)
print("fast_atan2", len(ang))
