"""Pins the oracle of Frame::ComputeStereoMatches (a17; oracle/match_oracle.cpp) to the REFERENCE's own function body: oracle/ref_build.py
slices the definition out of /root/reference/src/Frame.cc at build time (into the git-ignored oracle/_ref/gen/) and compiles it, together
with the reference's ORBextractor.cc, against the stand-ins.  Bit-exact mvuRight / mvDepth on the EuRoC-like geometry of BASELINE config 5.
Skipped when neither /root/reference nor a prebuilt oracle/_ref is present."""
import numpy as np
import pytest

from plvs_b200 import synth
from plvs_b200.matcher import Frame
from oracle import match as OM, orb as O

pytestmark = pytest.mark.skipif(not OM.stereo_ref_available(), reason="oracle/_ref/libstereo_ref.so not built (/root/reference absent)")


@pytest.mark.parametrize("w,h,nfeat,frame,baseline", [(752, 480, 1200, 0, 0.11), (752, 480, 1200, 4, 0.11), (640, 480, 1000, 2, 0.08), (320, 240, 500, 1, 0.2)])
def test_compute_stereo_matches(w, h, nfeat, frame, baseline):
    K = synth.intrinsics(w, h)
    mbf = np.float32(K["fx"] * baseline)
    il, ir = synth.gray_frame(frame, w, h), synth.gray_frame(frame, w, h, eye=baseline)
    rkeys, rur, rdp, n_right = OM.ref_compute_stereo_matches(il, ir, nfeat, baseline, float(mbf))
    # the oracle on the oracle's own extraction (which the other tests pin to the reference's extractor)
    kl, dl, _, _, intl = O.extract_cv2(il, nfeat, angle_impl="c", return_internals=True)
    kr, dr, _, _, intr = O.extract_cv2(ir, nfeat, angle_impl="c", return_internals=True)
    assert len(kl) == len(rkeys) and len(kr) == n_right
    for i, f in enumerate(("x", "y", "size", "angle", "response")):
        assert np.array_equal(kl[f], rkeys[:, i])
    tab = O.Tables(nfeat)
    L = Frame(kl, dl, w, h, tab.scale, bf=float(mbf)); R = Frame(kr, dr, w, h, tab.scale, bf=float(mbf))
    our, odp, kept = OM.compute_stereo_matches(L, R, intl["pyramid"], intr["pyramid"], tab.scale, tab.inv_scale, baseline, float(mbf))
    assert np.array_equal(our.view(np.uint32), rur.view(np.uint32))
    assert np.array_equal(odp.view(np.uint32), rdp.view(np.uint32))
    assert (rdp > 0).sum() == kept and kept > (60 if w < 400 else 250)
