"""-m gpu: the stereo front end of BASELINE.json configs[4] (752x480, 1200 features per eye) with one extractor handle per eye, as the reference's two
extraction threads (src/Frame.cc:314-315), on one GPU and -- when the box has two -- split over two GPUs with the right eye's pyramid and
keypoints read in place over NVLink by Frame::ComputeStereoMatches on the first.  Both must give the oracle's mvuRight / mvDepth
bit for bit."""
import numpy as np
import pytest

from plvs_b200 import synth, _lib
from plvs_b200.matcher import Frame
from plvs_b200.stereo import StereoFrontEnd
from oracle import orb as O, match as OM

pytestmark = pytest.mark.gpu


def _check(devices):
    w, h, nfeat, baseline = 752, 480, 1200, 0.11
    K = synth.intrinsics(w, h)
    fe = StereoFrontEnd(nfeat, w, h, baseline, K["fx"], devices=devices)
    tab = O.Tables(nfeat)
    for frame in (0, 4):
        il, ir = synth.gray_frame(frame, w, h), synth.gray_frame(frame, w, h, eye=baseline)
        kl, dl, kr, dr, ur, dp, kept = fe(il, ir)
        okl, odl, _, _ = O.extract_port(il, nfeat); okr, odr, _, _ = O.extract_port(ir, nfeat)
        assert np.array_equal(kl, okl) and np.array_equal(dl, odl) and np.array_equal(kr, okr) and np.array_equal(dr, odr)
        L = Frame(okl, odl, w, h, tab.scale, bf=fe.mbf); R = Frame(okr, odr, w, h, tab.scale, bf=fe.mbf)
        pl = [fe.left.pyramid_level(l) for l in range(8)]; pr = [fe.right.pyramid_level(l) for l in range(8)]
        our, odp, okept = OM.compute_stereo_matches(L, R, pl, pr, tab.scale, fe.left.GetInverseScaleFactors(), baseline, fe.mbf)
        assert kept == okept > 300
        assert np.array_equal(ur.view(np.uint32), our.view(np.uint32)) and np.array_equal(dp.view(np.uint32), odp.view(np.uint32))


def test_two_extractor_handles_one_gpu(gpu):
    _check((0, 0))


def test_eye_split_over_two_gpus(gpu):
    if _lib.load().plvs_device_count() < 2:
        pytest.skip("one GPU on this box: the two-GPU eye split needs `gpurun --gpus 2`")
    _check((0, 1))
