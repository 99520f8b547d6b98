"""Stereo frame construction (BASELINE.json configs[4], SURVEY.md §8e "C5"): the reference extracts the two eyes of a stereo frame in two threads
(src/Frame.cc:314-315: threadLeft / threadRight -> ExtractORB) and then matches them row by row (Frame::ComputeStereoMatches, :1780).  Here each eye
has its own extractor handle -- on the same GPU, or on two GPUs: the right eye's pyramid and keypoints then stay on the second GPU and the stereo
matcher on the first reads them in place over NVLink (peer access); only the right eye's image crosses the host bus."""
import threading
import numpy as np

from . import _lib
from .orb import ORBextractor
from .matcher import ORBmatcher, Frame, ComputeStereoMatches


class StereoFrontEnd:
    def __init__(self, nfeatures, width, height, baseline, fx, devices=(0, 0), scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        self.devices, self.w, self.h = tuple(devices), width, height
        self.mb, self.mbf = baseline, fx * baseline
        self.left = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=devices[0])
        self.right = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=devices[1])
        self.matcher = ORBmatcher(0.8, True, device=devices[0])
        if devices[0] != devices[1]:
            _lib.check(_lib.load().plvs_enable_peer_access(devices[0], devices[1]), "plvs_enable_peer_access")

    def __call__(self, im_left, im_right):
        """-> (keys_left, desc_left, keys_right, desc_right, mvuRight, mvDepth, n_stereo)"""
        out = [None, None]
        err = []

        def eye(i, ex, img):
            try:
                out[i] = ex(img)
            except Exception as e:          # surface worker failures in the caller
                err.append(e)
        t = threading.Thread(target=eye, args=(1, self.right, im_right))
        t.start()
        eye(0, self.left, im_left)
        t.join()
        if err:
            raise err[0]
        (_, kl, dl), (_, kr, dr) = out
        sf, isf = self.left.GetScaleFactors(), self.left.GetInverseScaleFactors()
        L = Frame(kl, dl, self.w, self.h, sf, bf=self.mbf)
        dv = self.right.device_result(0)
        # the right eye's keypoints / descriptors are read where the extractor left them (device memory of devices[1])
        R = Frame(None, None, self.w, self.h, sf, bf=self.mbf, device_ptrs=(dv.n, dv.keys, dv.desc, 0)) if dv.n == len(kr) else Frame(kr, dr, self.w, self.h, sf, bf=self.mbf)
        ur, dp, kept = ComputeStereoMatches(self.matcher, L, R, self.left.pyramid_view(0), self.right.pyramid_view(0), isf, self.mb, self.mbf)
        return kl, dl, kr, dr, ur, dp, kept
