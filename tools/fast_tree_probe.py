#!/usr/bin/env python3
"""Open issue: the min/max-tree corner score inside k_fast_cells (PLVS_FAST_TREE=1) disagreed with the bisection on the B200 although it
is right on the host, standalone on the device and on the CPU execution model.  This prints, per pyramid level, how many pixels of the device's score map
(inspection build, PLVS_ORB_DEBUG=1) differ from the oracle's, and for the first few the centre, the 16 ring differences and both scores -- the data needed
to see which arc the device's tree gets wrong.  Run on the GPU box: PLVS_FAST_TREE=1 PLVS_ORB_DEBUG=1 python tools/fast_tree_probe.py"""
import os
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
os.environ.setdefault("PLVS_FAST_TREE", "1")
os.environ.setdefault("PLVS_ORB_DEBUG", "1")

from plvs_b200 import synth                  # noqa: E402
from plvs_b200.orb import ORBextractor       # noqa: E402
from oracle import orb as O                  # noqa: E402  (tools are not product code)

RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def main():
    img = synth.gray_frame(0)
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    ex(img)
    tab = O.Tables(2000)
    pyr = O.pyramid_cv2(img, tab)
    total = 0
    for l in range(8):
        got = ex.pyramid_level(l, blurred=2)
        want = O.fast_score_map(pyr[l], 7)
        h, w = want.shape
        bad = np.argwhere(got[19:h - 19, 19:w - 19] != want[19:h - 19, 19:w - 19]) + 19
        total += len(bad)
        print("level %d: %d of %d scores differ" % (l, len(bad), (h - 38) * (w - 38)))
        for y, x in bad[:4]:
            v = int(pyr[l][y, x])
            d = [v - int(pyr[l][y + dy, x + dx]) for dx, dy in RING]
            print("   (%d,%d) centre %d  device %d  oracle %d  d = %s" % (x, y, v, int(got[y, x]), int(want[y, x]), d))
    print("total mismatches:", total)
    return total


if __name__ == "__main__":
    main()
