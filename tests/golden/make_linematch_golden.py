#!/usr/bin/env python3
"""Records tests/golden/linematch_ref.npz from the reference's own line-descriptor matcher (oracle/_ref/liblinematch_ref.so =
/root/reference/Thirdparty/line_descriptor/src/binary_descriptor_matcher_custom.cpp compiled unmodified): 2-NN indices in the library's order,
distances, ratio-test flags for the seeded cases of tests/linematch_cases.py.  Run where /root/reference exists: python tests/golden/make_linematch_golden.py"""
import pathlib, sys
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2]))
from oracle import linematch as L                  # noqa: E402
from tests.linematch_cases import cases            # noqa: E402

R = L.RefLineMatcher()
n, seed, ratio = 16, 5, 0.78
out = {"n_cases": n, "seed": seed, "nn_ratio": ratio}
for i, (q, t, mask) in enumerate(cases(n, seed=seed)):
    qi, ti, di, vi, nv = R.knn2(q, t, mask, ratio)
    out[f"qi{i}"], out[f"ti{i}"], out[f"di{i}"], out[f"vi{i}"] = qi, ti, di, vi
np.savez_compressed(pathlib.Path(__file__).parent / "linematch_ref.npz", **out)
print("written", n, "cases")
