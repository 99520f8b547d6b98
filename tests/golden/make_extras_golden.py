"""Generates tests/golden/extras_ref.npz from the REFERENCE's own code (oracle/_ref/*.so, compiled from /root/reference by oracle/ref_build.py) for the
rows widened last: SearchForInitialization (src/ORBmatcher.cc), the mesh read-out (Thirdparty/open_chisel) and the bag-of-words
transform (Thirdparty/DBoW2).  Run in the build container:   python tests/golden/make_extras_golden.py
Inputs are regenerated at test time (seeded synthetic data, tests/golden/extras_cases.py); only the reference's OUTPUTS are stored (meshes as counts
plus a digest of the float arrays, to keep the fixture small)."""
import hashlib
import pathlib
import sys
import tempfile

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import match as OM, tsdf as OT, bow as OB            # noqa: E402
from tests.golden import extras_cases as X                        # noqa: E402


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


def main():
    out = {}
    f1, f2, f3 = X.init_frames()
    prev = np.stack([f1.keys["x"], f1.keys["y"]], 1)
    for window, ratio, check in X.INIT_PARAMS:
        n, m, p = OM.ref_search_for_initialization(f1, f2, prev, window, ratio, check)
        n2, m2, p2 = OM.ref_search_for_initialization(f1, f3, p, window, ratio, check)
        tag = "init_%d_%g_%d" % (window, ratio, check)
        out[tag + "_n"] = np.array([n, n2]); out[tag + "_m"] = np.stack([m, m2]); out[tag + "_p"] = np.stack([p, p2])
    for name, kw, frames, color in X.MESH_CASES:
        r = X.mesh_map(OT.RefMap, kw, frames, color)
        keys, counts, V, N, C = r.extract_mesh()
        out["mesh_%s_keys" % name] = keys; out["mesh_%s_counts" % name] = counts; out["mesh_%s_digest" % name] = digest(V, N, C)
    desc = X.bow_descriptors()
    tmp = pathlib.Path(tempfile.mkdtemp())
    for k, L, levelsup, scoring, weighting, zero in X.BOW_CASES:
        path = tmp / "voc.txt"
        OB.write_vocabulary(path, k, L, seed=k * 10 + L, scoring=scoring, weighting=weighting, zero_weight_fraction=zero)
        r = OB.RefVocabulary(path).transform(desc, levelsup)
        tag = "bow_%d_%d_%d_%d_%d" % (k, L, levelsup, scoring, weighting)
        for name in ("word", "node", "bow_ids", "bow_vals", "fv_nodes", "fv_offsets"):
            out[tag + "_" + name] = r[name]
        out[tag + "_digest"] = digest(r["weight"], r["fv_features"])
    dst = pathlib.Path(__file__).resolve().parent / "extras_ref.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
