"""Generates tests/golden/match_ref.npz from the REFERENCE's own ORBmatcher.cc (oracle/_ref/libmatch_ref.so, compiled from
/root/reference by oracle/ref_build.py).  Run in the build container:   python tests/golden/make_match_golden.py
Inputs are regenerated at test time (seeded synthetic frames through the CPU extractor oracle); only the reference's
OUTPUTS (assign / match arrays and counts) are stored."""
import pathlib, sys
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from plvs_b200 import synth, scenario                       # noqa: E402
from plvs_b200.matcher import featvec                       # noqa: E402
from oracle import match as OM, orb as O                    # noqa: E402


def frames():
    K = synth.intrinsics(640, 480)
    tab = O.Tables(2000)
    out = []
    for f in (10, 11, 15):
        kp, desc, mono, _ = O.extract_port(synth.gray_frame(f), 2000)
        fr = scenario.make_frame(kp, desc, synth.depth_frame(f), K, tab.scale)
        fr.level_sigma2 = tab.sigma2
        out.append((fr, synth.pose(f)))
    return K, out


def cases(K, fr):
    """name -> (kind, args): the same inputs for the reference, the oracle and the CUDA path"""
    (last, Tl), (cur, Tc), (kf2, T2) = fr
    out = {}
    q, _ = scenario.map_queries(last, cur, K, Tl, Tc)
    out["map_th3"] = ("map", dict(F=cur, q=q, th=3.0, ratio=0.8, far=False, th_far=50.0, claimed=None))
    rng = np.random.default_rng(3)
    q2, _ = scenario.map_queries(last, cur, K, Tl, Tc, seed=3)
    claimed = (rng.random(cur.n) < 0.3).astype(np.uint8)
    q2["flags"] = (rng.random(len(q2)) < 0.9).astype(np.uint32)
    q2 = np.concatenate([q2, q2[::2], q2[::3]])
    out["map_competition"] = ("map", dict(F=cur, q=q2, th=5.0, ratio=0.8, far=True, th_far=3.0, claimed=claimed))
    ql, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    ql, z = OM.canonical_last_queries(ql)
    out["last_th15"] = ("last", dict(C=cur, q=ql, z=z, th=15.0, fwd=False, bwd=False, check=True, claimed=None))
    out["last_backward"] = ("last", dict(C=cur, q=ql, z=z, th=15.0, fwd=False, bwd=True, check=True, claimed=None))
    ql2 = np.concatenate([ql, ql[::2]])
    rng = np.random.default_rng(2)
    ql2["flags"] = (rng.random(len(ql2)) < 0.8).astype(np.uint32)
    ql2, z2 = OM.canonical_last_queries(ql2)
    out["last_competition"] = ("last", dict(C=cur, q=ql2, z=z2, th=15.0, fwd=False, bwd=False, check=True, claimed=(rng.random(cur.n) < 0.2).astype(np.uint8)))
    fv1, fv2 = featvec(scenario.node_ids(last.desc, 128)), featvec(scenario.node_ids(kf2.desc, 128))
    rng = np.random.default_rng(4)
    has1 = (rng.random(last.n) < 0.4).astype(np.uint8); has2 = (rng.random(kf2.n) < 0.4).astype(np.uint8)
    F12, ep = scenario.fundamental(K, Tl, T2)
    for nm, coarse, only in (("tri", False, False), ("tri_coarse", True, False), ("tri_stereo", False, True)):
        out[nm] = ("tri", dict(K1=last, K2=kf2, fv1=fv1, fv2=fv2, has1=has1, has2=has2, F12=F12, ep=ep, only=only, coarse=coarse, check=True))
    return out


def run_ref(kind, a):
    if kind == "map":
        return OM.ref_search_by_projection_map(a["F"], a["q"], a["th"], a["ratio"], a["far"], a["th_far"], a["claimed"])
    if kind == "last":
        return OM.ref_search_by_projection_last(a["C"], a["q"], a["z"], a["th"], a["fwd"], a["bwd"], a["check"], a["claimed"])
    return OM.ref_search_for_triangulation(a["K1"], a["K2"], a["fv1"], a["fv2"], a["has1"], a["has2"], a["F12"], a["ep"], a["only"], a["coarse"], a["check"])


def run_oracle(kind, a):
    if kind == "map":
        return OM.search_by_projection_map(a["F"], a["q"], a["th"], a["ratio"], a["far"], a["th_far"], a["claimed"])
    if kind == "last":
        return OM.search_by_projection_last(a["C"], a["q"], a["th"], a["fwd"], a["bwd"], a["check"], a["claimed"])
    return OM.search_for_triangulation(a["K1"], a["K2"], a["fv1"], a["fv2"], a["has1"], a["has2"], a["F12"], a["ep"], a["only"], a["coarse"], a["check"])


if __name__ == "__main__":
    K, fr = frames()
    store = {}
    for name, (kind, a) in cases(K, fr).items():
        n, arr = run_ref(kind, a)
        store[name + "_n"] = np.int32(n); store[name] = arr.astype(np.int32)
        print(name, n)
    out = pathlib.Path(__file__).with_name("match_ref.npz")
    np.savez_compressed(out, source="oracle/_ref/libmatch_ref.so (reference src/ORBmatcher.cc + stand-in data model)", **store)
    print(out, out.stat().st_size)
