"""Pins the matcher oracle (oracle/match_oracle.cpp, the restatement the CUDA searches are compared with) to the
REFERENCE's own src/ORBmatcher.cc, compiled from /root/reference by oracle/ref_build.py into oracle/_ref/libmatch_ref.so
(data-model stand-ins: oracle/plvs_standin/plvs_types.hpp).  Bit-exact assign arrays / match counts for the three searches
of SURVEY.md §8a (a11, a14, a15, a16), including claim competition, pre-claimed keypoints, map points without
observations, far-point gating, forward/backward level windows and the rotation histogram.  Skipped when neither
/root/reference nor a prebuilt oracle/_ref is present."""
import numpy as np
import pytest

from plvs_b200 import synth, scenario
from plvs_b200.matcher import featvec
from oracle import match as OM, orb as O

pytestmark = pytest.mark.skipif(not OM.ref_available(), reason="oracle/_ref/libmatch_ref.so not built (/root/reference absent)")


@pytest.fixture(scope="module")
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


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8); b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    for i in range(500):
        want = int(np.unpackbits(a[i] ^ b[i]).sum())
        assert OM.ref_hamming256(a[i], b[i]) == want == OM.hamming256(a[i], b[i])


@pytest.mark.parametrize("th", [1.0, 3.0, 5.0, 15.0])
def test_projection_map(frames, th):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.map_queries(last, cur, K, Tl, Tc)
    rn, ra = OM.ref_search_by_projection_map(cur, q, th, 0.8)
    on, oa = OM.search_by_projection_map(cur, q, th, 0.8)
    assert rn == on and np.array_equal(ra, oa)
    if th >= 3:
        assert rn > 200


def test_projection_map_claims_far_and_unobserved(frames):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    for seed in (3, 4, 5):
        q, _ = scenario.map_queries(last, cur, K, Tl, Tc, seed=seed)
        rng = np.random.default_rng(seed)
        claimed = (rng.random(cur.n) < 0.3).astype(np.uint8)
        q["flags"] = (rng.random(len(q)) < 0.9).astype(np.uint32)
        q = np.concatenate([q, q[::2], q[::3]])                     # duplicated queries fight for the same keypoints
        rn, ra = OM.ref_search_by_projection_map(cur, q, 5.0, 0.8, True, 3.0, claimed)
        on, oa = OM.search_by_projection_map(cur, q, 5.0, 0.8, True, 3.0, claimed)
        assert rn == on and np.array_equal(ra, oa)


@pytest.mark.parametrize("th,fwd,bwd", [(15.0, False, False), (7.0, False, False), (15.0, True, False), (15.0, False, True)])
def test_projection_last(frames, th, fwd, bwd):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    q, z = OM.canonical_last_queries(q)
    for check in (True, False):
        rn, ra = OM.ref_search_by_projection_last(cur, q, z, th, fwd, bwd, check)
        on, oa = OM.search_by_projection_last(cur, q, th, fwd, bwd, check)
        assert rn == on and np.array_equal(ra, oa)
    if not fwd:
        assert rn > 300


def test_projection_last_competition_and_claims(frames):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    q = np.concatenate([q, q[::2]])
    rng = np.random.default_rng(2)
    q["flags"] = (rng.random(len(q)) < 0.8).astype(np.uint32)
    q["invz"][::17] *= -1                                           # points behind the camera
    q, z = OM.canonical_last_queries(q)
    claimed = (rng.random(cur.n) < 0.2).astype(np.uint8)
    rn, ra = OM.ref_search_by_projection_last(cur, q, z, 15.0, claimed=claimed)
    on, oa = OM.search_by_projection_last(cur, q, 15.0, claimed=claimed)
    assert rn == on and np.array_equal(ra, oa)


@pytest.mark.parametrize("coarse,only_stereo", [(False, False), (True, False), (False, True)])
def test_triangulation(frames, coarse, only_stereo):
    K, fr = frames
    (k1, T1), (k2, T2) = fr[0], fr[2]
    for nodes in (128, 1024):
        fv1, fv2 = featvec(scenario.node_ids(k1.desc, nodes)), featvec(scenario.node_ids(k2.desc, nodes))
        rng = np.random.default_rng(4)
        has1 = (rng.random(k1.n) < 0.4).astype(np.uint8); has2 = (rng.random(k2.n) < 0.4).astype(np.uint8)
        F12, ep = scenario.fundamental(K, T1, T2)
        for check in (True, False):
            rn, rm = OM.ref_search_for_triangulation(k1, k2, fv1, fv2, has1, has2, F12, ep, only_stereo, coarse, check)
            on, om = OM.search_for_triangulation(k1, k2, fv1, fv2, has1, has2, F12, ep, only_stereo, coarse, check)
            assert rn == on and np.array_equal(rm, om)
        if coarse and nodes == 128:
            assert rn > 20


def _fuse_case(frames, seed, with_uright, dup):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    qm, _ = scenario.map_queries(last, cur, K, Tl, Tc, seed=seed)
    z = np.maximum(qm["track_depth"], 0.3).astype(np.float32)
    q = OM.fuse_queries(qm["proj_x"], qm["proj_y"], z, qm["level"], qm["desc"], K["bf"])
    if dup:                                       # several map points landing on the same keypoint: the Replace branches (:1412-1419)
        q = np.concatenate([q, q[::2]]); z = np.concatenate([z, z[::2]])
    kf = cur
    if not with_uright:                           # monocular keyframe: the 5.99 branch of the chi-square gate
        import copy
        kf = copy.copy(cur); kf.uright = None
    return K, kf, q, z


@pytest.mark.parametrize("with_uright,dup,th", [(True, False, 3.0), (True, True, 4.0), (False, False, 3.0), (False, True, 2.5)])
def test_fuse(frames, with_uright, dup, th):
    tab = O.Tables(2000)
    K, kf, q, z = _fuse_case(frames, 7, with_uright, dup)
    n, bi, bd = OM.fuse(kf, q, th, tab.inv_sigma2)
    rn, ridx = OM.ref_fuse(kf, q, z, K["bf"], th, tab.inv_sigma2)
    assert n == rn and np.array_equal(np.where(bd <= 50, bi, -1), ridx)
    assert n > 300


@pytest.mark.parametrize("nodes", [16, 128, 1024])
def test_search_by_bow(frames, nodes):
    """SearchByBoW(KF, F): 16 nodes => ~125 features per node, i.e. long chains of claims inside a node"""
    K, fr = frames
    for a, b in ((0, 1), (0, 2)):
        kf, f = fr[a][0], fr[b][0]
        fvK, fvF = featvec(scenario.node_ids(kf.desc, nodes)), featvec(scenario.node_ids(f.desc, nodes))
        has = (np.random.default_rng(nodes + a + b).random(kf.n) < 0.7).astype(np.uint8)
        for ratio in (0.7, 0.9):
            for check in (True, False):
                n, m = OM.search_by_bow(kf, f, fvK, fvF, has, ratio, check)
                rn, rm = OM.ref_search_by_bow(kf, f, fvK, fvF, has, ratio, check)
                assert n == rn and np.array_equal(m, rm)
        assert n > 100


@pytest.mark.parametrize("th,orb_dist", [(10.0, 100), (3.0, 64), (15.0, 50)])
def test_projection_reloc(frames, th, orb_dist):
    """SearchByProjection(Cur, KF, sAlreadyFound, th, ORBdist): any non-null map point blocks, window [l-1, l+1], no xR gate"""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    q = np.concatenate([q, q[::2]]); q["flags"] = 1; q["invz"] = 1
    claimed = (np.random.default_rng(0).random(cur.n) < 0.2).astype(np.uint8)
    for check in (True, False):
        n, a = OM.search_by_projection_reloc(cur, q, th, orb_dist, check, claimed)
        rn, ra = OM.ref_search_by_projection_reloc(cur, q, th, orb_dist, check, claimed)
        assert n == rn and np.array_equal(a, ra)
    assert n > 300


@pytest.mark.parametrize("dup,th", [(False, 4.0), (True, 3.0), (True, 10.0)])
def test_fuse_sim3(frames, dup, th):
    """Fuse(pKF, Scw, vpPoints, th, vpReplacePoint): no chi-square gate; keypoints that already hold a map point => vpReplacePoint"""
    K, kf, q, z = _fuse_case(frames, 9, True, dup)
    pre = (np.random.default_rng(5).random(kf.n) < 0.3).astype(np.uint8)
    n, bi, bd = OM.fuse_sim3(kf, q, th)
    rn, ridx = OM.ref_fuse_sim3(kf, q, z, th, pre)
    assert n == rn and np.array_equal(np.where(bd <= 50, bi, -1), ridx)
    assert n > 300


@pytest.mark.parametrize("th,ratio", [(8, 1.5), (4, 1.0), (10, 0.9)])
def test_projection_sim3(frames, th, ratio):
    """loop-closing SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming)"""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    q = np.concatenate([q, q[::2]]); q["flags"] = 1; q["invz"] = 1
    matched = (np.random.default_rng(3).random(cur.n) < 0.25).astype(np.uint8)
    n, a = OM.search_by_projection_sim3(cur, q, float(th), ratio, matched)
    rn, ra = OM.ref_search_by_projection_sim3(cur, q, float(th), ratio, matched)
    assert n == rn and np.array_equal(a, ra)
    assert n > 300


def _sim3_case(frames):
    """every keypoint of KF1 carries a map point that projects near its own position in KF2 and vice versa"""
    K, fr = frames
    (k1, T1), (k2, T2) = fr[0], fr[1]
    rng = np.random.default_rng(12)
    def mk(src, dst_frame):
        q = np.zeros(src.n, OM.FUSE_QUERY)
        q["u"] = src.keys["x"] + rng.normal(0, 2.0, src.n).astype(np.float32)
        q["v"] = src.keys["y"] + rng.normal(0, 2.0, src.n).astype(np.float32)
        q["level"] = src.keys["octave"]; q["desc"] = src.desc
        return q
    # KF1's map points carry descriptors of the nearest KF2 features where the scenario matcher finds them, else their own
    q12, q21 = mk(k1, k2), mk(k2, k1)
    has1 = (rng.random(k1.n) < 0.8).astype(np.uint8); has2 = (rng.random(k2.n) < 0.8).astype(np.uint8)
    return k1, k2, q12, q21, has1, has2


def test_search_by_sim3(frames):
    k1, k2, q12, q21, has1, has2 = _sim3_case(frames)
    for th in (7.5, 15.0):
        n, m = OM.search_by_sim3(k1, k2, q12, q21, has1, has2, th)
        rn, rm = OM.ref_search_by_sim3(k1, k2, q12, q21, has1, has2, th)
        assert n == rn and np.array_equal(m, rm)
    assert n > 100


@pytest.mark.parametrize("nodes", [16, 128, 1024])
def test_search_by_bow_kf(frames, nodes):
    """SearchByBoW(KF1, KF2): strict `< TH_LOW`, both sides need a map point, claims in vbMatched2"""
    K, fr = frames
    kf1, kf2 = fr[0][0], fr[2][0]
    fv1, fv2 = featvec(scenario.node_ids(kf1.desc, nodes)), featvec(scenario.node_ids(kf2.desc, nodes))
    rng = np.random.default_rng(nodes)
    has1 = (rng.random(kf1.n) < 0.7).astype(np.uint8); has2 = (rng.random(kf2.n) < 0.7).astype(np.uint8)
    for ratio in (0.8, 0.95):
        for check in (True, False):
            n, m = OM.search_by_bow_kf(kf1, kf2, fv1, fv2, has1, has2, ratio, check)
            rn, rm = OM.ref_search_by_bow_kf(kf1, kf2, fv1, fv2, has1, has2, ratio, check)
            assert n == rn and np.array_equal(m, rm)
    assert n > 50


def test_distinctive_descriptor_against_reference_text():
    """MapPoint::ComputeDistinctiveDescriptors: the numpy restatement (the known-answer check of the CUDA kernel) equals the reference's own
    function body (sliced out of src/MapPoint.cc at build time) on the chosen descriptor"""
    rng = np.random.default_rng(21)
    lists = []
    for n in [1, 2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 64, 100, 150]:
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.repeat(base[None], n, 0).copy()
        d ^= np.packbits(rng.random((n, 256)) < rng.uniform(0.02, 0.3), axis=1)
        lists.append(d)
    lists.append(np.repeat(rng.integers(0, 256, (1, 32), dtype=np.uint8), 6, 0))
    for _ in range(200):
        lists.append(rng.integers(0, 256, (int(rng.integers(2, 20)), 32), dtype=np.uint8))
    for d in lists:
        assert np.array_equal(d[OM.distinctive_descriptor(d)], OM.ref_distinctive_descriptor(d))


def test_stereo_from_rgbd_against_reference_text(frames):
    """Frame::ComputeStereoFromRGBD: scenario.uright_from_depth (what the GPU test compares the kernel with) == the reference's own body"""
    K, fr = frames
    for i, f in enumerate((10, 11, 15)):
        depth = synth.depth_frame(f).copy()
        depth[::7, ::5] = 0.0                                   # invalid readings
        keys = fr[i][0].keys
        ur, dz = scenario.uright_from_depth(keys, depth, K["bf"])
        rur, rdz = OM.ref_stereo_from_rgbd(keys, depth, K["bf"])
        assert np.array_equal(ur.view(np.uint32), rur.view(np.uint32)) and np.array_equal(dz.view(np.uint32), rdz.view(np.uint32))
        assert (ur > 0).sum() > 500 and (ur < 0).sum() > 10


@pytest.mark.parametrize("window,ratio,check", [(100, 0.9, True), (100, 0.9, False), (30, 0.7, True), (10, 0.9, True), (250, 1.0, True)])
def test_search_for_initialization(frames, window, ratio, check):
    """SearchForInitialization: vMatchedDistance gate, stolen matches (vnMatches21), histogram entries of stolen queries, vbPrevMatched update;
    two calls in a row like Tracking::MonocularInitialization (the second starts from the updated positions)"""
    K, fr = frames
    f1, f2, f3 = fr[0][0], fr[1][0], fr[2][0]
    prev = np.stack([f1.keys["x"], f1.keys["y"]], 1)
    n, m, p = OM.search_for_initialization(f1, f2, prev, window, ratio, check)
    rn, rm, rp = OM.ref_search_for_initialization(f1, f2, prev, window, ratio, check)
    assert n == rn and np.array_equal(m, rm) and np.array_equal(p.view(np.uint32), rp.view(np.uint32))
    assert n == (m >= 0).sum() and (f1.keys["octave"][m >= 0] == 0).all() and (f2.keys["octave"][m[m >= 0]] == 0).all()
    if window >= 30:
        assert n > 50
    n2, m2, p2 = OM.search_for_initialization(f1, f3, p, window, ratio, check)
    rn2, rm2, rp2 = OM.ref_search_for_initialization(f1, f3, rp, window, ratio, check)
    assert n2 == rn2 and np.array_equal(m2, rm2) and np.array_equal(p2.view(np.uint32), rp2.view(np.uint32))


def test_search_for_initialization_steals_and_empty():
    """hand-made case: two queries want the same feature, the later and better one takes it over; the loser stays in the rotation histogram"""
    from plvs_b200.matcher import Frame
    from plvs_b200.orb import KP_DTYPE
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, 32, dtype=np.uint8)
    def flip(d, k):
        d = d.copy(); bits = np.unpackbits(d); bits[:k] ^= 1; return np.packbits(bits)
    k1 = np.zeros(4, KP_DTYPE); k1["x"] = [100, 104, 300, 500]; k1["y"] = [100, 100, 300, 50]; k1["octave"] = [0, 0, 0, 1]; k1["angle"] = [10, 200, 10, 10]
    d1 = np.stack([flip(base, 20), flip(base, 3), rng.integers(0, 256, 32, dtype=np.uint8), base])
    k2 = np.zeros(3, KP_DTYPE); k2["x"] = [102, 400, 300]; k2["y"] = [101, 400, 302]; k2["octave"] = [0, 0, 1]; k2["angle"] = [10, 10, 10]
    d2 = np.stack([base, rng.integers(0, 256, 32, dtype=np.uint8), d1[2]])
    tab = O.Tables(1000)
    F1 = Frame(k1, d1, 640, 480, tab.scale, tab.sigma2); F2 = Frame(k2, d2, 640, 480, tab.scale, tab.sigma2)
    prev = np.stack([k1["x"], k1["y"]], 1)
    for check in (False, True):
        n, m, p = OM.search_for_initialization(F1, F2, prev, 20, 0.9, check)
        rn, rm, rp = OM.ref_search_for_initialization(F1, F2, prev, 20, 0.9, check)
        assert n == rn and np.array_equal(m, rm) and np.array_equal(p, rp)
        assert list(m) == [-1, 0, -1, -1] and n == 1                   # query 0 matched first (dist 20), query 1 (dist 3) stole feature 0
        assert tuple(p[1]) == (102.0, 101.0) and tuple(p[0]) == (100.0, 100.0)
    E = Frame(k2[:0], d2[:0], 640, 480, tab.scale, tab.sigma2)
    n, m, p = OM.search_for_initialization(F1, E, prev, 20, 0.9, True)
    rn, rm, rp = OM.ref_search_for_initialization(F1, E, prev, 20, 0.9, True)
    assert n == rn == 0 and np.array_equal(m, rm) and (m == -1).all()
