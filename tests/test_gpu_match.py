"""-m gpu: the three ORBmatcher searches on the GPU against the sequential CPU oracle (bit-exact pairs)."""
import numpy as np
import pytest

from plvs_b200 import synth, scenario
from plvs_b200.orb import ORBextractor
from plvs_b200.matcher import ORBmatcher, Frame, featvec, MP_QUERY, LAST_QUERY
from oracle import match as OM

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames(gpu):
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    K = synth.intrinsics(640, 480)
    out = []
    for f in (10, 11, 15):
        mono, kp, desc = ex(synth.gray_frame(f))
        fr = scenario.make_frame(kp, desc, synth.depth_frame(f), K, ex.GetScaleFactors())
        fr.level_sigma2 = ex.GetScaleSigmaSquares()
        out.append((fr, synth.pose(f)))
    return K, out


def test_hamming_helper(gpu):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8); b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for i in range(200):
        assert ORBmatcher.DescriptorDistance(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())


@pytest.mark.parametrize("th", [1.0, 3.0, 5.0, 15.0])
def test_projection_map(frames, th):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.map_queries(last, cur, K, Tl, Tc)
    m = ORBmatcher(0.8, True)
    n, assign = m.SearchByProjectionMap(cur, q, th)
    on, oassign = OM.search_by_projection_map(cur, q, th, 0.8)
    assert on == n and np.array_equal(assign, oassign)
    if th >= 3:
        assert n > 200


def test_projection_map_claims_and_far(frames):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.map_queries(last, cur, K, Tl, Tc, seed=3)
    rng = np.random.default_rng(1)
    claimed = (rng.random(cur.n) < 0.3).astype(np.uint8)
    q["flags"] = (rng.random(len(q)) < 0.9).astype(np.uint32)          # some map points without observations
    # duplicate queries => heavy competition for the same keypoints (sequential claims matter)
    q = np.concatenate([q, q[::2], q[::3]])
    m = ORBmatcher(0.8, True)
    n, assign = m.SearchByProjectionMap(cur, q, 5.0, True, 3.0, claimed)
    on, oassign = OM.search_by_projection_map(cur, q, 5.0, 0.8, True, 3.0, claimed)
    assert on == n and np.array_equal(assign, oassign)


@pytest.mark.parametrize("th,fwd,bwd", [(15.0, False, False), (7.0, False, False), (15.0, True, False), (15.0, False, True)])
def test_projection_last(frames, th, fwd, bwd):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    for check in (True, False):
        m = ORBmatcher(0.9, check)
        n, assign = m.SearchByProjectionLast(cur, q, th, fwd, bwd)
        on, oassign = OM.search_by_projection_last(cur, q, th, fwd, bwd, check)
        assert on == n and np.array_equal(assign, oassign)
    if not fwd:
        assert n > 300


def test_projection_last_competition(frames):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    q = np.concatenate([q, q[::2]])
    rng = np.random.default_rng(2)
    q["flags"] = (rng.random(len(q)) < 0.8).astype(np.uint32)
    m = ORBmatcher(0.9, True)
    n, assign = m.SearchByProjectionLast(cur, q, 15.0)
    on, oassign = OM.search_by_projection_last(cur, q, 15.0)
    assert on == n and np.array_equal(assign, oassign)


@pytest.mark.parametrize("knobs", [{}, {"PLVS_MATCH_RESOLVE": "cluster"}], ids=["one-cta", "cluster"])
def test_projection_resolve_variants(frames, monkeypatch, knobs):
    """both forms of the claim resolution give the sequential answer under heavy competition (every query duplicated: watch sets overflow, long
    displacement chains, claimants without observations, pre-claimed keypoints): the incremental one-CTA kernel (the default) and the 8-CTA
    cluster kernel that takes over for very large frames"""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    ql, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    ql = np.concatenate([ql, ql[::2]])
    rng = np.random.default_rng(12)
    ql["flags"] = (rng.random(len(ql)) < 0.8).astype(np.uint32)
    claimed = (rng.random(cur.n) < 0.1).astype(np.uint8)
    m = ORBmatcher(0.9, True)
    for rep in range(2):
        n, assign = m.SearchByProjectionLast(cur, ql, 15.0, claimed=claimed)
        on, oassign = OM.search_by_projection_last(cur, ql, 15.0, claimed=claimed)
        assert on == n and np.array_equal(assign, oassign)
    qm, _ = scenario.map_queries(last, cur, K, Tl, Tc, seed=3)
    qm = np.concatenate([qm, qm[::2], qm[::3]])
    qm["flags"] = (rng.random(len(qm)) < 0.85).astype(np.uint32)
    for th in (5.0, 15.0):                # th 15: windows of dozens of candidates
        n, assign = m.SearchByProjectionMap(cur, qm, th, claimed=claimed)
        on, oassign = OM.search_by_projection_map(cur, qm, th, 0.9, claimed=claimed)
        assert on == n and np.array_equal(assign, oassign) and n > 100


@pytest.mark.parametrize("coarse,only_stereo", [(False, False), (True, False), (False, True)])
def test_triangulation(frames, coarse, only_stereo):
    K, fr = frames
    (k1, T1), (k2, T2) = fr[0], fr[2]
    fv1, fv2 = featvec(scenario.node_ids(k1.desc, 128)), featvec(scenario.node_ids(k2.desc, 128))
    rng = np.random.default_rng(4)
    has1 = (rng.random(k1.n) < 0.4).astype(np.uint8); has2 = (rng.random(k2.n) < 0.4).astype(np.uint8)
    F12, ep = scenario.fundamental(K, T1, T2)
    for check in (True, False):
        m = ORBmatcher(0.6, check)
        n, m12 = m.SearchForTriangulation(k1, k2, fv1, fv2, has1, has2, F12, ep, only_stereo, coarse)
        on, om12 = OM.search_for_triangulation(k1, k2, fv1, fv2, has1, has2, F12, ep, only_stereo, coarse, check)
        assert on == n and np.array_equal(m12, om12)
    if coarse:
        assert n > 20


def test_match_empty_inputs(frames):
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    m = ORBmatcher(0.8, True)
    n, assign = m.SearchByProjectionMap(cur, np.zeros(0, MP_QUERY), 3.0)
    assert n == 0 and (assign == -1).all()
    n, assign = m.SearchByProjectionLast(cur, np.zeros(0, LAST_QUERY), 15.0)
    assert n == 0 and (assign == -1).all()


def test_match_device_resident_view(frames):
    """extract -> match without a host round trip of keypoints/descriptors."""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    mono, kp, desc = ex(synth.gray_frame(11))
    assert np.array_equal(kp, cur.keys)
    dv = ex.device_result(0)
    dcur = Frame(None, None, K["w"], K["h"], ex.GetScaleFactors(), bf=K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, 0))
    hcur = Frame(cur.keys, cur.desc, K["w"], K["h"], ex.GetScaleFactors(), bf=K["bf"])
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    m = ORBmatcher(0.9, True)
    n, assign = m.SearchByProjectionLast(dcur, q, 15.0)
    on, oassign = OM.search_by_projection_last(hcur, q, 15.0)
    assert on == n and np.array_equal(assign, oassign)


def test_compute_stereo_matches(gpu):
    """a17 (config 5 geometry: 752x480, 1200 features per eye): row matching + SAD refinement + median cut, bit-exact"""
    from plvs_b200.matcher import ComputeStereoMatches
    from oracle import orb as O
    w, h, nfeat = 752, 480, 1200
    K = synth.intrinsics(w, h)
    baseline = 0.11
    mbf = K["fx"] * baseline
    exl, exr = ORBextractor(nfeat, 1.2, 8, 20, 7), ORBextractor(nfeat, 1.2, 8, 20, 7)
    for frame in (0, 4):
        il, ir = synth.gray_frame(frame, w, h), synth.gray_frame(frame, w, h, eye=baseline)
        _, kl, dl = exl(il); _, kr, dr = exr(ir)
        sf, isf = exl.GetScaleFactors(), exl.GetInverseScaleFactors()
        L = Frame(kl, dl, w, h, sf, bf=mbf); R = Frame(kr, dr, w, h, sf, bf=mbf)
        m = ORBmatcher(0.8, True)
        ur, dp, kept = ComputeStereoMatches(m, L, R, exl.pyramid_view(0), exr.pyramid_view(0), isf, baseline, mbf)
        pl = [exl.pyramid_level(l) for l in range(8)]; pr = [exr.pyramid_level(l) for l in range(8)]
        our, odp, okept = OM.compute_stereo_matches(L, R, pl, pr, sf, isf, baseline, mbf)
        assert kept == okept
        assert np.array_equal(ur.view(np.uint32), our.view(np.uint32)) and np.array_equal(dp.view(np.uint32), odp.view(np.uint32))
        assert kept > 300
        # sanity: recovered depth agrees with the rendered scene where both exist
        z = synth.depth_frame(frame, w, h, noise=False)
        ok = dp > 0
        zt = z[kl["y"][ok].astype(int), kl["x"][ok].astype(int)]
        good = zt > 0
        assert np.median(np.abs(dp[ok][good] - zt[good]) / zt[good]) < 0.03


def test_grid_cache_key_reuses_grid_and_stays_exact(frames):
    """two searches on the same device-resident frame through one handle: the second reuses the cached feature grid"""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    mono, kp, desc = ex(synth.gray_frame(11))
    dv = ex.device_result(0)
    assert dv.cache_key != 0
    dcur = Frame(None, None, K["w"], K["h"], ex.GetScaleFactors(), bf=K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, 0, dv.cache_key))
    hcur = Frame(cur.keys, cur.desc, K["w"], K["h"], ex.GetScaleFactors(), bf=K["bf"])
    ql, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    qm, _ = scenario.map_queries(last, cur, K, Tl, Tc)
    m = ORBmatcher(0.9, True)
    n1, a1 = m.SearchByProjectionLast(dcur, ql, 15.0)
    claimed = (a1 >= 0).astype(np.uint8)
    n2, a2 = m.SearchByProjectionMap(dcur, qm, 3.0, claimed=claimed, nnratio=0.8)
    on1, oa1 = OM.search_by_projection_last(hcur, ql, 15.0)
    on2, oa2 = OM.search_by_projection_map(hcur, qm, 3.0, 0.8, claimed=(oa1 >= 0).astype(np.uint8))
    assert (n1, n2) == (on1, on2) and np.array_equal(a1, oa1) and np.array_equal(a2, oa2)
    # a new extraction gets a new key: no stale grid
    mono, kp, desc = ex(synth.gray_frame(12))
    assert ex.device_result(0).cache_key != dv.cache_key


@pytest.mark.parametrize("with_uright,dup,th", [(True, False, 3.0), (True, True, 4.0), (False, False, 3.0), (False, True, 2.5)])
def test_fuse(frames, with_uright, dup, th):
    """search part of ORBmatcher::Fuse (a "next" row of SURVEY.md §8f rank 1): CUDA == oracle (bit-exact best index and
    distance per map point) == the reference's own Fuse where it is observable (fused pairs)"""
    import copy
    from oracle import orb as O
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    qm, _ = scenario.map_queries(last, cur, K, Tl, Tc, seed=7)
    z = np.maximum(qm["track_depth"], 0.3).astype(np.float32)
    q = OM.fuse_queries(qm["proj_x"], qm["proj_y"], z, qm["level"], qm["desc"], K["bf"])
    if dup:
        q = np.concatenate([q, q[::2]]); z = np.concatenate([z, z[::2]])
    kf = cur
    if not with_uright:
        kf = copy.copy(cur); kf.uright = None
    inv = O.Tables(2000).inv_sigma2
    n, bi, bd = ORBmatcher(0.6, True).Fuse(kf, q, th, inv)
    on, obi, obd = OM.fuse(kf, q, th, inv)
    assert n == on and np.array_equal(bi, obi) and np.array_equal(bd, obd)
    assert n > 300
    if OM.ref_available():
        rn, ridx = OM.ref_fuse(kf, q, z, K["bf"], th, inv)
        assert n == rn and np.array_equal(np.where(bd <= 50, bi, -1), ridx)
    n0, bi0, bd0 = ORBmatcher(0.6, True).Fuse(kf, q[:0], th, inv)
    assert n0 == 0 and len(bi0) == 0


@pytest.mark.parametrize("nodes", [16, 128, 1024])
def test_search_by_bow(frames, nodes):
    """ORBmatcher::SearchByBoW(KF, F) (§8f rank 1): CUDA == oracle == the reference's own ORBmatcher.cc"""
    K, fr = frames
    for a, b in ((0, 1), (0, 2)):
        kf, f = fr[a][0], fr[b][0]
        fvK, fvF = featvec(scenario.node_ids(kf.desc, nodes)), featvec(scenario.node_ids(f.desc, nodes))
        has = (np.random.default_rng(nodes + a + b).random(kf.n) < 0.7).astype(np.uint8)
        for ratio in (0.7, 0.9):
            for check in (True, False):
                n, m = ORBmatcher(ratio, check).SearchByBoW(kf, f, fvK, fvF, has)
                on, om = OM.search_by_bow(kf, f, fvK, fvF, has, ratio, check)
                assert n == on and np.array_equal(m, om)
                if OM.ref_available():
                    rn, rm = OM.ref_search_by_bow(kf, f, fvK, fvF, has, ratio, check)
                    assert n == rn and np.array_equal(m, rm)
        assert n > 100


@pytest.mark.parametrize("th,orb_dist", [(10.0, 100), (3.0, 64), (15.0, 50)])
def test_projection_reloc(frames, th, orb_dist):
    """relocalisation overload SearchByProjection(Cur, KF, sAlreadyFound, th, ORBdist) (§8f rank 1)"""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    q = np.concatenate([q, q[::2]]); q["flags"] = 1; q["invz"] = 1
    claimed = (np.random.default_rng(0).random(cur.n) < 0.2).astype(np.uint8)
    for check in (True, False):
        n, a = ORBmatcher(0.9, check).SearchByProjectionReloc(cur, q, th, orb_dist, claimed)
        on, oa = OM.search_by_projection_reloc(cur, q, th, orb_dist, check, claimed)
        assert n == on and np.array_equal(a, oa)
        if OM.ref_available():
            rn, ra = OM.ref_search_by_projection_reloc(cur, q, th, orb_dist, check, claimed)
            assert n == rn and np.array_equal(a, ra)
    assert n > 300


@pytest.mark.parametrize("dup,th", [(False, 4.0), (True, 3.0), (True, 10.0)])
def test_fuse_sim3(frames, dup, th):
    """loop-closing overload Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (§8f rank 1)"""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    qm, _ = scenario.map_queries(last, cur, K, Tl, Tc, seed=9)
    z = np.maximum(qm["track_depth"], 0.3).astype(np.float32)
    q = OM.fuse_queries(qm["proj_x"], qm["proj_y"], z, qm["level"], qm["desc"], K["bf"])
    if dup:
        q = np.concatenate([q, q[::2]]); z = np.concatenate([z, z[::2]])
    n, bi, bd = ORBmatcher(0.8, True).FuseSim3(cur, q, th)
    on, obi, obd = OM.fuse_sim3(cur, q, th)
    assert n == on and np.array_equal(bi, obi) and np.array_equal(bd, obd)
    if OM.ref_available():
        pre = (np.random.default_rng(5).random(cur.n) < 0.3).astype(np.uint8)
        rn, ridx = OM.ref_fuse_sim3(cur, q, z, th, pre)
        assert n == rn and np.array_equal(np.where(bd <= 50, bi, -1), ridx)


@pytest.mark.parametrize("th,ratio", [(8, 1.5), (4, 1.0), (10, 0.9)])
def test_projection_sim3(frames, th, ratio):
    """loop-closing SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (§8f rank 1)"""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    q = np.concatenate([q, q[::2]]); q["flags"] = 1; q["invz"] = 1
    matched = (np.random.default_rng(3).random(cur.n) < 0.25).astype(np.uint8)
    n, a = ORBmatcher(0.75, True).SearchByProjectionSim3(cur, q, float(th), ratio, matched)
    on, oa = OM.search_by_projection_sim3(cur, q, float(th), ratio, matched)
    assert n == on and np.array_equal(a, oa)
    if OM.ref_available():
        rn, ra = OM.ref_search_by_projection_sim3(cur, q, float(th), ratio, matched)
        assert n == rn and np.array_equal(a, ra)


def test_search_by_sim3(frames):
    """SearchBySim3 = two device searches (the Fuse-Sim3 search) + the mutual-agreement check (§8f rank 1)"""
    K, fr = frames
    (k1, T1), (k2, T2) = fr[0], fr[1]
    rng = np.random.default_rng(12)
    def mk(src):
        q = np.zeros(src.n, OM.FUSE_QUERY)
        q["u"] = src.keys["x"] + rng.normal(0, 2.0, src.n).astype(np.float32)
        q["v"] = src.keys["y"] + rng.normal(0, 2.0, src.n).astype(np.float32)
        q["level"] = src.keys["octave"]; q["desc"] = src.desc
        return q
    q12, q21 = mk(k1), mk(k2)
    has1 = (rng.random(k1.n) < 0.8).astype(np.uint8); has2 = (rng.random(k2.n) < 0.8).astype(np.uint8)
    for th in (7.5, 15.0):
        n, m = ORBmatcher(0.75, True).SearchBySim3(k1, k2, q12, q21, has1, has2, th)
        on, om = OM.search_by_sim3(k1, k2, q12, q21, has1, has2, th)
        assert n == on and np.array_equal(m, om)
        if OM.ref_available():
            rn, rm = OM.ref_search_by_sim3(k1, k2, q12, q21, has1, has2, th)
            assert n == rn and np.array_equal(m, rm)
    assert n > 100


@pytest.mark.parametrize("nodes", [16, 128, 1024])
def test_search_by_bow_kf(frames, nodes):
    """ORBmatcher::SearchByBoW(KF1, KF2) (§8f rank 1)"""
    K, fr = frames
    kf1, kf2 = fr[0][0], fr[2][0]
    fv1, fv2 = featvec(scenario.node_ids(kf1.desc, nodes)), featvec(scenario.node_ids(kf2.desc, nodes))
    rng = np.random.default_rng(nodes)
    has1 = (rng.random(kf1.n) < 0.7).astype(np.uint8); has2 = (rng.random(kf2.n) < 0.7).astype(np.uint8)
    for ratio in (0.8, 0.95):
        for check in (True, False):
            n, m = ORBmatcher(ratio, check).SearchByBoWKF(kf1, kf2, fv1, fv2, has1, has2)
            on, om = OM.search_by_bow_kf(kf1, kf2, fv1, fv2, has1, has2, ratio, check)
            assert n == on and np.array_equal(m, om)
            if OM.ref_available():
                rn, rm = OM.ref_search_by_bow_kf(kf1, kf2, fv1, fv2, has1, has2, ratio, check)
                assert n == rn and np.array_equal(m, rm)
    assert n > 50


def test_distinctive_descriptors(gpu):
    """MapPoint::ComputeDistinctiveDescriptors, batched (§8f rank 1): known-answer check against an independent numpy restatement"""
    rng = np.random.default_rng(21)
    lists = []
    for n in [1, 2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 64, 100, 0, 150]:
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.repeat(base[None], n, 0).copy()
        flips = rng.random((n, 256)) < rng.uniform(0.02, 0.3)                  # noisy copies of one descriptor, some outliers
        d ^= np.packbits(flips, axis=1)
        lists.append(d)
    lists.append(np.repeat(rng.integers(0, 256, (1, 32), dtype=np.uint8), 6, 0))  # all identical: every median 0, first wins
    for _ in range(200):
        n = int(rng.integers(2, 20))
        lists.append(rng.integers(0, 256, (n, 32), dtype=np.uint8))
    got = ORBmatcher(0.6, True).ComputeDistinctiveDescriptors(lists)
    want = np.array([OM.distinctive_descriptor(d) for d in lists], np.int32)
    assert np.array_equal(got, want)


def test_device_view_with_host_uright(frames):
    """keypoints / descriptors device-resident from the extractor, mvuRight from the caller's host array (PLVS_VIEW_URIGHT_ON_HOST):
    the right-coordinate gate is applied exactly as with an all-host view"""
    K, fr = frames
    (last, Tl), (cur, Tc) = fr[0], fr[1]
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    mono, kp, desc = ex(synth.gray_frame(11))
    assert np.array_equal(kp, cur.keys)
    dv = ex.device_result(0)
    dcur = Frame(None, None, K["w"], K["h"], ex.GetScaleFactors(), uright=cur.uright, bf=K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, 0, dv.cache_key))
    q, _ = scenario.last_queries(last, cur, K, Tl, Tc)
    n, assign = ORBmatcher(0.9, True).SearchByProjectionLast(dcur, q, 15.0)
    on, oassign = OM.search_by_projection_last(cur, q, 15.0)
    assert on == n and np.array_equal(assign, oassign)
    nogate = Frame(cur.keys, cur.desc, K["w"], K["h"], ex.GetScaleFactors(), bf=K["bf"])
    n2, assign2 = OM.search_by_projection_last(nogate, q, 15.0)
    assert not np.array_equal(assign2, oassign)                 # the gate really changes the result on this stream
    qm, _ = scenario.map_queries(last, cur, K, Tl, Tc)
    n3, a3 = ORBmatcher(0.8, True).SearchByProjectionMap(dcur, qm, 3.0)
    on3, oa3 = OM.search_by_projection_map(cur, qm, 3.0, 0.8)
    assert n3 == on3 and np.array_equal(a3, oa3)
