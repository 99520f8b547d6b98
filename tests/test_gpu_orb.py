"""-m gpu: the CUDA ORB extractor against the CPU oracle, stage by stage and end to end (bit-exact)."""
import numpy as np
import pytest

from plvs_b200 import synth
from plvs_b200.orb import ORBextractor
from oracle import orb as O

pytestmark = pytest.mark.gpu

FIELDS = ("x", "y", "size", "angle", "response", "octave", "class_id")


def _compare(img, nfeat, ex=None, lap=(0, 0)):
    ex = ex or ORBextractor(nfeat, 1.2, 8, 20, 7)
    mono, kp, desc = ex(img, None, lap)
    okp, odesc, omono, ncand, internals = O.extract_cv2(img, nfeat, lapping=lap, return_internals=True, angle_impl="c")
    for l in range(8):
        assert np.array_equal(ex.pyramid_level(l), internals["pyramid"][l]), f"pyramid level {l}"
        if internals["blurred"][l] is not None:
            assert np.array_equal(ex.pyramid_level(l, blurred=True), internals["blurred"][l]), f"blur level {l}"
    for l in range(8):
        cx, cy, cs = ex.candidates(l)
        ox, oy, orr = internals["candidates"][l]
        if not (len(cx) == len(ox) and np.array_equal(cx - 16, ox.astype(np.int32)) and np.array_equal(cy - 16, oy.astype(np.int32))
                and np.array_equal(cs, orr.astype(np.int32))):
            import collections
            msg = f"level {l}: gpu {len(cx)} vs oracle {len(ox)} candidates"
            gs = set(zip((cx - 16).tolist(), (cy - 16).tolist(), cs.tolist())); os_ = set(zip(ox.astype(int).tolist(), oy.astype(int).tolist(), orr.astype(int).tolist()))
            msg += f"; only-gpu {sorted(gs - os_)[:12]} only-oracle {sorted(os_ - gs)[:12]} n_only_gpu {len(gs-os_)} n_only_oracle {len(os_-gs)}"
            raise AssertionError(msg)
    assert ex.last_stats()["candidates"] == ncand
    assert len(kp) == len(okp)
    for f in FIELDS:
        assert np.array_equal(kp[f], okp[f]), f"keypoint field {f}"
    assert np.array_equal(desc, odesc)
    assert mono == omono
    return len(kp)


def test_orb_score_map(gpu, monkeypatch):
    """inspection build (PLVS_ORB_DEBUG=1): FAST score map of every level vs the oracle's"""
    monkeypatch.setenv("PLVS_ORB_DEBUG", "1")
    img = synth.gray_frame(0)
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    ex(img)
    tab = O.Tables(2000)
    pyr = O.pyramid_cv2(img, tab)
    for l in range(8):
        got = ex.pyramid_level(l, blurred=2)
        want = O.fast_score_map(pyr[l], 7)
        h, w = want.shape
        bad = np.argwhere(got[19:h - 19, 19:w - 19] != want[19:h - 19, 19:w - 19])
        assert len(bad) == 0, f"level {l}: {len(bad)} score mismatches, first {bad[:5] + 19}, got {[int(got[y + 19, x + 19]) for y, x in bad[:5]]} want {[int(want[y + 19, x + 19]) for y, x in bad[:5]]}"


def test_orb_vga_2000(gpu):
    assert _compare(synth.gray_frame(0), 2000) > 1500


def test_orb_vga_1000_frames(gpu):
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    for f in (1, 7):
        _compare(synth.gray_frame(f), 1000, ex)


def test_orb_lapping_area(gpu):
    _compare(synth.gray_frame(3), 1000, lap=(200, 400))


def test_orb_low_texture_fallback(gpu):
    # smooth image: almost every cell needs the minThFAST pass, many cells stay empty
    img = (synth.gray_frame(2).astype(np.float32) * 0.15 + 100).astype(np.uint8)
    _compare(img, 1000)


def test_orb_flat_and_empty(gpu):
    ex = ORBextractor(500, 1.2, 8, 20, 7)
    mono, kp, desc = ex(np.full((480, 640), 127, np.uint8))
    assert mono == 0 and len(kp) == 0
    assert ex(np.zeros((0, 0), np.uint8))[0] == -1


def test_orb_odd_sizes(gpu):
    rng = np.random.default_rng(5)
    for (w, h) in ((752, 480), (333, 257)):
        img = synth.gray_frame(0, 1024, 768)[:h, :w].copy()
        _compare(img, 1200)


def test_orb_batch_matches_single(gpu):
    ex = ORBextractor(2000, 1.2, 8, 20, 7)
    imgs = np.stack([synth.gray_frame(f) for f in range(4)])
    mono, kps, descs = ex.extract_batch(imgs)
    for b in range(4):
        m1, k1, d1 = ORBextractor(2000, 1.2, 8, 20, 7)(imgs[b])
        assert m1 == mono[b] and np.array_equal(k1, kps[b]) and np.array_equal(d1, descs[b])


def test_orb_1080p(gpu):
    img = synth.gray_frame(0, 1920, 1080)
    ex = ORBextractor(4000, 1.2, 8, 20, 7)
    mono, kp, desc = ex(img)
    okp, odesc, omono, ncand = O.extract_port(img, 4000)
    assert len(kp) == len(okp) and mono == omono
    for f in FIELDS:
        assert np.array_equal(kp[f], okp[f]), f
    assert np.array_equal(desc, odesc)


@pytest.mark.parametrize("nfeat", [300, 1000, 2000, 5000])
def test_device_distributor_equals_host_distributor(gpu, monkeypatch, nfeat):
    """DistributeOctTree on the device (default) vs the host implementation (PLVS_ORB_HOST_DISTRIBUTE=1): same keypoints, same order"""
    imgs = [synth.gray_frame(5), (synth.gray_frame(6).astype(np.float32) * 0.2 + 80).astype(np.uint8), synth.gray_frame(0, 333, 257)]
    dev = ORBextractor(nfeat, 1.2, 8, 20, 7)
    monkeypatch.setenv("PLVS_ORB_HOST_DISTRIBUTE", "1")
    host = ORBextractor(nfeat, 1.2, 8, 20, 7)
    monkeypatch.delenv("PLVS_ORB_HOST_DISTRIBUTE")
    for img in imgs:
        m1, k1, d1 = dev(img)
        m2, k2, d2 = host(img)
        assert m1 == m2 and len(k1) == len(k2)
        assert np.array_equal(k1, k2) and np.array_equal(d1, d2)


@pytest.mark.parametrize("arena_nodes", [0, 24, 200])
def test_distributor_state_in_shared_memory_and_in_global_memory(gpu, monkeypatch, arena_nodes):
    """single-frame calls keep the distributor's node-level state in shared memory; a level that outgrows the arena starts over with the state
    in global memory (forced here with a tiny arena: 24 nodes overflow on every level, 200 on the lower ones); batches use global memory.  All the same keypoints."""
    imgs = [synth.gray_frame(3), synth.gray_frame(0, 333, 257)]
    want = [ORBextractor(2000, 1.2, 8, 20, 7).extract_batch(np.stack([im, im]))[1][0] for im in imgs]      # batch of two: global-memory state
    if arena_nodes:
        monkeypatch.setenv("PLVS_ORB_DIST_ARENA_NODES", str(arena_nodes))
    for im, w in zip(imgs, want):
        ex = ORBextractor(2000, 1.2, 8, 20, 7)
        mono, kp, desc = ex(im)
        assert np.array_equal(kp, w)
        okp = O.extract_port(im, 2000)[0]
        assert np.array_equal(kp, okp)


def test_cuda_vs_compiled_reference_live(gpu):
    """the CUDA extractor against the REFERENCE's own ORBextractor.cc (oracle/_ref/liborb_ref.so, prebuilt in the build
    container by oracle/ref_build.py; it travels to the GPU box with the snapshot)"""
    if not O.ref_available():
        pytest.skip("oracle/_ref/liborb_ref.so did not travel to this box")
    for (w, h, nfeat, frame, lap) in [(640, 480, 2000, 5, (0, 0)), (752, 480, 1200, 1, (0, 0)), (640, 480, 1000, 2, (0, 1000))]:
        img = synth.gray_frame(frame, w, h)
        ex = ORBextractor(nfeat, 1.2, 8, 20, 7)
        mono, kp, desc = ex(img, None, lap)
        ref = O.RefExtractor(nfeat)
        rkp, rdesc, rmono = ref(img, lap)
        assert mono == rmono and len(kp) == len(rkp)
        for f in FIELDS:
            assert np.array_equal(kp[f], rkp[f]), f
        assert np.array_equal(desc, rdesc)
        for l in range(8):
            assert np.array_equal(ex.pyramid_level(l), ref.level(l))


def test_color_input_and_stereo_from_rgbd(gpu):
    """steps either side of the extraction (§8f rank 2): cvtColor fused in front of the pyramid (== the real cv2.cvtColor), and
    Frame::ComputeStereoFromRGBD on the device-resident keypoints feeding the matcher's right-coordinate gate"""
    import cv2
    from plvs_b200 import scenario
    from plvs_b200.matcher import ORBmatcher, Frame
    from oracle import match as OM
    w, h = 640, 480
    K = synth.intrinsics(w, h)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    for nch, rgb, code in ((3, False, cv2.COLOR_BGR2GRAY), (3, True, cv2.COLOR_RGB2GRAY), (4, False, cv2.COLOR_BGRA2GRAY)):
        rng = np.random.default_rng(nch + int(rgb))
        gray0 = synth.gray_frame(7, w, h)
        col = np.stack([gray0, np.roll(gray0, 5, 1), (255 - gray0)] + ([gray0] if nch == 4 else []), -1)
        col = np.ascontiguousarray(np.clip(col.astype(np.int16) + rng.integers(-20, 20, col.shape), 0, 255).astype(np.uint8))
        want_gray = cv2.cvtColor(col, code)
        assert np.array_equal(want_gray, O.color_to_gray(col, rgb))
        mono, kps, descs = ex.extract_batch_color(col[None], rgb=rgb)
        assert np.array_equal(ex.pyramid_level(0), want_gray)
        mono2, kp2, desc2 = ORBextractor(1000, 1.2, 8, 20, 7)(want_gray)
        assert mono[0] == mono2 and np.array_equal(kps[0], kp2) and np.array_equal(descs[0], desc2)
    # ComputeStereoFromRGBD on the keypoints that are still on the device
    img = synth.gray_frame(11, w, h); depth = synth.depth_frame(11, w, h)
    mono, kp, desc = ex(img)
    ur, dz, dptr = ex.ComputeStereoFromRGBD(depth, K["bf"])
    our, odz = scenario.uright_from_depth(kp, depth, K["bf"])
    assert np.array_equal(ur, our) and np.array_equal(dz, odz) and (ur > 0).sum() > 500
    # ... and the device copy drives the matcher's xR gate exactly like the host array
    kp0, desc0, _, _ = O.extract_port(synth.gray_frame(10, w, h), 1000)
    tab = O.Tables(1000)
    last = scenario.make_frame(kp0, desc0, synth.depth_frame(10, w, h), K, tab.scale)
    cur = scenario.make_frame(kp, desc, depth, K, tab.scale)
    q, _ = scenario.last_queries(last, cur, K, synth.pose(10), synth.pose(11))
    dv = ex.device_result(0)
    dcur = Frame(None, None, w, h, ex.GetScaleFactors(), bf=K["bf"], device_ptrs=(dv.n, dv.keys, dv.desc, dptr, dv.cache_key))
    n, assign = ORBmatcher(0.9, True).SearchByProjectionLast(dcur, q, 15.0)
    on, oassign = OM.search_by_projection_last(cur, q, 15.0)
    assert n == on and np.array_equal(assign, oassign)
