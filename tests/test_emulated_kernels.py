"""CPU: the product's own kernel source (plvs_b200/csrc/*.cuh, device-only headers) executed on the small CPU model of CUDA in
tests/native/cuda_emu.hpp -- every CUDA thread a fiber, barriers / shuffles / ballots as rendezvous points -- and compared with the oracle.
It exists for the kernels written after the round's GPU budget was spent (tests/test_gpu_widened.py holds their GPU tests): index
arithmetic, warp-level reductions, prefix sums, barrier placement and float operation order are exercised here with the real text; what a CPU
model cannot show (memory-ordering races, launch configuration limits, nvcc code generation) stays for the GPU run.  k_build_grid and
k_in_frustum already passed on a B200 and double as a check of the model itself."""
import ctypes as C
import pathlib
import numpy as np
import pytest

from plvs_b200 import synth, scenario, tsdf as T, _lib as ABI
from plvs_b200.matcher import Frame
from plvs_b200.orb import KP_DTYPE
from oracle import match as OM, orb as O, tsdf as OT
from tests.native_build import build_emulated_kernels


ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emu():
    return C.CDLL(build_emulated_kernels())


@pytest.fixture(scope="module")
def frames():
    K = synth.intrinsics(640, 480)
    tab = O.Tables(3000)
    out = []
    for f in (10, 11, 14):
        kp, desc, _, _ = O.extract_port(synth.gray_frame(f), 3000)
        out.append(scenario.make_frame(kp, desc, synth.depth_frame(f), K, tab.scale))
    return K, tab, out


def test_model_on_a_verified_kernel_build_grid(emu, frames):
    """k_build_grid (1024 threads: shared-memory histogram, two-level shuffle scan, atomics, per-cell insertion sort) passed on the B200;
    here it has to reproduce Frame::AssignFeaturesToGrid's cell lists on the CPU model"""
    _, _, fr = frames
    f = fr[0]
    v = f.view()
    cs = np.zeros(64 * 48 + 1, np.int32); srt = np.zeros(f.n, np.int32)
    assert emu.emu_build_grid(C.byref(v), cs.ctypes.data_as(C.c_void_p), srt.ctypes.data_as(C.c_void_p)) == 0
    def c_round(v):                                            # C round(): halves away from zero (numpy rounds them to even)
        v = v.astype(np.float64)
        return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(int)
    px = c_round((f.keys["x"] - np.float32(f.min_x)) * np.float32(f.grid_inv_w))
    py = c_round((f.keys["y"] - np.float32(f.min_y)) * np.float32(f.grid_inv_h))
    ok = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    cell = px * 48 + py
    want = [np.nonzero(ok & (cell == c))[0] for c in range(64 * 48)]
    assert cs[-1] == ok.sum()
    for c in range(64 * 48):
        assert np.array_equal(srt[cs[c]:cs[c + 1]], want[c])


@pytest.mark.parametrize("window,ratio,check,cap", [(100, 0.9, False, 128), (30, 0.7, True, 16), (10, 0.9, True, 4)])
def test_search_for_initialization_kernels(emu, frames, window, ratio, check, cap):
    """k_init_candidates + k_init_resolve == the oracle (pinned to the reference's compiled SearchForInitialization): matches, count, updated
    vbPrevMatched; small initial caps force the redo-with-room loop of the entry point"""
    _, _, fr = frames
    f1, f2, f3 = fr
    prev = np.stack([f1.keys["x"], f1.keys["y"]], 1).astype(np.float32)
    emu.emu_match_initialization.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    for other in (f2, f3):
        v1, v2 = f1.view(), other.view()
        p = prev.copy(); m = np.full(f1.n, -1, np.int32); nm = C.c_int()
        assert emu.emu_match_initialization(C.byref(v1), C.byref(v2), p.ctypes.data_as(C.c_void_p), window, ratio, int(check), m.ctypes.data_as(C.c_void_p),
                                            C.byref(nm), cap) == 0
        on, om, op = OM.search_for_initialization(f1, other, prev, window, ratio, check)
        assert nm.value == on and np.array_equal(m, om) and np.array_equal(p.view(np.uint32), op.view(np.uint32))
        if window >= 30:
            assert on > 50
        prev = op


def test_in_frustum_and_compaction_kernels(emu, frames):
    """k_in_frustum (already green on the B200) and k_compact_queries (not yet): the in-view queries, in order, with their source indices;
    n is not a multiple of 1024 and spans several chunks of the block-wide scan"""
    K, tab, fr = frames
    last, cur = fr[0], fr[1]
    Tl, Tc = synth.pose(10), synth.pose(11)
    ok = last.depth_at_kp > 0
    Pw = scenario.backproject(last.keys[ok], last.depth_at_kp[ok], K, Tl)
    Pw = np.concatenate([Pw, Pw + np.float32(0.01), Pw[::-1] * np.float32(1.5)])           # ~ 3 x 1500 points, some out of view
    n = len(Pw)
    rng = np.random.default_rng(2)
    pts = np.zeros(n, OM.MAP_POINT)
    pts["xw"] = Pw
    Ow = np.asarray(Tl, np.float64).reshape(3, 4)[:, 3]
    v = Pw.astype(np.float64) - Ow; d = np.linalg.norm(v, axis=1)
    pts["normal"] = (v / d[:, None]).astype(np.float32)
    pts["max_dist"] = (d * rng.uniform(0.8, 2.0, n)).astype(np.float32); pts["min_dist"] = (pts["max_dist"] / tab.scale[7]).astype(np.float32)
    pts["flags"] = (rng.random(n) < 0.9).astype(np.uint32); pts["desc"] = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    frm = OM.make_frustum(Tc, K, (0.0, 0.0, 640.0, 480.0), K["bf"], 0.5, 1.2, 8)
    Tt = np.zeros(16, np.float32); Tt[:7] = OM.scale_thresholds(1.2, 8)
    q = np.zeros(n, OM.MP_QUERY); iv = np.zeros(n, np.uint8); cq = np.zeros(n, OM.MP_QUERY); src = np.full(n, -1, np.int32); cnt = C.c_int()
    assert emu.emu_in_frustum(C.byref(frm), Tt.ctypes.data_as(C.c_void_p), pts.ctypes.data_as(C.c_void_p), n, q.ctypes.data_as(C.c_void_p),
                              iv.ctypes.data_as(C.c_void_p), C.byref(cnt), cq.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p)) == 0
    on, oq, oiv = OM.in_frustum(frm, pts)
    assert cnt.value == on and np.array_equal(iv, oiv) and np.array_equal(q.tobytes(), oq.tobytes())
    assert 1000 < on < n - 100 and n > 3 * 1024
    idx = np.nonzero(oiv)[0]
    assert np.array_equal(src[:on], idx) and np.array_equal(cq[:on].tobytes(), oq[idx].tobytes())


def test_undistort_kernel(emu, frames):
    cv2 = pytest.importorskip("cv2")
    _, _, fr = frames
    kp = fr[0].keys
    xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float32)
    emu.emu_undistort.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    for K4, dist in (((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
                     ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
                     ((500.0, 500.0, 320.0, 240.0), (0.1, -0.2, 0.001, -0.002, 0.05, 0.01, -0.02, 0.003))):
        K4f = [float(np.float32(v)) for v in K4]                       # the camera matrix and the coefficients are CV_32F in PLVS (src/Frame.cc:1521-1527)
        k14 = np.zeros(14, np.float64); k14[:len(dist)] = np.array(dist, np.float32)
        out = np.zeros(len(kp), KP_DTYPE)
        assert emu.emu_undistort(kp.ctypes.data_as(C.c_void_p), len(kp), *K4f, k14.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        want = O.undistort_points(xy, K4, np.array(dist, np.float32))
        Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
        cvw = cv2.undistortPoints(xy.reshape(-1, 1, 2), Km, np.array(dist, np.float32), None, Km).reshape(-1, 2)
        assert np.array_equal(want.view(np.uint32), cvw.view(np.uint32))
        assert np.array_equal(out["x"].view(np.uint32), want[:, 0].view(np.uint32)) and np.array_equal(out["y"].view(np.uint32), want[:, 1].view(np.uint32))
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(out[f], kp[f])


@pytest.mark.parametrize("color,res,far", [(1, 0.04, 4.0), (0, 0.04, 4.0), (1, 0.5, 6.0)])
def test_mesh_kernels(emu, color, res, far):
    """k_mesh_count / k_mesh_emit / k_mesh_shade over the oracle's voxel blocks (handed over in a shuffled pool order) == the oracle's meshes, which
    are pinned bit-exactly to the compiled open_chisel: chunk order, vertex order, positions, normals, colours"""
    K = synth.intrinsics(160, 120)
    p = T.default_params(voxel_resolution=res, use_carving=1, near_plane=0.1, far_plane=far, max_blocks=8192, use_color=color)
    o = OT.Map(p, threads=8); o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], 160, 120)
    for f in (0, 1, 2, 6):
        o.integrate(synth.depth_frame(f, 160, 120), synth.pose(f), synth.bgr_frame(f, 160, 120) if color else None)
    keys, sdf, w, rgba = o.download()
    perm = np.random.default_rng(0).permutation(len(keys))
    keys, sdf, w, rgba = (np.ascontiguousarray(a[perm]) for a in (keys, sdf, w, rgba))
    emu.emu_update_meshes.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_longlong, C.c_void_p]
    ok, oc, oV, oN, oC = o.extract_mesh()
    nb = len(keys)
    mk = np.zeros((nb, 3), np.int32); mc = np.zeros(nb, np.int32); nv = C.c_longlong()
    V = np.zeros((len(oV) + 8, 3), np.float32); N = np.zeros_like(V); Cc = np.zeros_like(V)
    nm = emu.emu_update_meshes(nb, keys.ctypes.data_as(C.c_void_p), sdf.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), rgba.ctypes.data_as(C.c_void_p),
                               res, color, mk.ctypes.data_as(C.c_void_p), mc.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), N.ctypes.data_as(C.c_void_p),
                               Cc.ctypes.data_as(C.c_void_p), len(V), C.byref(nv))
    assert nm == len(ok) and nv.value == len(oV) and nm > 3
    assert np.array_equal(mk[:nm], ok) and np.array_equal(mc[:nm], oc)
    for a, b in ((V, oV), (N, oN), (Cc, oC)):
        assert np.array_equal(a[:len(oV)].view(np.uint32), b.view(np.uint32))


def _sibling_layout(parent, desc):
    """what plvs_voc_create builds on the host: children in id order, their descriptors contiguous"""
    n = len(parent)
    off = np.zeros(n + 1, np.int32)
    for i in range(1, n):
        off[parent[i] + 1] += 1
    off = np.cumsum(off).astype(np.int32)
    cur = off[:-1].copy()
    cid = np.zeros(max(n - 1, 1), np.int32); cdesc = np.zeros((max(n - 1, 1), 32), np.uint8)
    for i in range(1, n):
        s = cur[parent[i]]; cur[parent[i]] += 1
        cid[s] = i; cdesc[s] = desc[i]
    return off, cid, cdesc


@pytest.mark.parametrize("k,L,levelsup,zero", [(10, 3, 2, 0.0), (6, 4, 4, 0.3), (4, 5, 1, 0.1)])
def test_bow_kernels(emu, frames, tmp_path, k, L, levelsup, zero):
    """k_bow_descend / k_bow_rank / k_bow_offsets == the oracle (pinned to the compiled DBoW2): word, weight and node per feature, FeatureVector"""
    from oracle import bow as OB
    _, _, fr = frames
    desc = np.ascontiguousarray(np.concatenate([fr[0].desc, fr[1].desc[:700]]))
    path = tmp_path / "voc.txt"
    OB.write_vocabulary(path, k, L, seed=7 * k + L, zero_weight_fraction=zero)
    voc = OB.Vocabulary(path)
    (kk, LL, _, _), parent, wid, ndesc, w = voc.export()
    off, cid, cdesc = _sibling_layout(parent, ndesc)
    n = len(desc)
    word = np.zeros(n, np.uint32); weight = np.zeros(n, np.float64); node = np.zeros(n, np.uint32)
    fvn = np.zeros(n, np.uint32); fvo = np.zeros(n + 1, np.int32); fvf = np.zeros(n, np.int32); nn = C.c_int()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu_bow_transform.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.POINTER(C.c_int)]
    kept = emu.emu_bow_transform(p(off), p(cid), p(cdesc), p(wid), p(w), LL, p(desc), n, levelsup, p(word), p(weight), p(node), p(fvn), p(fvo), p(fvf), C.byref(nn))
    o = voc.transform(desc, levelsup)
    assert kept == (o["weight"] > 0).sum() and nn.value == len(o["fv_nodes"]) >= 1 and (levelsup >= L or nn.value > 3)
    assert np.array_equal(word, o["word"]) and np.array_equal(weight.view(np.uint64), o["weight"].view(np.uint64)) and np.array_equal(node, o["node"])
    m = nn.value
    assert np.array_equal(fvn[:m], o["fv_nodes"]) and np.array_equal(fvo[:m + 1], o["fv_offsets"]) and np.array_equal(fvf[:kept], o["fv_features"])


def test_bow_entry_points_whole_unit(tmp_path, frames):
    """plvs_b200/csrc/bow.cu as a whole -- text loader, tree flattening, kernel sequence, read-back, BowVector -- built for the CPU (launches rewritten to
    emu::launch, CUDA runtime calls on host memory: tests/native_build.py) and driven through its own C ABI: plvs_voc_load_text + plvs_voc_transform ==
    the oracle, for a vocabulary file with and without a trailing newline; plvs_voc_create from flat arrays gives the same handle behaviour"""
    from oracle import bow as OB
    from tests.native_build import build_emulated_library
    lib = C.CDLL(build_emulated_library())
    _, _, fr = frames
    desc = np.ascontiguousarray(fr[0].desc)
    n = len(desc)
    lib.plvs_voc_load_text.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    lib.plvs_voc_destroy.argtypes = [C.c_void_p]; lib.plvs_voc_destroy.restype = None
    lib.plvs_voc_size.argtypes = [C.c_void_p]
    lib.plvs_voc_create.argtypes = [C.c_int] * 6 + [C.c_void_p] * 5
    lib.plvs_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_int)] + [C.c_void_p] * 3 + [C.POINTER(C.c_int), C.c_void_p]
    lib.plvs_last_error.restype = C.c_char_p
    p = lambda a: a.ctypes.data_as(C.c_void_p)

    def run(h, levelsup):
        word = np.zeros(n, np.uint32); weight = np.zeros(n, np.float64); node = np.zeros(n, np.uint32)
        bi = np.zeros(n, np.uint32); bv = np.zeros(n, np.float64); nb = C.c_int()
        fvn = np.zeros(n, np.uint32); off = np.zeros(n + 1, np.int32); feat = np.zeros(n, np.int32); nn = C.c_int()
        rc = lib.plvs_voc_transform(h, p(desc), n, 0, levelsup, p(word), p(weight), p(node), p(bi), p(bv), C.byref(nb), p(fvn), p(off), p(feat), C.byref(nn), None)
        assert rc == 0, lib.plvs_last_error()
        k = nn.value
        return dict(word=word, weight=weight, node=node, bow_ids=bi[:nb.value], bow_vals=bv[:nb.value], fv_nodes=fvn[:k], fv_offsets=off[:k + 1], fv_features=feat[:off[k]])

    def same(a, b):
        for name in b:
            x, y = a[name], b[name]
            assert np.array_equal(x.view(np.uint64), y.view(np.uint64)) if x.dtype == np.float64 else np.array_equal(x, y), name

    for k, L, levelsup, zero, scoring, weighting in ((10, 3, 2, 0.2, 0, 0), (7, 3, 1, 0.1, 5, 3)):
        path = tmp_path / ("voc%d.txt" % k)
        OB.write_vocabulary(path, k, L, seed=k, scoring=scoring, weighting=weighting, zero_weight_fraction=zero)
        ov = OB.Vocabulary(path)
        want = ov.transform(desc, levelsup)
        h = C.c_void_p()
        assert lib.plvs_voc_load_text(str(path).encode(), 0, C.byref(h)) == 0, lib.plvs_last_error()
        assert lib.plvs_voc_size(h) == ov.size() == k ** L
        same(run(h, levelsup), want)
        lib.plvs_voc_destroy(h)
        with open(path, "a") as f:
            f.write("\n")                                  # the file as saveToTextFile leaves it: the trailing empty line is ignored
        h = C.c_void_p()
        assert lib.plvs_voc_load_text(str(path).encode(), 0, C.byref(h)) == 0
        same(run(h, levelsup), want)
        lib.plvs_voc_destroy(h)
        (kk, LL, sc, wt), parent, wid, ndesc, w = ov.export()
        h = C.c_void_p()
        assert lib.plvs_voc_create(0, kk, LL, sc, wt, len(parent), p(parent), p(wid), p(ndesc), p(w), C.byref(h)) == 0, lib.plvs_last_error()
        same(run(h, levelsup), want)
        lib.plvs_voc_destroy(h)
    h = C.c_void_p()
    assert lib.plvs_voc_load_text(str(tmp_path / "missing.txt").encode(), 0, C.byref(h)) != 0
    bad = tmp_path / "bad.txt"; bad.write_text("hello world\n")
    assert lib.plvs_voc_load_text(str(bad).encode(), 0, C.byref(h)) != 0


# ---- the matcher translation unit as a whole (plvs_b200/csrc/match.cu) on the CPU model, through the product's own Python mirror ------------------
class _Missing:
    """stands for the entry points of the other translation units while the signatures are declared"""
    def __setattr__(self, k, v):
        pass


class _Partial:
    def __init__(self, lib):
        object.__setattr__(self, "_lib", lib)

    def __getattr__(self, name):
        try:
            return getattr(object.__getattribute__(self, "_lib"), name)
        except AttributeError:
            return _Missing()


# ---- the GPU tests themselves, replayed on the CPU model -------------------------------------------------------------------------------------
# Every test function of the GPU test modules is run again here, unchanged, with plvs_b200._lib.load() handing out the emulated translation units
# (tests/native_build.py: launches rewritten to the model's launcher, CUDA runtime calls on host memory).  Left out: the cases whose ORACLE side takes
# minutes (VGA / 1080p TSDF scans against the brute-force oracle, the 1080p extraction) and the NCCL merge test.
_REPLAY_MODULES = ("tests.test_gpu_orb", "tests.test_gpu_match", "tests.test_gpu_tsdf", "tests.test_gpu_frustum")
_REPLAY_SKIP = {"test_vga_1cm_single_scan": "oracle brute force over a VGA / 1 cm frustum", "test_1080p_5mm_scan_pair": "oracle brute force at 1080p / 5 mm",
                "test_orb_1080p": "1080p extraction on the CPU model", "test_orb_vga_1000_frames": "many VGA frames on the CPU model"}


# the heaviest replays (10-20 s each on the model) run with PLVS_EMU_FULL=1; the default set keeps the CPU suite within a few minutes
_REPLAY_FULL_ONLY = {"test_orb_batch_matches_single", "test_color_input_and_stereo_from_rgbd", "test_many_scans_carvable_mask_stays_exact", "test_cloud_sequence",
                     "test_device_distributor_equals_host_distributor", "test_orb_odd_sizes", "test_orb_lapping_area"}


def _replay_cases():
    import importlib
    import itertools
    import os
    full = bool(os.environ.get("PLVS_EMU_FULL"))
    cases = []
    for modname in _REPLAY_MODULES:
        mod = importlib.import_module(modname)
        for name in sorted(n for n in vars(mod) if n.startswith("test_")):
            fn = getattr(mod, name)
            if name in _REPLAY_SKIP or (name in _REPLAY_FULL_ONLY and not full):
                continue
            axes = []
            for mark in getattr(fn, "pytestmark", []):
                if mark.name == "parametrize":
                    names = [a.strip() for a in mark.args[0].split(",")] if isinstance(mark.args[0], str) else list(mark.args[0])
                    axes.append([dict(zip(names, v if len(names) > 1 else (v,))) for v in mark.args[1]])
            for combo in itertools.product(*axes) if axes else [()]:
                kw = {}
                for d in combo:
                    kw.update(d)
                cases.append(pytest.param(modname, name, kw, id="%s::%s%s" % (modname.split(".")[-1], name, "[%s]" % ",".join(str(v) for v in kw.values()) if kw else "")))
    return cases


@pytest.fixture
def product_bound_to_emulated_units(monkeypatch):
    """plvs_b200._lib.load() hands out the emulated translation units for the duration of one test: every product class then runs on the CPU model"""
    from tests.native_build import build_emulated_library
    monkeypatch.setattr(ABI, "_lib", ABI.declare(_Partial(C.CDLL(build_emulated_library()))))


_module_fixture_cache = {}


def _module_fixture(mod, name):
    """a module-scoped fixture of a GPU test module (their `frames`), built once per module through the emulated library"""
    key = (mod.__name__, name)
    if key not in _module_fixture_cache:
        fx = getattr(mod, name)
        raw = getattr(fx, "__wrapped__", None) or fx._get_wrapped_function()
        _module_fixture_cache[key] = raw(True)
    return _module_fixture_cache[key]


@pytest.mark.parametrize("modname,name,kw", _replay_cases())
def test_gpu_tests_replayed_on_the_cpu_model(product_bound_to_emulated_units, monkeypatch, tmp_path, modname, name, kw):
    import importlib
    import inspect
    mod = importlib.import_module(modname)
    fn = getattr(mod, name)
    args = {}
    for p in inspect.signature(fn).parameters:
        if p in kw:
            args[p] = kw[p]
        elif p == "gpu":
            args[p] = True
        elif p == "monkeypatch":
            args[p] = monkeypatch
        elif p == "tmp_path":
            args[p] = tmp_path
        else:
            args[p] = _module_fixture(mod, p)
    fn(**args)


@pytest.mark.parametrize("name", ["_impl_tsdf_from_raw_u16_depth", "_impl_mesh_read_out", "_impl_search_for_initialization", "_impl_search_local_points_resident",
                                  "_impl_bow_transform", "_impl_undistort_keypoints_on_device", "_impl_reference_goldens", "_impl_keyframe_ids", "_impl_deform",
                                  "_impl_world_cloud", "_impl_line_knn2", "_impl_map_file_round_trip"])
def test_unverified_gpu_tests_replayed_on_the_cpu_model(product_bound_to_emulated_units, tmp_path, name):
    """the bodies of tests/test_gpu_widened.py (the rows that have not met a GPU), unchanged, against the emulated translation units: entry points,
    host sequencing and kernels of the 16-bit depth path, the mesh read-out, SearchForInitialization, the resident SearchLocalPoints, the BoW
    transform and UndistortKeyPoints (on keypoints left "on the device" by the emulated extractor) all give the oracle's results"""
    import tests.test_gpu_widened as Z
    fn = getattr(Z, name)
    fn(tmp_path) if name in ("_impl_bow_transform", "_impl_reference_goldens", "_impl_map_file_round_trip") else fn()


def test_smoke_replayed_on_the_cpu_model(product_bound_to_emulated_units):
    """__graft_entry__.smoke() -- one small frame through extract -> match -> TSDF, each stage against the oracle (and the compiled reference when it is
    there) -- with every translation unit of the library on the CPU model"""
    import __graft_entry__ as g
    g.smoke()


def test_memcheck_on_the_cpu_model(tmp_path):
    """PLVS_EMU_GUARD=1: every "device" allocation of the emulated library sits between inaccessible pages and buffers are sized exactly, so an
    out-of-bounds access by a kernel or a copy faults (the model's stand-in for compute-sanitizer memcheck).  In a child interpreter: first that an
    overrun really faults, then smoke() and the six not-yet-verified GPU test bodies run clean."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, PLVS_EMU_GUARD="1")
    pre = ("import sys, ctypes as C, pathlib; sys.path.insert(0, %r); from plvs_b200 import _lib as ABI; from tests.native_build import build_emulated_library; "
           "from tests.test_emulated_kernels import _Partial; lib = C.CDLL(build_emulated_library()); ABI._lib = ABI.declare(_Partial(lib)); " % str(ROOT))
    bad = subprocess.run([sys.executable, "-c", pre + "p = C.c_void_p(); lib.plvs_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]; "
                          "assert lib.plvs_host_alloc(C.byref(p), 1000) == 0; b = (C.c_ubyte * 2000).from_address(p.value); b[999] = 1; print('in bounds', flush=True); b[1008] = 1; print('missed')"],
                         capture_output=True, text=True, env=env, cwd=str(ROOT))
    assert bad.returncode != 0 and "in bounds" in bad.stdout and "missed" not in bad.stdout
    body = ("import __graft_entry__ as g, tests.test_gpu_widened as Z; g.smoke(); "
            "[getattr(Z, n)() for n in ('_impl_tsdf_from_raw_u16_depth', '_impl_mesh_read_out', '_impl_search_for_initialization', '_impl_search_local_points_resident', "
            "'_impl_undistort_keypoints_on_device', '_impl_keyframe_ids')]; Z._impl_bow_transform(pathlib.Path(%r)); print('clean')" % str(tmp_path))
    ok = subprocess.run([sys.executable, "-c", pre + body], capture_output=True, text=True, env=env, cwd=str(ROOT), timeout=1200)
    assert ok.returncode == 0 and "clean" in ok.stdout, ok.stderr[-3000:]


def test_image_pyramids_of_the_extractor_on_the_cpu_model(product_bound_to_emulated_units):
    """mvImagePyramid / mvImagePyramidFiltered as ORBextractor::PrecomputeGaussianPyramid leaves them (src/ORBextractor.cc:1401-1427): every level of the
    device pyramid == cv2.resize chain, every filtered level == the 7x7 sigma-2 blur of that level's clone (the oracle's gauss7, pinned to cv2)"""
    from plvs_b200.orb import ORBextractor
    img = synth.gray_frame(2, 480, 360)
    ex = ORBextractor(800, 1.2, 8, 20, 7)
    ex(img)
    tab = O.Tables(800)
    pyr = O.pyramid_cv2(img, tab)
    for l in range(8):
        h, w = pyr[l].shape
        got = ex.pyramid_level(l, blurred=0)
        flt = ex.pyramid_level(l, blurred=1)
        assert got.shape == flt.shape
        oy, ox = (got.shape[0] - h) // 2, (got.shape[1] - w) // 2            # the device levels may carry their border
        assert np.array_equal(got[oy:oy + h, ox:ox + w], pyr[l]), l
        assert np.array_equal(flt[oy:oy + h, ox:ox + w], O.gauss7(pyr[l])), l



def test_randomised_scheduling_exposes_a_missing_barrier():
    """PLVS_EMU_SCHED_SEED: with round-robin scheduling a kernel whose threads read what thread 0 wrote WITHOUT a barrier looks fine (thread 0 always runs first);
    with a randomised order of thread segments some threads read too early.  The product's kernels give the oracle's results under such schedules too
    (run by hand over the whole file with seeds 1 and 2; profiles/r01_cpu_model_runs.md)."""
    import os
    import subprocess
    import sys
    code = ("import sys, ctypes as C; sys.path.insert(0, %r); from tests.native_build import build_emulated_kernels; "
            "print(C.CDLL(build_emulated_kernels()).emu_missing_barrier_demo())" % str(ROOT))
    plain = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT), env={k: v for k, v in os.environ.items() if k != "PLVS_EMU_SCHED_SEED"})
    assert plain.returncode == 0 and int(plain.stdout.split()[-1]) == 256
    seen = []
    for seed in ("1", "2", "3"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT), env=dict(os.environ, PLVS_EMU_SCHED_SEED=seed))
        assert r.returncode == 0
        seen.append(int(r.stdout.split()[-1]))
    assert all(0 < s < 256 for s in seen), seen


def test_bench_step_dataflow_on_the_cpu_model(product_bound_to_emulated_units, monkeypatch):
    """bench.py's unit of work -- plvs_b200.pipeline.HotPath.step and the threaded HotPath.run_stream: batch extraction, the two tracking searches
    on the device-resident frame, the triangulation search, the colour depth-scan integration -- on the CPU model over a short synthetic stream,
    compared with the oracle stage by stage (the body of tests/test_gpu_bench_config.py::test_hot_path_step_at_bench_config, reduced): the bench
    measures the path the parity tests pin, with nothing skipped"""
    import tests.test_gpu_bench_config as B
    monkeypatch.setenv("PLVS_PIPELINE_SERIAL", "1")      # the model runs one host thread at a time: the driver's stages in sequence
    nblk, nm = B._impl_hot_path(320, 240, 500, 0.04, 4.0, 2, 2, 4096, threaded=False, native="host")
    assert nblk > 20 and nm > 100


def test_dataset_stream_through_the_hot_path_on_the_cpu_model(product_bound_to_emulated_units, monkeypatch, tmp_path):
    """`bench.py --dataset DIR` (BASELINE configs[0]): a sequence in the TUM RGB-D benchmark's on-disk format goes through StreamData.from_tum, HotPath.prepare and
    the native stream driver exactly like the synthetic stream; same totals as the step-by-step form, map equal to the oracle fed the loaded frames"""
    pytest.importorskip("cv2")
    from plvs_b200 import tsdf as T
    from plvs_b200.pipeline import StreamData, HotPath
    from oracle import tsdf as OT
    from tests.test_tum_loader import _write_sequence
    monkeypatch.setenv("PLVS_PIPELINE_SERIAL", "1")
    _write_sequence(tmp_path, 5, 320, 240, True)
    d = StreamData.from_tum(tmp_path, 5, pinned=False)
    hp = HotPath(d, nfeatures=500, voxel=0.04, far=4.0, max_blocks=4096, batch=2)
    hp.prepare()
    hp.tsdf.Reset(); hp.tsdf.integrate(d.depth[0], d.poses[0], d.bgr[0])
    got = {}
    for s_ in range(2):
        for k, v in hp.step(1 + 2 * s_, 2, resident=False, concurrent=False).items():
            got[k] = got.get(k, 0) + v
    assert got["keypoints"] > 400 and got["matches"] > 20
    p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=1)
    o = OT.Map(p, threads=4); o.set_camera(d.K["fx"], d.K["fy"], d.K["cx"], d.K["cy"], d.w, d.h)
    for f in range(5):
        o.integrate(d.depth[f], d.poses[f], d.bgr[f])
    gk, gs, gw, gc = hp.tsdf.download(); ok, os_, ow, oc = o.download()
    assert np.array_equal(gk, ok) and len(gk) > 20 and np.allclose(gs, os_, atol=1e-4) and np.array_equal(gw.view(np.uint32), ow.view(np.uint32)) and np.array_equal(gc, oc)
    hp.tsdf.Reset(); hp.tsdf.integrate(d.depth[0], d.poses[0], d.bgr[0])
    agg = hp.run_stream_native(1, 2, False)
    assert agg["keypoints"] == got["keypoints"] and agg["matches"] == got["matches"]


def test_bench_geometry_scan_sequence_on_the_cpu_model(product_bound_to_emulated_units):
    """the body of tests/test_gpu_bench_config.py::test_c2_bench_geometry_ten_scans at 160x120 / 4 cm: consecutive colour scans with carving on one
    map, planes 0.1-5 m, compared after every scan"""
    import tests.test_gpu_bench_config as B
    n, s, exact = B._impl_depth_scan_sequence(160, 120, 0.04, 5.0, 5, 8192)
    assert n > 50 and exact


def test_error_paths_of_the_new_entry_points(product_bound_to_emulated_units, tmp_path):
    """argument validation and capacity errors of the entry points written last return codes (PLVS_EINVAL / PLVS_ECAP / PLVS_ESTATE) with a message --
    nothing aborts, nothing is written out of bounds (the reference quick_exit()s in places like these)"""
    from plvs_b200.matcher import ORBmatcher
    from plvs_b200.bow import ORBVocabulary
    from oracle import bow as OB
    lib = ABI.load()
    K = synth.intrinsics(160, 120)
    g = T.ChiselServer(T.default_params(voxel_resolution=0.04, max_blocks=1024, use_color=1))
    d16 = np.full((120, 160), 5000, np.uint16); pose = np.ascontiguousarray(synth.pose(0), np.float32).reshape(12)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.plvs_tsdf_integrate_depth_u16(g._h, p(d16), 160, 120, 320, 0.0002, None, 0, 0, p(pose), T.SCAN) != 0            # no camera yet
    g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], 160, 120)
    assert lib.plvs_tsdf_integrate_depth_u16(g._h, p(d16), 160, 120, 100, 0.0002, None, 0, 0, p(pose), T.SCAN) != 0            # row stride shorter than a row
    assert lib.plvs_tsdf_integrate_depth_u16(g._h, p(d16), 160, 120, 320, 0.0002, None, 0, 3, p(pose), T.SCAN_COLOR) != 0      # colour mode without an image
    assert b"colour" in lib.plvs_last_error() or b"color" in lib.plvs_last_error()
    g.integrate_u16(d16, 0.0002, synth.pose(0), synth.bgr_frame(0, 160, 120))
    nm, nv = g.UpdateMesh()
    assert nv > 0
    small = np.zeros(3, np.float32)
    assert lib.plvs_tsdf_get_meshes(g._h, None, None, 0, p(small), None, None, 1, 0) == -4                                    # PLVS_ECAP
    k1 = np.zeros(1, np.uint32)
    assert lib.plvs_tsdf_get_mesh_kfids(g._h, p(k1), 1, 0) == -4
    n = C.c_int()
    assert lib.plvs_tsdf_download_kfid(g._h, p(k1), 0, C.byref(n)) == -4 and n.value > 0
    assert (g.download_kfid() == 0).all()                                                                                   # no cloud with ids yet
    m = ORBmatcher(0.9, True)
    kp = np.zeros(0, KP_DTYPE); de = np.zeros((0, 32), np.uint8)
    e = Frame(kp, de, 640, 480, O.Tables(1000).scale)
    n0, a0, p0 = m.SearchForInitialization(e, e, np.zeros((0, 2), np.float32), 100)
    assert n0 == 0 and len(a0) == 0
    v = e.view(); nmatch = C.c_int(); out = np.zeros(1, np.int32)
    assert lib.plvs_match_initialization(m._h, C.byref(v), C.byref(v), None, -5, 0.9, 1, p(out), C.byref(nmatch)) != 0        # negative window
    voc = ORBVocabulary()
    assert not voc.loadFromTextFile(tmp_path / "nope.txt")
    path = tmp_path / "v.txt"; OB.write_vocabulary(path, 4, 2, seed=1)
    assert voc.loadFromTextFile(path)
    r = voc.transform(np.zeros((0, 32), np.uint8), 1)
    assert len(r["word"]) == 0 and len(r["bow_ids"]) == 0 and len(r["fv_nodes"]) == 0
    h = C.c_void_p()
    par = np.array([0, 5, 0], np.int32); wid = np.array([-1, 0, 1], np.int32); dsc = np.zeros((3, 32), np.uint8); wgt = np.ones(3)
    assert lib.plvs_voc_create(0, 2, 1, 0, 0, 3, p(par), p(wid), p(dsc), p(wgt), C.byref(h)) != 0                             # a parent that does not precede its child
