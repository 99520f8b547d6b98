"""The C++ drop-in shims (shim/plvs_shim.hpp) EXECUTED: PLVS2::ORBextractor, PLVS2::ORBmatcher (the three hot-path overloads) and
chisel_server::ChiselServer (both fusion routes + UpdateMesh/GetPointCloud) are driven by tests/native/shim_runtime.cpp through stand-in
Frame / MapPoint / KeyFrame objects and what they leave in those objects is compared with the oracle and with the Python mirror
(reference surfaces: include/ORBextractor.h:86, include/ORBmatcher.h:68-97, Thirdparty/chisel_server/include/chisel_server/ChiselServer.h:190-286).
The gather order (`src`), the claim flags and the scatter `F.mvpMapPoints[i] = vpMapPoints[src[assign[i]]]` are what is at stake.

-m gpu: the harness links against plvs_b200/libplvs_b200.so and runs on the device.  CPU suite: the same harness linked against the library's
translation units compiled for the CPU execution model of tests/native/cuda_emu.hpp (tests/native_build.py)."""
import ctypes as C
import pathlib
import subprocess
import numpy as np
import pytest

from plvs_b200 import synth, scenario, _lib as ABI
from plvs_b200.matcher import MP_QUERY, LAST_QUERY, featvec, featvec_struct
from plvs_b200.orb import KP_DTYPE
from oracle import orb as O, match as OM, tsdf as OT

ROOT = pathlib.Path(__file__).resolve().parent.parent


class FrameIn(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys", C.c_void_p), ("desc", C.c_void_p), ("uright", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float), ("grid_inv_w", C.c_float), ("grid_inv_h", C.c_float),
                ("nlevels", C.c_int32), ("scale", C.c_void_p), ("sigma2", C.c_void_p), ("bf", C.c_float), ("b", C.c_float)]


def _frame_in(fr, b=0.08):
    keep = [fr.keys, fr.desc, fr.uright, np.ascontiguousarray(fr.scale_factors, np.float32), np.ascontiguousarray(fr.level_sigma2, np.float32)]
    s = FrameIn(fr.n, keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data if keep[2] is not None else None,
                fr.min_x, fr.min_y, fr.max_x, fr.max_y, float(fr.grid_inv_w), float(fr.grid_inv_h), len(keep[3]), keep[3].ctypes.data, keep[4].ctypes.data,
                fr.bf, b)
    s._keep = keep
    return s


def build_harness(link_lib):
    """g++ shim_runtime.cpp against `link_lib` (the product library, or its CPU-model build); one .so per library"""
    link_lib = pathlib.Path(link_lib)
    out = ROOT / "tests" / "native" / ("libshim_runtime_%s.so" % link_lib.stem.replace("lib", "", 1))
    src = ROOT / "tests" / "native" / "shim_runtime.cpp"
    deps = [src, ROOT / "shim" / "plvs_shim.hpp", ROOT / "shim" / "standin.hpp", ROOT / "include" / "plvs_b200.h", link_lib]
    if not out.exists() or any(out.stat().st_mtime < d.stat().st_mtime for d in deps):
        cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-fPIC", "-shared", str(src), "-o", str(out),
               f"-L{link_lib.parent}", f"-l:{link_lib.name}", f"-Wl,-rpath,{link_lib.parent}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(out))
    lib.shim_rt_error.restype = C.c_char_p
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _frames(nfeatures=1000):
    w, h = 640, 480
    K = synth.intrinsics(w, h)
    tab = O.Tables(nfeatures)
    fr = []
    for f in (4, 5):
        kp, desc, _, _ = O.extract_port(synth.gray_frame(f, w, h), nfeatures)
        x = scenario.make_frame(kp, desc, synth.depth_frame(f, w, h), K, tab.scale)
        x.level_sigma2 = np.asarray(tab.sigma2, np.float32)
        fr.append(x)
    return K, fr


def run_extractor(H):
    img = synth.gray_frame(3, 640, 480)
    cap = 2 * 1000 + 64 * 8
    kp = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); mono = C.c_int(); l1 = np.zeros(640 * 480, np.uint8); wh = np.zeros(2, np.int32)
    n = H.shim_rt_extract(_p(img), 640, 480, 1000, 0, 0, _p(kp), _p(desc), cap, C.byref(mono), _p(l1), l1.size, _p(wh))
    assert n > 0, H.shim_rt_error()
    okp, odesc, omono, _ = O.extract_port(img, 1000)
    assert n == len(okp) and mono.value == omono and np.array_equal(kp[:n], okp) and np.array_equal(desc[:n], odesc)
    tab = O.Tables(1000)
    lw, lh = tab.level_size(640, 480, 1)
    assert (int(wh[0]), int(wh[1])) == (lw, lh)
    assert np.array_equal(l1[:lw * lh].reshape(lh, lw), O.resize_linear(img, lw, lh))          # mvImagePyramid[1] as Frame::ComputeStereoMatches reads it
    # lapping area: monoIndex / ordering as the reference's stereo-fisheye path (include/ORBextractor.h:86 vLappingArea)
    n2 = H.shim_rt_extract(_p(img), 640, 480, 1000, 0, 400, _p(kp), _p(desc), cap, C.byref(mono), _p(l1), l1.size, _p(wh))
    okp, odesc, omono, _ = O.extract_port(img, 1000, lapping=(0, 400))
    assert n2 == len(okp) and mono.value == omono and np.array_equal(kp[:n2], okp) and np.array_equal(desc[:n2], odesc)


def run_search_map(H):
    K, (last, cur) = _frames()
    q, _ = scenario.map_queries(last, cur, K, synth.pose(4), synth.pose(5), seed=3)
    rng = np.random.default_rng(5)
    nq = len(q)
    # vpMapPoints as Tracking holds it: in-view points interleaved with points not in view, bad points and points without observations
    state = np.full(nq, 1 | 4, np.uint8)
    state[rng.random(nq) < 0.15] &= ~np.uint8(1)          # mbTrackInView = false
    state[rng.random(nq) < 0.05] |= 2                     # isBad()
    state[rng.random(nq) < 0.10] &= ~np.uint8(4)          # Observations() == 0
    pre = np.zeros(cur.n, np.uint8)
    r = rng.random(cur.n); pre[r < 0.10] = 1; pre[(r >= 0.10) & (r < 0.15)] = 2
    fin = _frame_in(cur)
    for th, far in ((3.0, 0), (5.0, 1)):
        out = np.zeros(cur.n, np.int32)
        n = H.shim_rt_search_map(C.byref(fin), _p(q), _p(state), nq, _p(pre), C.c_float(th), C.c_float(0.8), far, C.c_float(4.0), _p(out))
        assert n >= 0, H.shim_rt_error()
        # what the reference does with the same objects: skip !mbTrackInView / isBad (src/ORBmatcher.cc:80-95), claims = Observations() > 0
        src = np.nonzero(((state & 1) != 0) & ((state & 2) == 0))[0]
        qq = q[src].copy(); qq["flags"] = ((state[src] & 4) != 0).astype(np.uint32)
        on, oassign = OM.search_by_projection_map(cur, qq, th, 0.8, bool(far), 4.0, claimed=(pre == 1).astype(np.uint8))
        want = np.where(oassign >= 0, src[np.maximum(oassign, 0)], np.where(pre > 0, -2, -1))
        assert n == on and np.array_equal(out, want), (th, far, n, on, int((out != want).sum()))
        assert on > 100


def run_search_last(H):
    K, (last, cur) = _frames()
    cam4 = np.array([K["fx"], K["fy"], K["cx"], K["cy"]], np.float32)
    rng = np.random.default_rng(11)
    # camera-frame positions of the last frame's points in the CURRENT camera (any consistent set will do: the shim projects them itself)
    z = np.where(last.depth_at_kp > 0, last.depth_at_kp, np.float32(1.0)).astype(np.float32)
    shift = np.float32(6.0)
    pc = np.stack([(last.keys["x"] + shift - cam4[2]) * z / cam4[0], (last.keys["y"] - cam4[3]) * z / cam4[1], z], 1).astype(np.float32)
    mp = np.where(last.depth_at_kp > 0, 1, 0).astype(np.uint8)
    r = rng.random(last.n); mp[(mp == 1) & (r < 0.08)] = 2; mp[(mp == 1) & (r > 0.95)] = 3
    mpdesc = last.desc.copy()
    for cur_t, last_t, mono in (((0, 0, 0), (0, 0, 0), 0), ((0, 0, 0), (0, 0, 0.3), 0), ((0, 0, 0), (0, 0, -0.3), 0), ((0, 0, 0.5), (0, 0, 0.9), 1)):
        cur_t, last_t = np.array(cur_t, np.float32), np.array(last_t, np.float32)
        fcur, flast = _frame_in(cur), _frame_in(last)
        out = np.zeros(cur.n, np.int32)
        n = H.shim_rt_search_last(C.byref(fcur), C.byref(flast), _p(mp), _p(pc), _p(mpdesc), _p(cam4), _p(cur_t), _p(last_t), C.c_float(15.0), mono, 1, _p(out))
        assert n >= 0, H.shim_rt_error()
        # the same expressions in numpy float32: world = pc - cur_t, x3Dc = world + cur_t, uv = fx*x/z + cx, invz = (float)(1.0 / z)
        src = np.nonzero((mp == 1) | (mp == 2))[0]
        xc = ((pc[src] - cur_t) + cur_t).astype(np.float32)
        q = np.zeros(len(src), LAST_QUERY)
        q["u"] = cam4[0] * xc[:, 0] / xc[:, 2] + cam4[2]; q["v"] = cam4[1] * xc[:, 1] / xc[:, 2] + cam4[3]
        q["invz"] = (1.0 / xc[:, 2].astype(np.float64)).astype(np.float32)
        q["last_octave"] = last.keys["octave"][src]; q["angle"] = last.keys["angle"][src]; q["flags"] = (mp[src] == 1); q["desc"] = mpdesc[src]
        tlc_z = np.float32(np.float32(-cur_t[2]) + last_t[2])          # (Tlw * twc)(2) with translation-only poses
        fwd, bwd = bool(tlc_z > np.float32(0.08)) and not mono, bool(-tlc_z > np.float32(0.08)) and not mono
        on, oassign = OM.search_by_projection_last(cur, q, 15.0, fwd, bwd, True)
        want = np.where(oassign >= 0, src[np.maximum(oassign, 0)], -1)
        assert n == on and np.array_equal(out, want), (cur_t, last_t, mono, n, on)
        assert on > 100


def run_triangulation(H):
    K, (last, cur) = _frames()
    fv1, fv2 = featvec(scenario.node_ids(cur.desc, 256)), featvec(scenario.node_ids(last.desc, 256))
    F12, ep = scenario.fundamental(K, synth.pose(5), synth.pose(4))
    rng = np.random.default_rng(2)
    has1, has2 = (rng.random(cur.n) < 0.3).astype(np.uint8), (rng.random(last.n) < 0.3).astype(np.uint8)
    s1, s2 = featvec_struct(fv1), featvec_struct(fv2)
    f1, f2 = _frame_in(cur), _frame_in(last)
    for only_stereo, coarse, ori in ((0, 0, 0), (0, 1, 1), (1, 0, 1)):
        pairs = np.zeros((cur.n, 2), np.int32)
        n = H.shim_rt_triangulation(C.byref(f1), C.byref(f2), C.byref(s1), C.byref(s2), _p(has1), _p(has2), _p(F12), _p(ep), only_stereo, coarse,
                                    C.c_float(0.6), ori, _p(pairs), cur.n)
        assert n >= 0, H.shim_rt_error()
        on, om12 = OM.search_for_triangulation(cur, last, fv1, fv2, has1, has2, F12, ep, bool(only_stereo), bool(coarse), bool(ori))
        idx = np.nonzero(om12 >= 0)[0]
        assert n == on and np.array_equal(pairs[:n], np.stack([idx, om12[idx]], 1))
    assert on > 20


def run_chisel(H):
    from plvs_b200 import tsdf as T
    w, h, ns = 160, 120, 3
    K = synth.intrinsics(w, h)
    cam4 = np.array([K["fx"], K["fy"], K["cx"], K["cy"]], np.float64)
    depth = np.stack([synth.depth_frame(f, w, h) for f in range(ns)]); bgr = np.stack([synth.bgr_frame(f, w, h) for f in range(ns)])
    poses = np.stack([np.ascontiguousarray(synth.pose(f), np.float32).reshape(12) for f in range(ns)])
    for route, color in ((0, 1), (0, 0), (1, 1)):
        cap = 8192
        keys = np.zeros((cap, 3), np.int32); sdf = np.zeros((cap, 4096), np.float32); wt = np.zeros((cap, 4096), np.float32); rgba = np.zeros((cap, 4096, 4), np.uint8)
        capc = 1 << 20
        cx, cn, cc = np.zeros((capc, 3), np.float32), np.zeros((capc, 3), np.float32), np.zeros((capc, 4), np.uint8)
        nb, nc = C.c_int(), C.c_int()
        rc = H.shim_rt_chisel(C.c_float(0.04), C.c_float(0.1), C.c_float(4.0), 1, C.c_float(0.05), color, cap, _p(cam4), w, h, ns, _p(depth), _p(bgr), _p(poses),
                              route, 2, _p(keys), _p(sdf), _p(wt), _p(rgba), cap, C.byref(nb), _p(cx), _p(cn), _p(cc), capc, C.byref(nc))
        assert rc == 0, H.shim_rt_error()
        p = T.default_params(voxel_resolution=0.04, use_carving=1, carving_dist=0.05, near_plane=0.1, far_plane=4.0, max_blocks=cap, use_color=color)
        o = OT.Map(p, threads=8); o.set_camera(*cam4, w, h)
        for f in range(ns):
            if route == 0:
                o.integrate(depth[f], synth.pose(f), bgr[f] if color else None)
            else:
                # the cloud the harness builds: every 2nd pixel with z > 0, x = (u - cx) * z / fx in float; colours r,g,b * (1/255)
                v, u = np.mgrid[0:h:2, 0:w:2]
                z = depth[f][v, u]; ok = z > 0
                uu, vv, z = u[ok].astype(np.float32), v[ok].astype(np.float32), z[ok]
                xyz = np.stack([(uu - np.float32(cam4[0 + 2])) * z / np.float32(cam4[0]), (vv - np.float32(cam4[3])) * z / np.float32(cam4[1]), z], 1).astype(np.float32)
                c = bgr[f][v[ok], u[ok]].astype(np.float32) * (np.float32(1.0) / np.float32(255.0))
                o.integrate_cloud(xyz, np.ascontiguousarray(c[:, ::-1]), synth.pose(f), depth[f])
        ok_, os_, ow, oc = o.download()
        n = nb.value
        assert n == len(ok_)
        order = np.lexsort((keys[:n, 2], keys[:n, 1], keys[:n, 0]))          # plvs_tsdf_download_blocks returns pool order
        assert np.array_equal(keys[:n][order], ok_)
        assert np.array_equal(wt[:n][order].view(np.uint32), ow.view(np.uint32)) and np.array_equal(sdf[:n][order].view(np.uint32), os_.view(np.uint32))
        if color:
            assert np.array_equal(rgba[:n][order], oc)
        # GetPointCloud: one point per mesh vertex of the oracle's marching cubes, chunks in key order
        mk, mcnt, mv, mn, mc = o.extract_mesh()
        assert nc.value == len(mv) > 100
        assert np.array_equal(cx[:len(mv)].view(np.uint32), mv.view(np.uint32))
        if color:
            assert np.array_equal(cn[:len(mv)].view(np.uint32), mn.view(np.uint32))
            want = (mc * np.float32(255)).astype(np.uint8)                # point.r = color * 255 (ChiselServer.cpp:956-958): float -> uint8 truncation
            assert np.array_equal(cc[:len(mv), :3], want)


ALL = (run_extractor, run_search_map, run_search_last, run_triangulation, run_chisel)


@pytest.mark.gpu
@pytest.mark.parametrize("body", ALL, ids=lambda f: f.__name__)
def test_shim_runtime_on_the_gpu(gpu, body):
    body(build_harness(ROOT / "plvs_b200" / "libplvs_b200.so"))


@pytest.mark.parametrize("body", ALL, ids=lambda f: f.__name__)
def test_shim_runtime_on_the_cpu_model(body):
    from tests.native_build import build_emulated_library
    body(build_harness(build_emulated_library()))
