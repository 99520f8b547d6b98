import fcntl, os, pathlib, subprocess

HERE = pathlib.Path(__file__).resolve().parent


def _build_once(out, deps, cmd):
    """run `cmd` (which must write out + '.tmp') unless `out` is newer than `deps`; pytest-xdist workers serialise on a lock file and the result appears atomically"""
    with open(str(out) + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not out.exists() or any(out.stat().st_mtime < d.stat().st_mtime for d in deps):
            subprocess.check_call(cmd)
            os.replace(str(out) + ".tmp", str(out))
    return str(out)


def build_host_checks():
    src = HERE / "native" / "host_checks.cpp"
    out = HERE / "native" / "libhost_checks.so"
    deps = [src, HERE.parent / "plvs_b200" / "csrc" / "orb_distribute.hpp", HERE.parent / "plvs_b200" / "csrc" / "libm_sincosf.cuh",
            HERE.parent / "plvs_b200" / "csrc" / "stdsort_emul.cuh"]
    return _build_once(out, deps, ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", str(src), "-o", str(out) + ".tmp", "-lm"])


def build_emulated_kernels():
    """tests/native/emu_kernels.cpp: the product's device-only kernel headers compiled for the CPU execution model of cuda_emu.hpp"""
    src = HERE / "native" / "emu_kernels.cpp"
    out = HERE / "native" / "libemu_kernels.so"
    csrc = HERE.parent / "plvs_b200" / "csrc"
    deps = [src, HERE / "native" / "cuda_emu.hpp", HERE.parent / "include" / "plvs_b200.h"] + \
        [csrc / n for n in ("match_common.cuh", "match_frustum.cuh", "match_init.cuh", "orb_undistort.cuh", "tsdf_hash.cuh", "mesh_kernels.cuh", "mc_tables.inc", "bow_kernels.cuh")]
    return _build_once(out, deps, ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", str(src), "-o", str(out) + ".tmp", "-lm"])


CLUSTER_KERNELS = {"k_resolve": 8}          # __cluster_dims__ of the kernels that need their CTAs resident together


def build_emulated_library(units=("core.cu", "bow.cu", "match.cu", "tsdf.cu", "orb.cu", "pipeline.cu")):
    """Whole translation units of the product on the CPU: `kernel<<<grid, block, smem, stream>>>(args);` is rewritten to emu::launch, the CUDA runtime
    calls resolve to tests/native/fake_cuda/cuda_runtime.h (host memory), device code runs on cuda_emu.hpp.  Only units without inline PTX qualify.
    -> tests/native/libemu_units.so exporting the same C ABI as the real library for those units."""
    import re
    csrc = HERE.parent / "plvs_b200" / "csrc"
    gen = HERE / "native" / "_gen"
    out = HERE / "native" / "libemu_units.so"
    deps = [csrc / u for u in units] + list(csrc.glob("*.cuh")) + [ HERE / "native" / "cuda_emu.hpp", HERE / "native" / "fake_cuda" / "cuda_runtime.h",
                                        HERE.parent / "include" / "plvs_b200.h", pathlib.Path(__file__)]
    if out.exists() and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(out)
    gen.mkdir(parents=True, exist_ok=True)
    import fcntl
    lock = open(gen / ".build.lock", "w")
    fcntl.flock(lock, fcntl.LOCK_EX)                      # pytest-xdist workers: one of them builds, the others find the result
    if out.exists() and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(out)
    launch = re.compile(r"(\b[A-Za-z_][\w:]*(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
    srcs = []
    for u in units:
        text = (csrc / u).read_text()
        res, pos = [], 0
        for m in launch.finditer(text):
            cfg, depth0, cur = [], 0, ""
            for ch in m.group(2):             # split the launch configuration at its top-level commas
                if ch == "," and depth0 == 0:
                    cfg.append(cur.strip()); cur = ""
                else:
                    depth0 += {"(": 1, ")": -1}.get(ch, 0); cur += ch
            cfg.append(cur.strip())
            assert len(cfg) >= 2, (u, m.group(0))
            depth, j = 1, m.end()
            while depth:                      # the matching parenthesis of the argument list
                depth += {"(": 1, ")": -1}.get(text[j], 0); j += 1
            args = text[m.end():j - 1]
            assert text[j:].lstrip().startswith(";"), (u, text[m.start():j + 20])
            grid, block, smem = cfg[0], cfg[1], cfg[2] if len(cfg) > 2 else "0"
            cl = CLUSTER_KERNELS.get(m.group(1).split("<")[0].strip(), 1)
            res.append(text[pos:m.start()] + "emu::launch_cluster(%d, dim3(%s), dim3(%s), (size_t)(%s), [&] { %s(%s); })" % (cl, grid, block, smem, m.group(1), args))
            pos = j
        res.append(text[pos:])
        body = "".join(res).replace('#include "', '#include "%s/' % csrc)
        dst = gen / (u.replace(".cu", "_emu.cpp"))
        dst.write_text('#include "%s"\n' % (HERE / "native" / "cuda_emu.hpp") + body)
        srcs.append(str(dst))
    tmp = out.with_suffix(".so.tmp")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", "-I", str(HERE / "native" / "fake_cuda")] + srcs +
                          ["-o", str(tmp), "-lm", "-pthread"])
    tmp.replace(out)
    return str(out)
