import pathlib, subprocess

HERE = pathlib.Path(__file__).resolve().parent


def build_host_checks():
    src = HERE / "native" / "host_checks.cpp"
    out = HERE / "native" / "libhost_checks.so"
    deps = [src, HERE.parent / "plvs_b200" / "csrc" / "orb_distribute.hpp", HERE.parent / "plvs_b200" / "csrc" / "libm_sincosf.cuh",
            HERE.parent / "plvs_b200" / "csrc" / "stdsort_emul.cuh"]
    if not out.exists() or any(out.stat().st_mtime < d.stat().st_mtime for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", str(src), "-o", str(out), "-lm"])
    return str(out)


def build_emulated_kernels():
    """tests/native/emu_kernels.cpp: the product's device-only kernel headers compiled for the CPU execution model of cuda_emu.hpp"""
    src = HERE / "native" / "emu_kernels.cpp"
    out = HERE / "native" / "libemu_kernels.so"
    csrc = HERE.parent / "plvs_b200" / "csrc"
    deps = [src, HERE / "native" / "cuda_emu.hpp", HERE.parent / "include" / "plvs_b200.h"] + \
        [csrc / n for n in ("match_common.cuh", "match_frustum.cuh", "match_init.cuh", "orb_undistort.cuh", "tsdf_hash.cuh", "mesh_kernels.cuh", "mc_tables.inc", "bow_kernels.cuh")]
    if not out.exists() or any(out.stat().st_mtime < d.stat().st_mtime for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", str(src), "-o", str(out), "-lm"])
    return str(out)
