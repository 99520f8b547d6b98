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
