"""Build libplvs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os, pathlib, subprocess, sys

HERE = pathlib.Path(__file__).resolve().parent
OUT = HERE.parent / "libplvs_b200.so"
SRCS = ["core.cu", "orb.cu", "match.cu", "tsdf.cu", "bow.cu", "pipeline.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--fmad=false",                      # bit-exact fp32 paths: contraction is opted into per kernel, never implicit
         "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off,-pthread", "-shared", "-Xptxas", "-v"]


def build(force=False, verbose=False):
    srcs = [HERE / s for s in SRCS if (HERE / s).exists()]
    deps = list(HERE.glob("*.cu")) + list(HERE.glob("*.cuh")) + list(HERE.glob("*.hpp")) + list(HERE.glob("*.inc")) + \
        [HERE.parent.parent / "include" / "plvs_b200.h"]
    if OUT.exists() and not force and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(OUT)
    cmd = [NVCC] + FLAGS + ["-o", str(OUT)] + [str(s) for s in srcs] + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed")
    (HERE.parent / "build_ptxas.log").write_text(r.stdout + r.stderr)
    return str(OUT)


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
