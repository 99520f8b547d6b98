"""Build the CPU oracle (test infrastructure) -> oracle/liboracle.so.

Flags mirror the reference's Ubuntu-24.04 configuration (config.sh:19-21: no
-march=native => no FMA contraction); -ffp-contract=off makes that explicit.
The reference's own sources for the path are compiled separately by oracle/ref_build.py
(against stand-in headers for the libraries this image lacks) into oracle/_ref/.
"""
import os, subprocess, sys, pathlib

HERE = pathlib.Path(__file__).resolve().parent
SRCS = ["orb_oracle.cpp", "match_oracle.cpp", "tsdf_oracle.cpp", "bow_oracle.cpp"]
OUT = HERE / "liboracle.so"


def build(force=False):
    srcs = [HERE / s for s in SRCS if (HERE / s).exists()]
    deps = srcs + [HERE.parent / "plvs_b200" / "csrc" / "orb_pattern.inc", HERE.parent / "plvs_b200" / "csrc" / "mc_tables.inc"]
    if OUT.exists() and not force and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return str(OUT)
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
           "-o", str(OUT)] + [str(s) for s in srcs] + ["-lm"]
    subprocess.check_call(cmd)
    return str(OUT)


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
