"""CPU: the header-only C++ shims (shim/plvs_shim.hpp) compile against stand-in OpenCV/Eigen types and link
against libplvs_b200.so -- every template that gathers from Frame/MapPoint/KeyFrame is instantiated."""
import ctypes as C
import pathlib
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_shim_compiles_and_links():
    from plvs_b200 import _lib
    _lib.load()
    out = ROOT / "tests" / "native" / "libshim_check.so"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-fPIC", "-shared", str(ROOT / "tests/native/shim_compile.cpp"), "-o", str(out),
           f"-L{ROOT / 'plvs_b200'}", "-lplvs_b200", f"-Wl,-rpath,{ROOT / 'plvs_b200'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(str(out))
    assert lib.shim_instantiate(0) > 0
