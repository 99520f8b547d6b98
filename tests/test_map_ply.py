"""CPU: the volumetric map on disk ("PLY save / load").  plvs_map_save_ply against the reference's own writer --
PointCloudMap<PointT>::WritePLY, sliced out of src/PointCloudMap.cc and compiled into oracle/_ref/libmapply_ref.so -- byte for byte (binary and ASCII, with
and without faces), and plvs_map_load_ply on the files the reference writes.  Host I/O only: no device involved."""
import ctypes as C
import filecmp
import numpy as np
import pytest

from plvs_b200 import tsdf as T


def _ref():
    try:
        from oracle import ref_build
        path = ref_build.build_mapply()
        if path is None:
            return None
        lib = C.CDLL(path)
        lib.ref_write_map_ply.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int]
        return lib
    except Exception:
        return None


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    xyz = (rng.standard_normal((n, 3)) * 3).astype(np.float32)
    xyz[:: 7] = np.round(xyz[:: 7])                     # integers, zeros and tiny / huge magnitudes: every branch of the ostream float formatting
    if n > 3:
        xyz[1] = (1e-7, -2.5e9, 0.0)
    nrm = rng.standard_normal((n, 3)).astype(np.float32)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-9)
    bgra = rng.integers(0, 256, (n, 4), dtype=np.uint8)
    label = rng.integers(0, 1 << 31, n, dtype=np.uint32); kfid = rng.integers(0, 5000, n, dtype=np.uint32)
    return xyz, bgra, nrm.astype(np.float32), label, kfid


@pytest.mark.skipif(_ref() is None, reason="oracle/_ref/libmapply_ref.so not built (no /root/reference here)")
@pytest.mark.parametrize("n", [0, 3, 300, 3000, 1, 3001])
@pytest.mark.parametrize("is_mesh,binary", [(1, 1), (0, 1), (1, 0), (0, 0)])
def test_writer_equals_the_reference_writer_byte_for_byte(tmp_path, n, is_mesh, binary):
    ref = _ref()
    if is_mesh and n % 3:
        pytest.skip("with faces the reference reads past its index array when the vertex count is not a multiple of three (src/PointCloudMap.cc:339-346,403-418)")
    xyz, bgra, nrm, label, kfid = _cloud(n, n + 2 * is_mesh + binary)
    a, b = tmp_path / "ours.ply", tmp_path / "ref.ply"
    T.save_map_ply(a, xyz, bgra, nrm, label, kfid, bool(is_mesh), bool(binary))
    assert ref.ref_write_map_ply(str(b).encode(), xyz.ctypes.data, bgra.ctypes.data, nrm.ctypes.data, label.ctypes.data, kfid.ctypes.data, n, is_mesh, binary) == 0
    assert filecmp.cmp(a, b, shallow=False), (n, is_mesh, binary)


@pytest.mark.parametrize("binary", [1, 0])
def test_reader_returns_what_was_written(tmp_path, binary):
    """binary files hold every value exactly; ASCII files hold 6 significant digits (the reference writes floats with the default ostream precision)"""
    ref = _ref()
    n = 777
    xyz, bgra, nrm, label, kfid = _cloud(n, 5)
    f = tmp_path / "m.ply"
    if ref is not None:
        assert ref.ref_write_map_ply(str(f).encode(), xyz.ctypes.data, bgra.ctypes.data, nrm.ctypes.data, label.ctypes.data, kfid.ctypes.data, n, 1, binary) == 0
    else:
        T.save_map_ply(f, xyz, bgra, nrm, label, kfid, True, bool(binary))
    m = T.load_map_ply(f)
    assert m["fields"] == 31 and len(m["xyz"]) == n
    # what the file calls red, green, blue: the binary writer stores PCL's first three bytes (b, g, r), the ASCII writer r, g, b
    want_rgb = bgra[:, :3] if binary else bgra[:, 2::-1]
    assert np.array_equal(m["rgb"], want_rgb) and np.array_equal(m["label"], label) and np.array_equal(m["kfid"], kfid)
    if binary:
        assert np.array_equal(m["xyz"].view(np.uint32), xyz.view(np.uint32)) and np.array_equal(m["normals"].view(np.uint32), nrm.view(np.uint32))
    else:
        assert np.allclose(m["xyz"], xyz, rtol=1e-5, atol=0) and np.allclose(m["normals"], nrm, rtol=1e-5, atol=1e-12)


def test_reader_on_other_ply_files_and_errors(tmp_path):
    """properties are matched by name: a PLY with double coordinates, an extra property and no normals / ids; plus the error returns"""
    f = tmp_path / "other.ply"
    f.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\nproperty double x\nproperty double y\nproperty double z\nproperty float intensity\n"
                 "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n"
                 "1.5 -2 3 0.25 10 20 30\n4 5 6e-1 0.5 255 0 1\n")
    m = T.load_map_ply(f)
    assert m["fields"] == 3 and np.allclose(m["xyz"], [[1.5, -2, 3], [4, 5, 0.6]]) and np.array_equal(m["rgb"], [[10, 20, 30], [255, 0, 1]])
    from plvs_b200 import _lib
    lib = _lib.load()
    n, fl = C.c_longlong(), C.c_int()
    assert lib.plvs_map_load_ply(str(tmp_path / "missing.ply").encode(), None, None, None, None, None, 0, C.byref(n), C.byref(fl)) == -1
    bad = tmp_path / "bad.ply"; bad.write_text("plx\n")
    assert lib.plvs_map_load_ply(str(bad).encode(), None, None, None, None, None, 0, C.byref(n), C.byref(fl)) == -1
    trunc = tmp_path / "trunc.ply"; trunc.write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nend_header\n1 2 3\n")
    buf = np.zeros((3, 3), np.float32)
    assert lib.plvs_map_load_ply(str(trunc).encode(), buf.ctypes.data, None, None, None, None, 3, C.byref(n), C.byref(fl)) == -1
