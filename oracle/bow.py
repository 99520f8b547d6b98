"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/bow_oracle.cpp (the restated DBoW2 transform) and of the reference's own DBoW2
(oracle/_ref/libbow_ref.so, oracle/ref_build.py), plus a generator of synthetic vocabularies in the ORBvoc.txt text format."""
import ctypes as C
import numpy as np

from .orb import lib


def write_vocabulary(path, k, L, seed=0, scoring=0, weighting=0, zero_weight_fraction=0.0, clustered=True):
    """a full k-ary tree of depth L in the format TemplatedVocabulary::saveToTextFile writes (one line per node: parent, leaf flag, 32 bytes, weight),
    WITHOUT a trailing newline.  Children are noisy copies of their parent (like k-means centres) unless clustered is False."""
    rng = np.random.default_rng(seed)
    lines = ["%d %d %d %d" % (k, L, scoring, weighting)]
    nodes = [(0, rng.integers(0, 256, 32, dtype=np.uint8))]          # (depth, descriptor) of the root, id 0
    frontier = [0]
    for depth in range(1, L + 1):
        nxt = []
        for pid in frontier:
            for _ in range(k):
                if clustered:
                    flip = np.packbits((rng.random(256) < 0.5 / depth).astype(np.uint8))
                    d = nodes[pid][1] ^ flip
                else:
                    d = rng.integers(0, 256, 32, dtype=np.uint8)
                nid = len(nodes)
                nodes.append((depth, d))
                leaf = depth == L
                w = 0.0 if (leaf and rng.random() < zero_weight_fraction) else (float(rng.uniform(0.5, 9.0)) if leaf else 0.0)
                lines.append("%d %d %s %r" % (pid, 1 if leaf else 0, " ".join(str(int(b)) for b in d), w))
                nxt.append(nid)
        frontier = nxt
    with open(path, "w") as f:
        f.write("\n".join(lines))
    return len(nodes)


def _transform(fn, h, desc, levelsup):
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(d)
    word = np.zeros(max(n, 1), np.uint32); weight = np.zeros(max(n, 1), np.float64); node = np.zeros(max(n, 1), np.uint32)
    bi = np.zeros(max(n, 1), np.uint32); bv = np.zeros(max(n, 1), np.float64)
    fn_nodes = np.zeros(max(n, 1), np.uint32); off = np.zeros(n + 2, np.int32); feat = np.zeros(max(n, 1), np.int32); nn = C.c_int()
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_int)] + [C.c_void_p] * 3
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    nb = fn(h, p(d), n, levelsup, p(word), p(weight), p(node), p(bi), p(bv), C.byref(nn), p(fn_nodes), p(off), p(feat))
    m = nn.value
    return dict(word=word[:n], weight=weight[:n], node=node[:n], bow_ids=bi[:nb], bow_vals=bv[:nb], fv_nodes=fn_nodes[:m], fv_offsets=off[:m + 1],
                fv_features=feat[:off[m]])


class Vocabulary:
    """the restatement (oracle/bow_oracle.cpp)"""
    def __init__(self, path):
        self._l = lib()
        self._l.orc_voc_load.restype = C.c_void_p; self._l.orc_voc_load.argtypes = [C.c_char_p]
        self._l.orc_voc_destroy.argtypes = [C.c_void_p]; self._l.orc_voc_size.argtypes = [C.c_void_p]
        self._h = self._l.orc_voc_load(str(path).encode())
        assert self._h, path

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.orc_voc_destroy(self._h); self._h = None

    def size(self):
        return self._l.orc_voc_size(self._h)

    def transform(self, desc, levelsup=4):
        return _transform(self._l.orc_voc_transform, self._h, desc, levelsup)

    def export(self):
        """flat arrays for plvs_voc_create: (k, L, scoring, weighting), parent[n], word_id[n] (-1 = inner node), desc[n,32], weight[n]"""
        self._l.orc_voc_export.argtypes = [C.c_void_p] * 6
        hdr = np.zeros(4, np.int32)
        n = self._l.orc_voc_export(self._h, None, None, None, None, hdr.ctypes.data_as(C.c_void_p))
        parent = np.zeros(n, np.int32); wid = np.zeros(n, np.int32); desc = np.zeros((n, 32), np.uint8); w = np.zeros(n, np.float64)
        self._l.orc_voc_export(self._h, parent.ctypes.data_as(C.c_void_p), wid.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p),
                               w.ctypes.data_as(C.c_void_p), hdr.ctypes.data_as(C.c_void_p))
        return tuple(int(x) for x in hdr), parent, wid, desc, w


def ref_available():
    from . import ref_build
    return ref_build.build_bow() is not None


class RefVocabulary:
    """the reference's own DBoW2 (oracle/_ref/libbow_ref.so)"""
    _lib = None

    def __init__(self, path):
        if RefVocabulary._lib is None:
            from . import ref_build
            RefVocabulary._lib = C.CDLL(ref_build.build_bow())
        l = RefVocabulary._lib
        l.ref_voc_load.restype = C.c_void_p; l.ref_voc_load.argtypes = [C.c_char_p]
        l.ref_voc_destroy.argtypes = [C.c_void_p]; l.ref_voc_size.argtypes = [C.c_void_p]
        self._h = l.ref_voc_load(str(path).encode())
        assert self._h, path

    def __del__(self):
        if getattr(self, "_h", None):
            RefVocabulary._lib.ref_voc_destroy(self._h); self._h = None

    def size(self):
        return RefVocabulary._lib.ref_voc_size(self._h)

    def transform(self, desc, levelsup=4):
        return _transform(RefVocabulary._lib.ref_voc_transform, self._h, desc, levelsup)
