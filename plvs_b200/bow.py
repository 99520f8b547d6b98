"""Host-side mirror of the reference's ORBVocabulary (include/ORBVocabulary.h = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) for the calls on
the path: loadFromTextFile (src/System.cc) and transform(features, BowVector, FeatureVector, levelsup) as Frame::ComputeBoW makes it
(src/Frame.cc:1498-1505).  The tree descents and the FeatureVector grouping run on the device (plvs_b200/csrc/bow.cu)."""
import ctypes as C
import numpy as np

from . import _lib


def bow_vector(scoring, weighting, word, weight):
    """the BowVector step alone (host arithmetic in the reference's order): ascending word ids, values"""
    lib = _lib.load()
    word = np.ascontiguousarray(word, np.uint32); weight = np.ascontiguousarray(weight, np.float64)
    n = len(word)
    ids = np.zeros(max(n, 1), np.uint32); vals = np.zeros(max(n, 1), np.float64); m = C.c_int()
    _lib.check(lib.plvs_bow_vector(scoring, weighting, word.ctypes.data_as(C.c_void_p), weight.ctypes.data_as(C.c_void_p), n, ids.ctypes.data_as(C.c_void_p),
                                   vals.ctypes.data_as(C.c_void_p), C.byref(m)), "plvs_bow_vector")
    return ids[:m.value], vals[:m.value]


class ORBVocabulary:
    def __init__(self, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self._device = device

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.plvs_voc_destroy(self._h); self._h = None

    def loadFromTextFile(self, path):
        """TemplatedVocabulary::loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1351-1436) -> bool"""
        if self._h:
            self._lib.plvs_voc_destroy(self._h); self._h = C.c_void_p()
        return self._lib.plvs_voc_load_text(str(path).encode(), self._device, C.byref(self._h)) == 0

    def create(self, k, L, scoring, weighting, parent, word_id, desc, weight):
        parent = np.ascontiguousarray(parent, np.int32); word_id = np.ascontiguousarray(word_id, np.int32)
        desc = np.ascontiguousarray(desc, np.uint8); weight = np.ascontiguousarray(weight, np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(self._lib.plvs_voc_create(self._device, k, L, scoring, weighting, len(parent), p(parent), p(word_id), p(desc), p(weight), C.byref(self._h)),
                   "plvs_voc_create")

    def size(self):
        return self._lib.plvs_voc_size(self._h)

    def transform(self, desc, levelsup=4, on_device=False, n=None):
        """-> dict(word, weight, node per feature; bow_ids, bow_vals = BowVector; fv_nodes, fv_offsets, fv_features = FeatureVector; fv_device)"""
        if on_device:
            ptr = C.c_void_p(int(desc))
        else:
            d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
            n = len(d); ptr = d.ctypes.data_as(C.c_void_p)
        m = max(n, 1)
        word = np.zeros(m, np.uint32); weight = np.zeros(m, np.float64); node = np.zeros(m, np.uint32)
        bi = np.zeros(m, np.uint32); bv = np.zeros(m, np.float64); nb = C.c_int()
        fvn = np.zeros(m, np.uint32); off = np.zeros(n + 2, np.int32); feat = np.zeros(m, np.int32); nn = C.c_int()
        dev = _lib.FeatVec()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(self._lib.plvs_voc_transform(self._h, ptr, n, int(bool(on_device)), levelsup, p(word), p(weight), p(node), p(bi), p(bv), C.byref(nb),
                                                p(fvn), p(off), p(feat), C.byref(nn), C.byref(dev)), "plvs_voc_transform")
        k = nn.value
        return dict(word=word[:n], weight=weight[:n], node=node[:n], bow_ids=bi[:nb.value], bow_vals=bv[:nb.value], fv_nodes=fvn[:k], fv_offsets=off[:k + 1],
                    fv_features=feat[:off[k]], fv_device=dev)
