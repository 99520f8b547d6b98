"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the line-descriptor k-NN the reference calls in LineMatcher::ComputeDescriptorMatches
(/root/reference/src/LineMatcher.cc:2567-2615): cv::line_descriptor_c::BinaryDescriptorMatcher::knnMatch(query, train, matches, 2, mask, true)
(/root/reference/Thirdparty/line_descriptor/src/binary_descriptor_matcher_custom.cpp:258-337) followed by the ratio test.

The library answers the query with multi-index hashing (Mihasher(256, 32): 32 substrings of 8 bits = the 32 bytes of the descriptor, :596-788).  The
result is the k nearest train descriptors by Hamming distance; AMONG EQUAL DISTANCES the order is the order in which Mihasher::query discovers the
items, and the callers' ratio test reads it.  Restated: an item t is first met at the smallest substring radius s* = min_k popcount(q[k] ^ t[k]),
in the first substring k* that attains it (:668-673 outer loops), at the position of the byte q[k*] ^ t[k*] in the library's enumeration of the
8-bit strings with s* ones (:683-760), and inside a bucket in insertion order = ascending train index (:801-818, BucketGroup::insert :928-950).
So the neighbours are the k smallest (hamming, s*, k*, enumeration rank, train index) tuples.  Pinned bit-exactly to the compiled library by
tests/test_oracle_vs_reference_linematch.py; the product never imports this module."""
import ctypes as C
import pathlib
import numpy as np

HERE = pathlib.Path(__file__).resolve().parent


def enumeration_rank():
    """rank[mask] of every 8-bit string in Mihasher::query's enumeration of the strings with popcount(mask) ones (its `power` / `bit` loop, :683-760)"""
    rank = np.zeros(256, np.int32)
    curb = 8
    for s in range(0, 9):
        power = [0] * 100
        bitstr = 0
        for i in range(s):
            power[i] = i
        power[s] = curb + 1
        bit = s - 1
        r = 0
        while True:
            if bit != -1:
                bitstr ^= (1 << power[bit]) if power[bit] == bit else (3 << (power[bit] - 1))
                power[bit] += 1
                bit -= 1
            else:
                rank[bitstr & 0xff] = r
                r += 1
                while True:
                    bit += 1
                    if bit < s and power[bit] == power[bit + 1] - 1:
                        bitstr ^= 1 << (power[bit] - 1)
                        power[bit] = bit
                    else:
                        break
                if bit == s:
                    break
    return rank


_POP8 = np.array([bin(i).count("1") for i in range(256)], np.int64)
_RANK = None


def knn2(query, train, mask=None, nn_ratio=0.78):
    """-> (query_idx[M], train_idx[M,2], dist[M,2] float32, valid[M] uint8, n_valid): the rows of lmatches (compact result: masked-out queries dropped)
    and vValidMatch of LineMatcher::ComputeDescriptorMatches"""
    global _RANK
    if _RANK is None:
        _RANK = enumeration_rank().astype(np.int64)
    q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
    nt = len(t)
    assert nt >= 2, "the library reads uninitialised results with fewer train descriptors than k"
    qi, ti, di, vi = [], [], [], []
    for i in range(len(q)):
        if mask is not None and not mask[i]:
            continue
        x = q[i][None, :] ^ t                       # nt x 32
        pc = _POP8[x]                               # per-substring distances
        ham = pc.sum(1)
        smin = pc.min(1)
        kfirst = pc.argmin(1)                       # first substring that attains the minimum
        rank = _RANK[x[np.arange(nt), kfirst]]
        key = ((((ham * 16 + smin) * 32 + kfirst) * 128 + rank) << 20) + np.arange(nt)
        best = np.argsort(key, kind="stable")[:2]
        qi.append(i); ti.append(best.astype(np.int32)); di.append(ham[best].astype(np.float32))
        vi.append(1 if np.float32(ham[best[0]]) < np.float32(nn_ratio) * np.float32(ham[best[1]]) else 0)
    m = len(qi)
    return (np.asarray(qi, np.int32), np.asarray(ti, np.int32).reshape(m, 2), np.asarray(di, np.float32).reshape(m, 2), np.asarray(vi, np.uint8), int(np.sum(vi)))


class RefLineMatcher:
    """the reference's own matcher (oracle/_ref/liblinematch_ref.so)"""

    def __init__(self):
        from . import ref_build
        path = ref_build.build_linematch()
        if path is None:
            raise RuntimeError("oracle/_ref/liblinematch_ref.so is not built and /root/reference is absent")
        self.lib = C.CDLL(path)
        self.lib.ref_line_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]

    def knn2(self, query, train, mask=None, nn_ratio=0.78):
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        nq = len(q)
        qi = np.zeros(nq, np.int32); ti = np.zeros((nq, 2), np.int32); di = np.zeros((nq, 2), np.float32); vi = np.zeros(nq, np.uint8); nv = C.c_int()
        mk = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        import os, sys
        sys.stdout.flush()
        m = self.lib.ref_line_knn2(q.ctypes.data, nq, t.ctypes.data, len(t), None if mk is None else mk.ctypes.data, nn_ratio, qi.ctypes.data, ti.ctypes.data, di.ctypes.data,
                                   vi.ctypes.data, C.byref(nv))
        return qi[:m], ti[:m], di[:m], vi[:m], nv.value
