"""Host-side expectation of the voxel-block merge: fold per-map downloads (keys, sdf, w) in list order exactly as
k_merge_fold does -- per voxel with w_in > 0: acc = (w > 0 ? w*sdf : 0) + w_in*sdf_in ; w += w_in ; sdf = acc / w."""
import numpy as np


def fold(maps):
    state = {}
    for keys, sdf, wt in maps:
        for k, s, wv in zip(map(tuple, keys), sdf, wt):
            a = np.where(wv > 0, wv * s, 0).astype(np.float32)
            if k not in state:
                state[k] = (np.zeros(4096, np.float32), np.full(4096, 99999.0, np.float32))
            w0, s0 = state[k]
            hit = wv > 0
            acc = (np.where(w0 > 0, w0 * s0, 0).astype(np.float32) + a).astype(np.float32)
            w1 = (w0 + wv).astype(np.float32)
            s1 = np.where(hit, acc / np.where(hit, w1, 1), s0).astype(np.float32)
            state[k] = (np.where(hit, w1, w0).astype(np.float32), s1)
    return state


def compare(state, keys, sdf, wt, atol=1e-6):
    """number of blocks that differ between the folded expectation and a download"""
    got = {tuple(k): (s, w) for k, s, w in zip(keys, sdf, wt)}
    bad = int(set(got) != set(state))
    for k, (s, w) in got.items():
        if k not in state:
            continue
        ew, es = state[k]
        if not np.array_equal(w, ew) or not np.allclose(s[ew > 0], es[ew > 0], rtol=0, atol=atol) or not np.all(s[ew == 0] == 99999.0):
            bad += 1
    return bad
