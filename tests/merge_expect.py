"""Host-side expectation of the voxel-block merge: fold per-map downloads (keys, sdf, w[, rgba]) in list order exactly as
k_merge_fold does -- per voxel with w_in > 0: acc = (w > 0 ? w*sdf : 0) + w_in*sdf_in ; w += w_in ; sdf = acc / w; colour (r,g,b,cw):
an empty voxel takes the incoming colour, otherwise channel = (uint8)((float)(cw*c + cw_in*c_in) * (1.f / (cw + cw_in))), cw = min(cw + cw_in, 255)."""
import numpy as np


def fold(maps):
    state = {}
    for m in maps:
        keys, sdf, wt = m[0], m[1], m[2]
        rgba = m[3] if len(m) > 3 else [None] * len(keys)
        for k, s, wv, cv in zip(map(tuple, keys), sdf, wt, rgba):
            a = np.where(wv > 0, wv * s, 0).astype(np.float32)
            if k not in state:
                state[k] = (np.zeros(4096, np.float32), np.full(4096, 99999.0, np.float32), np.zeros((4096, 4), np.uint8))
            w0, s0, c0 = state[k]
            hit = wv > 0
            acc = (np.where(w0 > 0, w0 * s0, 0).astype(np.float32) + a).astype(np.float32)
            w1 = (w0 + wv).astype(np.float32)
            s1 = np.where(hit, acc / np.where(hit, w1, 1), s0).astype(np.float32)
            c1 = c0
            if cv is not None:
                cw0, cwi = c0[:, 3].astype(np.uint32), cv[:, 3].astype(np.uint32)
                inv = (np.float32(1.0) / np.maximum(cw0 + cwi, 1).astype(np.float32)).astype(np.float32)
                mix = ((cw0[:, None] * c0[:, :3] + cwi[:, None] * cv[:, :3].astype(np.uint32)).astype(np.float32) * inv[:, None]).astype(np.float32)
                mixed = np.concatenate([mix.astype(np.uint32).astype(np.uint8), np.minimum(cw0 + cwi, 255).astype(np.uint8)[:, None]], 1)
                c1 = np.where((cwi == 0)[:, None], c0, np.where((cw0 == 0)[:, None], cv, mixed)).astype(np.uint8)
            state[k] = (np.where(hit, w1, w0).astype(np.float32), s1, c1)
    return state


def compare(state, keys, sdf, wt, rgba=None, atol=1e-6):
    """number of blocks that differ between the folded expectation and a download"""
    got = {tuple(k): i for i, k in enumerate(keys)}
    bad = int(set(got) != set(state))
    for k, i in got.items():
        if k not in state:
            continue
        s, w = sdf[i], wt[i]
        ew, es, ec = state[k]
        if not np.array_equal(w, ew) or not np.allclose(s[ew > 0], es[ew > 0], rtol=0, atol=atol) or not np.all(s[ew == 0] == 99999.0):
            bad += 1
        elif rgba is not None and not np.array_equal(rgba[i], ec):
            bad += 1
    return bad
