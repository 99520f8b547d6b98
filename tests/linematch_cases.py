"""Seeded query / train descriptor sets for the line-descriptor k-NN tests: random, near-duplicate (almost everything ties), sparse (few bits set),
perturbed copies of the train set -- with and without a query mask."""
import numpy as np


def cases(n_trials=24, seed=1, nq_max=80, nt_max=120):
    rng = np.random.default_rng(seed)
    out = []
    for trial in range(n_trials):
        nq = int(rng.integers(1, nq_max)); nt = int(rng.integers(2, nt_max))
        kind = trial % 4
        if kind == 0:
            q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        elif kind == 1:
            base = rng.integers(0, 256, (4, 32), dtype=np.uint8)
            t = base[rng.integers(0, 4, nt)].copy(); q = base[rng.integers(0, 4, nq)].copy()
            for a in (t, q):
                for r in range(len(a)):
                    for _ in range(int(rng.integers(0, 3))):
                        a[r, rng.integers(0, 32)] ^= np.uint8(1 << rng.integers(0, 8))
        elif kind == 2:
            t = np.zeros((nt, 32), np.uint8); q = np.zeros((nq, 32), np.uint8)
            for a in (t, q):
                for r in range(len(a)):
                    for _ in range(int(rng.integers(0, 6))):
                        a[r, rng.integers(0, 32)] |= np.uint8(1 << rng.integers(0, 8))
        else:
            t = rng.integers(0, 256, (nt, 32), dtype=np.uint8); q = t[rng.integers(0, nt, nq)].copy()
            for r in range(nq):
                for _ in range(int(rng.integers(0, 20))):
                    q[r, rng.integers(0, 32)] ^= np.uint8(1 << rng.integers(0, 8))
        mask = None if trial % 3 else (rng.random(nq) < 0.7).astype(np.uint8)
        out.append((q, t, mask))
    return out
