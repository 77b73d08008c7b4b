"""The Myers/Hyyro edit-distance restatement in oracle/ (config 5's "edlib CPU" timing column; edlib has no
call site in the reference) against a plain O(nm) dynamic program."""
import numpy as np

import oracle


def _dp(a, b):
    prev = np.arange(len(b) + 1)
    bb = np.frombuffer(b.encode(), np.uint8)
    for i, ch in enumerate(a.encode(), 1):
        sub = prev[:-1] + (bb != ch)
        cur = np.minimum(sub, prev[1:] + 1)
        cur = np.concatenate([[i], cur])
        # resolve the left-to-right dependency D[i][j-1] + 1
        cur = np.minimum.accumulate(cur - np.arange(len(cur))) + np.arange(len(cur))
        prev = cur
    return int(prev[-1])


def test_edit_distance_vs_plain_dp():
    rng = np.random.default_rng(5)
    shapes = [(0, 0), (0, 7), (5, 0), (1, 1), (63, 64), (64, 64), (65, 63), (128, 129), (200, 77), (77, 200), (511, 640)]
    for (n, m) in shapes + [(int(rng.integers(1, 400)), int(rng.integers(1, 400))) for _ in range(40)]:
        a = "".join("ACGT"[x] for x in rng.integers(0, 4, n))
        b = "".join("ACGT"[x] for x in rng.integers(0, 4, m))
        assert oracle.edit_distance(a, b) == _dp(a, b), (n, m)
        if n:
            c = list(a)
            for _ in range(max(1, n // 15)):
                c[int(rng.integers(0, n))] = "ACGT"[int(rng.integers(0, 4))]
            c = "".join(c)[: max(1, n - 3)]
            assert oracle.edit_distance(a, c) == _dp(a, c)
            assert oracle.edit_distance(c, a) == oracle.edit_distance(a, c)
    assert oracle.edit_distance("ACGT" * 50, "ACGT" * 50) == 0


def test_edit_distance_pairs_batch():
    rng = np.random.default_rng(6)
    seqs = ["".join("ACGT"[x] for x in rng.integers(0, 4, int(rng.integers(1, 300)))) for _ in range(30)]
    cat = np.frombuffer("".join(seqs).encode(), np.uint8)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    ia = rng.integers(0, 30, 100).astype(np.int32)
    ib = rng.integers(0, 30, 100).astype(np.int32)
    d = oracle.edit_distance_pairs(cat, off, ia, ib, threads=2)
    for p in range(0, 100, 7):
        assert int(d[p]) == _dp(seqs[ia[p]], seqs[ib[p]])
