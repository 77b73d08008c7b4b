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


def _path_py(a, b):
    """Plain matrix + traceback (diagonal, then up = 'I', then left = 'D') in Python, small sizes only."""
    n, m = len(a), len(b)
    D = [[0] * (m + 1) for _ in range(n + 1)]
    for j in range(m + 1):
        D[0][j] = j
    for i in range(1, n + 1):
        D[i][0] = i
        for j in range(1, m + 1):
            D[i][j] = min(D[i - 1][j - 1] + (a[i - 1] != b[j - 1]), D[i - 1][j] + 1, D[i][j - 1] + 1)
    ops = []
    i, j = n, m
    while i > 0 or j > 0:
        if i > 0 and j > 0 and D[i][j] == D[i - 1][j - 1] + (a[i - 1] != b[j - 1]):
            ops.append("X" if a[i - 1] != b[j - 1] else "=")
            i -= 1; j -= 1
        elif i > 0 and (j == 0 or D[i][j] == D[i - 1][j] + 1):
            ops.append("I"); i -= 1
        else:
            ops.append("D"); j -= 1
    ops.reverse()
    out, k = "", 0
    while k < len(ops):
        e = k
        while e < len(ops) and ops[e] == ops[k]:
            e += 1
        out += "%d%s" % (e - k, ops[k])
        k = e
    return D[n][m], out


def test_edit_path_vs_python_and_known_answers():
    rng = np.random.default_rng(8)
    for a, b, d in [("kitten", "sitting", 3), ("Saturday", "Sunday", 3), ("", "abc", 3), ("abc", "", 3), ("", "", 0),
                    ("intention", "execution", 5)]:
        got_d, runs = oracle.edit_path(a, b)
        assert got_d == d == oracle.edit_distance(a, b)
        assert oracle.cigar_runs_to_string(runs) == _path_py(a, b)[1]
    for _ in range(60):
        n, m = int(rng.integers(0, 90)), int(rng.integers(0, 90))
        a = "".join("ACGT"[x] for x in rng.integers(0, 4, n))
        b = "".join("ACGT"[x] for x in rng.integers(0, 4, m)) if rng.random() < 0.4 else a[m // 3:] + a[:m // 4]
        d, runs = oracle.edit_path(a, b)
        want_d, want = _path_py(a, b)
        assert d == want_d == oracle.edit_distance(a, b)
        assert oracle.cigar_runs_to_string(runs) == want


def test_oracle_against_the_committed_golden_vectors():
    """tests/golden/edit_distance.json: distances and paths from the definition (pure-Python Wagner-Fischer,
    scripts/make_edit_golden.py) -- the oracle's Myers restatement and its path function must reproduce them."""
    from conftest import load_golden
    g = load_golden("edit_distance.json")
    assert len(g["cases"]) >= 50
    for c in g["cases"]:
        assert oracle.edit_distance(c["query"], c["target"]) == c["distance"]
        d, runs = oracle.edit_path(c["query"], c["target"])
        assert d == c["distance"] and oracle.cigar_runs_to_string(runs) == c["cigar"]
