"""Known-answer vectors for the unit-cost edit distance (kernel (b') and its oracle): textbook examples plus seeded pairs
whose distance and path come from the DEFINITION -- a plain O(nm) Wagner-Fischer matrix in pure Python, walked back
preferring the diagonal, then 'I' (consumes the query), then 'D' -- independent of oracle/isof_oracle.c and of the CUDA
kernel.  (edlib, which the reference declares in setup.py:81 and never calls, is not installable here; the integer
distance is algorithm-independent.)      python scripts/make_edit_golden.py  ->  tests/golden/edit_distance.json"""
import json
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def wagner_fischer(a, b):
    n, m = len(a), len(b)
    D = [list(range(m + 1))] + [[i] + [0] * m for i in range(1, n + 1)]
    for i in range(1, n + 1):
        Di, Dp, ai = D[i], D[i - 1], a[i - 1]
        for j in range(1, m + 1):
            v = Dp[j - 1] + (ai != b[j - 1])
            if Dp[j] + 1 < v:
                v = Dp[j] + 1
            if Di[j - 1] + 1 < v:
                v = Di[j - 1] + 1
            Di[j] = v
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
    cigar, k = "", 0
    while k < len(ops):
        e = k
        while e < len(ops) and ops[e] == ops[k]:
            e += 1
        cigar += "%d%s" % (e - k, ops[k])
        k = e
    return D[n][m], cigar


def mutate(rng, s, rate, alphabet):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(rng.choice(alphabet)); continue
        out.append(ch)
        if r > 1 - rate / 3:
            out.append(rng.choice(alphabet))
    return "".join(out)


def main():
    rng = random.Random(20261020)
    cases = []
    for a, b in [("kitten", "sitting"), ("Saturday", "Sunday"), ("intention", "execution"), ("", "abc"), ("abc", ""), ("", ""),
                 ("GATTACA", "GATTACA"), ("ACGT", "TGCA"), ("a", "b"), ("flaw", "lawn"), ("gumbo", "gambol"),
                 ("AGGCTATCACCTGACCTCCAGGCCGATGCCC", "TAGCTATCACGACCGCGGTCGATTTGCCCGAC")]:
        cases.append((a, b))
    for n in [1, 2, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256, 257, 300, 511, 512, 513, 700, 1023, 1024, 1025, 1300]:
        alphabet = "ACGT" if n % 3 else "ACGTN"
        s = "".join(rng.choice(alphabet) for _ in range(n))
        cases.append((s, mutate(rng, s, 0.06, alphabet)))
        if n % 2:
            cases.append((mutate(rng, s, 0.3, alphabet), s))
        if n in (64, 257, 700):
            cases.append((s, "".join(rng.choice(alphabet) for _ in range(n + 9))))
            cases.append((s[n // 3:], s))
    out = []
    for a, b in cases:
        d, cigar = wagner_fischer(a, b)
        out.append({"query": a, "target": b, "distance": d, "cigar": cigar})
    path = os.path.join(ROOT, "tests", "golden", "edit_distance.json")
    with open(path, "w") as fh:
        json.dump({"about": "unit-cost global edit distance: plain Wagner-Fischer matrix in pure Python (scripts/make_edit_golden.py); "
                            "path walked back diagonal > I > D", "cases": out}, fh, indent=0)
    print(len(out), "cases ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
