"""Kernel (b'), csrc/myers.cu: batched global unit-cost edit distance (Myers/Hyyro bit-parallel) through the C ABI against
the oracle's Myers restatement (itself checked against a plain O(nm) DP in test_oracle_edit_distance.py) and, for the
path, against the oracle's plain-matrix traceback with the same tie rule.  edlib has no call site in the reference
(setup.py:81 only): the integer distance is the contract; the textbook known answers pin both sides."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

KNOWN = [("kitten", "sitting", 3), ("Saturday", "Sunday", 3), ("", "abc", 3), ("abc", "", 3), ("", "", 0),
         ("GATTACA", "GATTACA", 0), ("intention", "execution", 5), ("ACGT", "TGCA", 4), ("a", "b", 1)]


@pytest.fixture(scope="module")
def ctx():
    from isonform_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def _rand(rng, n, alphabet="ACGT"):
    return "".join(alphabet[int(x)] for x in rng.integers(0, len(alphabet), size=n))


def _mutate(rng, s, rate, alphabet="ACGT"):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(alphabet[int(rng.integers(0, len(alphabet)))])
            continue
        out.append(ch)
        if r > 1 - rate / 3:
            out.append(alphabet[int(rng.integers(0, len(alphabet)))])
    return "".join(out)


def _pool(seqs):
    from isonform_b200._lib import pack_sequences
    return pack_sequences(seqs)


def _apply(runs, q, t):
    """Walk a (len<<4|op) run list over q and t: returns (cost, consumed_q, consumed_t, letters_ok)."""
    i = j = cost = 0
    ok = True
    for r in runs:
        ln, op = int(r) >> 4, int(r) & 15
        if op == 7:
            ok &= q[i:i + ln] == t[j:j + ln]
            i += ln; j += ln
        elif op == 8:
            ok &= all(a != b for a, b in zip(q[i:i + ln], t[j:j + ln]))
            i += ln; j += ln; cost += ln
        elif op == 1:
            i += ln; cost += ln
        elif op == 2:
            j += ln; cost += ln
        else:
            ok = False
    return cost, i, j, ok


def test_known_answers(ctx):
    from isonform_b200 import hotpath
    for a, b, d in KNOWN:
        assert oracle.edit_distance(a, b) == d
        assert hotpath.edlib_align(a, b, ctx=ctx)["editDistance"] == d, (a, b)
        r = hotpath.edlib_align(a, b, task="path", ctx=ctx)
        assert r["editDistance"] == d
        assert r["cigar"] == oracle.cigar_runs_to_string(oracle.edit_path(a, b)[1]), (a, b)


def test_golden_vectors(ctx):
    """tests/golden/edit_distance.json (pure-Python Wagner-Fischer, scripts/make_edit_golden.py): distances, paths, and the
    distances again under a bound."""
    from conftest import load_golden
    from isonform_b200 import hotpath
    g = load_golden("edit_distance.json")["cases"]
    seqs = [c["query"] for c in g] + [c["target"] for c in g]
    pa = list(range(len(g)))
    pb = [len(g) + i for i in pa]
    got = hotpath.edit_distance_pairs(seqs, pa, pb, task="path", ctx=ctx)
    for c, r in zip(g, got):
        assert r["editDistance"] == c["distance"] and r["cigar"] == (c["cigar"] or ""), (len(c["query"]), len(c["target"]))
    plain = hotpath.edit_distance_pairs(seqs, pa, pb, ctx=ctx)
    assert [r["editDistance"] for r in plain] == [c["distance"] for c in g]
    for k in (0, 10, 60):
        bounded = hotpath.edit_distance_pairs(seqs, pa, pb, k=k, ctx=ctx)
        assert [r["editDistance"] for r in bounded] == [c["distance"] if c["distance"] <= k else -1 for c in g]


def test_distance_vs_oracle_all_group_sizes_and_stripes(ctx):
    """Every class of the kernel (1, 2, 4, 8, 16, 32 lanes per pair; 1 to 4 stripes of 1024 rows), word and stripe
    boundaries, ragged targets inside a warp, similar and unrelated pairs."""
    rng = np.random.default_rng(101)
    seqs = [""]
    edges = [1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 159, 160, 161, 191, 192, 193, 223, 224, 225, 255, 256,
             257, 320, 321, 384, 385, 448, 449, 511, 512, 513, 640, 641, 768, 897, 1023, 1024, 1025, 1281, 1537, 1793, 2047,
             2048, 2049, 2561, 3100, 3585, 4097, 5121, 6145, 7169, 8191, 8192, 8193, 9000]
    for n in edges:
        s = _rand(rng, n)
        seqs += [s, _mutate(rng, s, 0.1), _mutate(rng, s, 0.3)]
    for _ in range(120):
        n = int(rng.integers(1, 700))
        s = _rand(rng, n)
        seqs += [s, _mutate(rng, s, float(rng.random()) * 0.4)]
    seqs += ["A" * 300, "A" * 1500, "AC" * 700, "ACGT" * 1000, "T" * 40]
    cat, off = _pool(seqs)
    S = len(seqs)
    pa = rng.integers(0, S, 6000).astype(np.int32)
    pb = rng.integers(0, S, 6000).astype(np.int32)
    # neighbours (a sequence and its mutated copies) so that small distances occur in every class
    nb = np.arange(S - 1, dtype=np.int32)
    pa = np.concatenate([pa, nb, nb + 1, nb])
    pb = np.concatenate([pb, nb + 1, nb, nb])
    got = ctx.edit_distance_pairs(cat, off, pa, pb)["dist"]
    want = oracle.edit_distance_pairs(cat, off, pa, pb, threads=8)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, [(int(p), len(seqs[pa[p]]), len(seqs[pb[p]]), int(got[p]), int(want[p])) for p in bad[:10]]


def test_arbitrary_bytes_and_case(ctx):
    """Bytes compare by value: lower case differs from upper case, N equals only N, any byte value is a symbol."""
    rng = np.random.default_rng(7)
    seqs = [bytes(rng.integers(0, 256, int(rng.integers(0, 400))).astype(np.uint8)) for _ in range(60)]
    seqs += [b"ACGTNacgtn" * 20, b"acgtnACGTN" * 20, b"\x00" * 70, b"\x00\xff" * 35, b"N" * 100, b"n" * 100]
    cat, off = _pool(seqs)
    S = len(seqs)
    pa = np.repeat(np.arange(S, dtype=np.int32), S)
    pb = np.tile(np.arange(S, dtype=np.int32), S)
    got = ctx.edit_distance_pairs(cat, off, pa, pb)["dist"]
    want = oracle.edit_distance_pairs(cat, off, pa, pb, threads=8)
    assert (got == want).all()
    assert got[(S - 2) * S + (S - 1)] == 100           # 'N' * 100 vs 'n' * 100


def test_threshold_k(ctx):
    rng = np.random.default_rng(9)
    seqs = []
    for _ in range(100):
        s = _rand(rng, int(rng.integers(1, 1500)))
        seqs += [s, _mutate(rng, s, 0.05)]
    cat, off = _pool(seqs)
    pa = np.arange(0, 200, 2, dtype=np.int32)
    pb = pa + 1
    full = ctx.edit_distance_pairs(cat, off, pa, pb)["dist"]
    for k in (0, 5, 40, 10 ** 6):
        got = ctx.edit_distance_pairs(cat, off, pa, pb, k=k)["dist"]
        assert (got == np.where(full <= k, full, -1)).all()
    r = ctx.edit_distance_pairs(cat, off, pa, pb, k=5, want_cigar=True)
    n_runs = np.diff(r["cigar_off"])
    assert ((n_runs == 0) == (r["dist"] < 0)).all()            # no path above k; every other pair (non-empty) has runs


def test_path_vs_oracle(ctx):
    """Runs equal the oracle's plain-matrix traceback (diagonal, then I, then D) and are a valid witness of the distance."""
    rng = np.random.default_rng(33)
    seqs = ["", "A", "ACGT"]
    for n in [1, 31, 32, 33, 64, 65, 100, 128, 129, 256, 300, 512, 513, 1024, 1025, 1500, 2049, 2600]:
        s = _rand(rng, n)
        seqs += [s, _mutate(rng, s, 0.08), _mutate(rng, s, 0.3), s[n // 3:], s[: 2 * n // 3] + _rand(rng, 40)]
    seqs += ["A" * 200, "A" * 230, "AC" * 100, "CA" * 100]
    cat, off = _pool(seqs)
    S = len(seqs)
    pa = rng.integers(0, S, 700).astype(np.int32)
    pb = rng.integers(0, S, 700).astype(np.int32)
    nb = np.arange(S - 1, dtype=np.int32)
    pa = np.concatenate([pa, nb, nb + 1, nb])
    pb = np.concatenate([pb, nb + 1, nb, nb])
    r = ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
    want_d = oracle.edit_distance_pairs(cat, off, pa, pb, threads=8)
    assert (r["dist"] == want_d).all()
    for p in range(len(pa)):
        q, t = seqs[pa[p]], seqs[pb[p]]
        runs = r["cigar"][r["cigar_off"][p]:r["cigar_off"][p + 1]]
        cost, i, j, ok = _apply(runs, q, t)
        assert ok and i == len(q) and j == len(t) and cost == int(want_d[p]), (p, len(q), len(t))
        if len(q) * len(t) <= 1_500_000 or p % 9 == 0:
            d, want = oracle.edit_path(q, t)
            assert d == int(want_d[p])
            assert runs.tolist() == want.tolist(), (p, len(q), len(t))


def test_path_capacity_protocol_and_chunked_trace(ctx):
    """ISOF_ERR_CAPACITY reports the needed size; a trace budget smaller than the call forces chunks (the DP then runs
    once for the run counts and once for the runs) with identical output."""
    from isonform_b200 import _lib
    rng = np.random.default_rng(5)
    seqs = []
    for _ in range(60):
        s = _rand(rng, int(rng.integers(50, 900)))
        seqs += [s, _mutate(rng, s, 0.15)]
    cat, off = _pool(seqs)
    pa = rng.integers(0, len(seqs), 500).astype(np.int32)
    pb = rng.integers(0, len(seqs), 500).astype(np.int32)
    ref = ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
    L = ctx._L
    dist = np.empty(len(pa), np.int32)
    coff = np.zeros(len(pa) + 1, np.int64)
    small = np.empty(10, np.uint32)
    rc = L.isof_edit_distance_pairs(ctx._h, cat.ctypes.data, off.ctypes.data, len(off) - 1, pa.ctypes.data, pb.ctypes.data,
                                    len(pa), -1, dist.ctypes.data, small.ctypes.data, 10, coff.ctypes.data)
    assert rc == _lib.ERR_CAPACITY and coff[-1] == len(ref["cigar"]) and (coff == ref["cigar_off"]).all()
    c2 = _lib.Context(0)
    try:
        c2.set_trace_budget(3 << 20)
        r2 = c2.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
        assert c2.profile()["dp_launches"] >= 2
    finally:
        c2.close()
    assert (r2["dist"] == ref["dist"]).all() and (r2["cigar_off"] == ref["cigar_off"]).all()
    assert (r2["cigar"] == ref["cigar"]).all()


def test_bad_arguments(ctx):
    from isonform_b200 import _lib
    cat, off = _pool(["ACGT", "AC"])
    with pytest.raises(_lib.IsofError):
        ctx.edit_distance_pairs(cat, off, [0], [2])
    with pytest.raises(_lib.IsofError):
        ctx.edit_distance_pairs(cat, off, [-1], [0])
    assert len(ctx.edit_distance_pairs(cat, off, [], [])["dist"]) == 0


def test_large_batch_checksum_property(ctx):
    """Full-size property check (no oracle at this size in seconds): symmetry d(a,b) = d(b,a), d(a,a) = 0,
    | |a| - |b| | <= d <= max(|a|, |b|), triangle inequality on sampled triples; 200 k pairs of 100-300 bp."""
    rng = np.random.default_rng(77)
    base = [_rand(rng, int(rng.integers(100, 300))) for _ in range(40)]
    seqs = [_mutate(rng, base[i % 40], 0.1) for i in range(2000)]
    cat, off = _pool(seqs)
    lens = np.diff(off)
    pa = rng.integers(0, 2000, 200_000).astype(np.int32)
    pb = rng.integers(0, 2000, 200_000).astype(np.int32)
    d_ab = ctx.edit_distance_pairs(cat, off, pa, pb)["dist"]
    d_ba = ctx.edit_distance_pairs(cat, off, pb, pa)["dist"]
    assert (d_ab == d_ba).all()
    assert (d_ab[pa == pb] == 0).all()
    assert (d_ab >= np.abs(lens[pa] - lens[pb])).all() and (d_ab <= np.maximum(lens[pa], lens[pb])).all()
    pc = rng.integers(0, 2000, 200_000).astype(np.int32)
    d_bc = ctx.edit_distance_pairs(cat, off, pb, pc)["dist"]
    d_ac = ctx.edit_distance_pairs(cat, off, pa, pc)["dist"]
    assert (d_ac <= d_ab + d_bc).all()
    sub = rng.integers(0, 200_000, 3000)
    want = oracle.edit_distance_pairs(cat, off, pa[sub], pb[sub], threads=8)
    assert (d_ab[sub] == want).all()


def test_debug_build_assertions_stay_silent_for_myers():
    """The assertion build (-DISOF_CHECK) carries bounds checks on the trace stores / loads and the run writes of
    kernel (b'): all classes, multi-stripe pairs and a chunked call must leave the violation mask empty."""
    from isonform_b200 import _lib
    dctx = _lib.Context(0, debug=True)
    try:
        assert dctx.debug_violations() == (0, True)
        rng = np.random.default_rng(12)
        seqs = [""]
        for n in [1, 32, 33, 64, 100, 128, 129, 200, 256, 257, 500, 512, 513, 1024, 1100, 2048, 2100, 4096, 4100, 9000]:
            s = _rand(rng, n)
            seqs += [s, _mutate(rng, s, 0.1)]
        cat, off = _pool(seqs)
        S = len(seqs)
        pa = rng.integers(0, S, 400).astype(np.int32)
        pb = rng.integers(0, S, 400).astype(np.int32)
        nb = np.arange(S - 1, dtype=np.int32)
        pa = np.concatenate([pa, nb, nb + 1])
        pb = np.concatenate([pb, nb + 1, nb])
        want = oracle.edit_distance_pairs(cat, off, pa, pb, threads=8)
        for budget in (0, 48 << 20):
            dctx.set_trace_budget(budget)
            r = dctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
            assert (r["dist"] == want).all()
            assert (dctx.edit_distance_pairs(cat, off, pa, pb)["dist"] == want).all()
            for p in range(0, len(pa), 5):
                cost, i, j, ok = _apply(r["cigar"][r["cigar_off"][p]:r["cigar_off"][p + 1]], seqs[pa[p]], seqs[pb[p]])
                assert ok and i == len(seqs[pa[p]]) and j == len(seqs[pb[p]]) and cost == int(want[p])
            assert dctx.debug_violations()[0] == 0
        dctx.set_trace_budget(0)
        assert (dctx.edit_distance_pairs(cat, off, pa, pb, k=30)["dist"] == np.where(want <= 30, want, -1)).all()
        assert dctx.debug_violations()[0] == 0
    finally:
        dctx.close()


def test_randomised_differential(ctx):
    """Seeded random calls: lengths 0-3000 drawn log-uniformly, alphabets of 1-20 symbols, mutation rates 0-50 %, random k,
    distances always against the oracle's Myers, paths against the plain-matrix traceback when it fits."""
    import os
    n_seeds = int(os.environ.get("ISOF_FUZZ_SEEDS", "12"))
    for seed in range(n_seeds):
        rng = np.random.default_rng(1000 + seed)
        alpha = "ACGTNacgtnRYKMSW*-.x"[: int(rng.integers(1, 21))]
        seqs = []
        for _ in range(int(rng.integers(2, 60))):
            n = int(np.exp(rng.uniform(0, np.log(3000)))) - 1
            s = _rand(rng, n, alpha)
            seqs += [s, _mutate(rng, s, float(rng.random()) * 0.5, alpha)]
        cat, off = _pool(seqs)
        P = int(rng.integers(1, 400))
        pa = rng.integers(0, len(seqs), P).astype(np.int32)
        pb = np.where(rng.random(P) < 0.5, pa ^ 1, rng.integers(0, len(seqs), P)).astype(np.int32)
        want = oracle.edit_distance_pairs(cat, off, pa, pb, threads=8)
        k = int(rng.integers(0, 200)) if seed % 3 == 0 else -1
        r = ctx.edit_distance_pairs(cat, off, pa, pb, k=k, want_cigar=True)
        exp = want if k < 0 else np.where(want <= k, want, -1)
        assert (r["dist"] == exp).all(), seed
        # distances only: another kernel instantiation (up to eight words per lane instead of four)
        assert (ctx.edit_distance_pairs(cat, off, pa, pb, k=k)["dist"] == exp).all(), seed
        for p in range(P):
            runs = r["cigar"][r["cigar_off"][p]:r["cigar_off"][p + 1]]
            if exp[p] < 0:
                assert len(runs) == 0
                continue
            q, t = seqs[pa[p]], seqs[pb[p]]
            if len(q) * len(t) <= 400_000:
                assert runs.tolist() == oracle.edit_path(q, t)[1].tolist(), (seed, p)
            else:
                cost, i, j, ok = _apply(runs, q, t)
                assert ok and i == len(q) and j == len(t) and cost == int(want[p]), (seed, p)


def test_warp_walk_equals_thread_walk(ctx):
    """The path walk has two forms (one warp per pair with a ballot over the diagonal look-ahead, one thread per pair);
    isof_set_traceback_warp picks one or lets the pair count decide.  Same runs either way, also for multi-stripe pairs
    and under a threshold k."""
    rng = np.random.default_rng(44)
    seqs = []
    for n in [129, 130, 200, 255, 256, 400, 512, 513, 900, 1024, 1500, 2048, 3000, 4096, 4097, 7000]:
        s = _rand(rng, n)
        seqs += [s, _mutate(rng, s, 0.05), _mutate(rng, s, 0.35), s[: n // 2], _rand(rng, 60)]
    cat, off = _pool(seqs)
    S = len(seqs)
    pa = rng.integers(0, S, 500).astype(np.int32)
    pb = rng.integers(0, S, 500).astype(np.int32)
    nb = np.arange(S - 1, dtype=np.int32)
    pa = np.concatenate([pa, nb, nb + 1])
    pb = np.concatenate([pb, nb + 1, nb])
    out = {}
    try:
        for mode in (0, 1, 2):
            ctx.set_traceback_warp(mode)
            out[mode] = [ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=True),
                         ctx.edit_distance_pairs(cat, off, pa, pb, k=150, want_cigar=True)]
    finally:
        ctx.set_traceback_warp(2)
    for mode in (1, 2):
        for x in (0, 1):
            assert (out[mode][x]["dist"] == out[0][x]["dist"]).all()
            assert (out[mode][x]["cigar_off"] == out[0][x]["cigar_off"]).all()
            assert (out[mode][x]["cigar"] == out[0][x]["cigar"]).all()
    want = oracle.edit_distance_pairs(cat, off, pa, pb, threads=8)
    assert (out[1][0]["dist"] == want).all()
    for p in range(0, len(pa), 3):
        q, t = seqs[pa[p]], seqs[pb[p]]
        if len(q) * len(t) <= 2_000_000:
            runs = out[1][0]["cigar"][out[1][0]["cigar_off"][p]:out[1][0]["cigar_off"][p + 1]]
            assert runs.tolist() == oracle.edit_path(q, t)[1].tolist(), p


def test_extreme_shapes(ctx):
    """Corners of the layout: a 256-symbol pool with paths (one warp per block: 128 KB of match vectors), a multi-stripe
    query against a much longer target (line buffer of 20 k columns), empty sides in every class, a one-row query against
    a long target and the transpose."""
    rng = np.random.default_rng(2)
    allb = bytes(range(256))
    seqs = [allb, allb[::-1], bytes(rng.integers(0, 256, 700).astype(np.uint8)), bytes(rng.integers(0, 256, 150).astype(np.uint8))]
    cat, off = _pool(seqs)
    pa = np.array([0, 0, 1, 2, 3, 2], np.int32)
    pb = np.array([1, 0, 2, 3, 2, 2], np.int32)
    r = ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
    for p in range(len(pa)):
        d, runs = oracle.edit_path(seqs[pa[p]], seqs[pb[p]])
        assert int(r["dist"][p]) == d
        assert r["cigar"][r["cigar_off"][p]:r["cigar_off"][p + 1]].tolist() == runs.tolist()
    big = _rand(rng, 5000)
    long_t = _mutate(rng, big, 0.1) + _rand(rng, 15000)
    seqs = [big, long_t, "", "A", _rand(rng, 40), _rand(rng, 200), _rand(rng, 1000), _rand(rng, 4500)]
    cat, off = _pool(seqs)
    S = len(seqs)
    pa = np.repeat(np.arange(S, dtype=np.int32), S)
    pb = np.tile(np.arange(S, dtype=np.int32), S)
    want = oracle.edit_distance_pairs(cat, off, pa, pb, threads=8)
    assert (ctx.edit_distance_pairs(cat, off, pa, pb)["dist"] == want).all()
    for mode in (0, 1):
        ctx.set_traceback_warp(mode)
        try:
            r = ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
        finally:
            ctx.set_traceback_warp(2)
        assert (r["dist"] == want).all()
        for p in range(len(pa)):
            q, t = seqs[pa[p]], seqs[pb[p]]
            cost, i, j, ok = _apply(r["cigar"][r["cigar_off"][p]:r["cigar_off"][p + 1]], q, t)
            assert ok and i == len(q) and j == len(t) and cost == int(want[p]), (mode, p)


def test_release_scratch_returns_memory_and_calls_still_work():
    """isof_release_scratch gives the grow-only device scratch back (a traced call of 2 GB here) and the next calls of
    both aligners simply allocate again with identical results."""
    import torch
    from isonform_b200 import _lib, hotpath
    c = _lib.Context(0)
    try:
        rng = np.random.default_rng(3)
        seqs = [_rand(rng, 2000) for _ in range(64)]
        cat, off = _pool(seqs)
        pa = rng.integers(0, 64, 2000).astype(np.int32)
        pb = rng.integers(0, 64, 2000).astype(np.int32)
        a = c.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
        b1 = hotpath.align_pairs(seqs, pa[:100], pb[:100], 2, -2, 12, 1, c)
        free_before = torch.cuda.mem_get_info(0)[0]
        c.release_scratch()
        free_after = torch.cuda.mem_get_info(0)[0]
        assert free_after - free_before > (1 << 30)
        a2 = c.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
        b2 = hotpath.align_pairs(seqs, pa[:100], pb[:100], 2, -2, 12, 1, c)
        assert (a["dist"] == a2["dist"]).all() and (a["cigar"] == a2["cigar"]).all()
        assert (b1["score"] == b2["score"]).all() and (b1["cigar"] == b2["cigar"]).all()
    finally:
        c.close()


def test_banded_classes_under_a_bound_k(ctx):
    """With a bound k, long pairs of similar length run on a ring of lanes that slides down the band | i - j | <= k
    (ed_run_banded): exact whenever the distance is <= k, -1 otherwise.  Window widths 2 ... 32 words, words entering and
    leaving the ring, lengths around word borders, targets shorter and longer than the query, pairs far above the bound."""
    rng = np.random.default_rng(21)
    seqs = []
    for n in [161, 192, 200, 256, 500, 777, 1000, 1024, 1025, 2000, 3001, 4096, 5000, 9000]:
        s = _rand(rng, n)
        seqs += [s, _mutate(rng, s, 0.01), _mutate(rng, s, 0.03), _mutate(rng, s, 0.10), s[:-7], "ACG" + s, _rand(rng, n)]
    cat, off = _pool(seqs)
    S = len(seqs)
    base = np.arange(0, S, 7, dtype=np.int32)
    pa = np.concatenate([np.repeat(base, 7), np.repeat(base, 7) + np.tile(np.arange(7, dtype=np.int32), len(base))])
    pb = np.concatenate([np.repeat(base, 7) + np.tile(np.arange(7, dtype=np.int32), len(base)), np.repeat(base, 7)])
    full = oracle.edit_distance_pairs(cat, off, pa, pb, threads=8)
    assert (ctx.edit_distance_pairs(cat, off, pa, pb)["dist"] == full).all()
    for k in (0, 1, 3, 15, 16, 17, 31, 47, 64, 100, 200, 480, 495, 496, 600, 5000):
        got = ctx.edit_distance_pairs(cat, off, pa, pb, k=k)["dist"]
        want = np.where(full <= k, full, -1)
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (k, [(int(p), len(seqs[pa[p]]), len(seqs[pb[p]]), int(got[p]), int(want[p])) for p in bad[:8]])


def test_device_pointer_variant_with_an_unaligned_run_buffer(ctx):
    """isof_edit_distance_pairs_dev on the caller's device buffers; the run buffer starts 4 bytes off a 16-byte boundary, so
    the runs kernel must fall back from its 16-byte stores."""
    import torch
    rng = np.random.default_rng(91)
    seqs = []
    for _ in range(40):
        s = _rand(rng, int(rng.integers(1, 900)))
        seqs += [s, _mutate(rng, s, 0.2)]
    cat, off = _pool(seqs)
    pa = rng.integers(0, len(seqs), 600).astype(np.int32)
    pb = rng.integers(0, len(seqs), 600).astype(np.int32)
    ref = ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=True)
    dev = torch.device("cuda", 0)
    d_cat = torch.from_numpy(np.array(cat)).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    d_pa, d_pb = torch.from_numpy(pa).to(dev), torch.from_numpy(pb).to(dev)
    d_dist = torch.empty(len(pa), dtype=torch.int32, device=dev)
    d_coff = torch.empty(len(pa) + 1, dtype=torch.int64, device=dev)
    cap = len(ref["cigar"]) + 8
    for shift in (0, 1, 3):
        d_buf = torch.zeros(cap + 4, dtype=torch.int32, device=dev)
        d_cig = d_buf[shift:]
        torch.cuda.synchronize()
        total = ctx.edit_distance_pairs_dev(d_cat.data_ptr(), d_off.data_ptr(), len(seqs), len(cat), d_pa.data_ptr(), d_pb.data_ptr(),
                                            len(pa), -1, d_dist.data_ptr(), d_cig.data_ptr(), cap, d_coff.data_ptr())
        torch.cuda.synchronize()
        assert total == len(ref["cigar"])
        assert (d_dist.cpu().numpy() == ref["dist"]).all()
        assert (d_coff.cpu().numpy() == ref["cigar_off"]).all()
        assert (d_cig[:total].cpu().numpy().view(np.uint32) == ref["cigar"]).all()
        assert int(d_buf[:shift].abs().sum()) == 0 and int(d_cig[total:].abs().sum()) == 0      # nothing written outside
