"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs, and against the committed golden fixtures."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from conftest import lcg, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from isonform_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def _rand_dna(rng, n, alphabet="ACGT"):
    return "".join(alphabet[int(x)] for x in rng.integers(0, len(alphabet), size=n))


# ------------------------------------------------------------------------------- minimizers
def test_minimizers_golden(ctx):
    from isonform_b200 import hotpath
    g = load_golden("minimizers.json")
    by_kw = {}
    for c in g["cases"]:
        by_kw.setdefault((c["k"], c["w"]), []).append(c)
    for (k, w), cases in by_kw.items():
        got = hotpath.minimizer_positions_batch([c["seq"] for c in cases], k, w, ctx)
        for c, pos in zip(cases, got):
            assert pos.tolist() == c["pos"], c["name"]


def test_minimizers_kat3_digest(ctx):
    from isonform_b200 import hotpath
    g = load_golden("minimizers.json")["kat3"]
    pos = hotpath.minimizer_positions_batch([lcg(g["n"], g["seed"])], g["k"], g["w"], ctx)[0].tolist()
    assert len(pos) == g["count"]
    assert hashlib.sha256(",".join(map(str, pos)).encode()).hexdigest()[:16] == g["sha256_16"]


def test_minimizers_vs_oracle_ragged(ctx):
    from isonform_b200 import hotpath
    rng = np.random.default_rng(17)
    seqs = []
    for i in range(300):
        n = int(rng.integers(0, 1500)) if i % 7 else int(rng.integers(0, 60))
        alpha = ["ACGT", "AC", "ACGTN", "A"][i % 4] if i % 5 == 0 else "ACGT"
        seqs.append(_rand_dna(rng, n, alpha))
    seqs += ["", "A", "ACGT" * 300, ("ACGTTGCA" * 40)[:255], ("ACGTTGCA" * 40)[:256], ("ACGTTGCA" * 40)[:257]]
    for (k, w) in [(20, 31), (9, 20), (13, 13), (31, 100)]:
        got = hotpath.minimizer_positions_batch(seqs, k, w, ctx)
        for s, pos in zip(seqs, got):
            assert pos.tolist() == oracle.minimizers(s, k, w).tolist(), (k, w, len(s))


def test_minimizers_dict_interface(ctx):
    from isonform_b200 import hotpath
    s = lcg(200, 42)
    reads = {1: ("a", s, "I" * 200), 2: ("b", s[:90] + s[95:], "I" * 195)}
    M = hotpath.get_minimizers_and_positions(reads, 31, 20, "lex", ctx)
    assert [p for _, p in M[1]] == [9, 13, 16, 19, 29, 40, 43, 54, 60, 68, 75, 76, 86, 94, 100, 108, 118, 121, 131,
                                    139, 142, 152, 157, 158, 170, 171]
    assert all(kmer == reads[r][1][p:p + 20] for r in M for kmer, p in M[r])
    assert M == oracle.get_minimizers_and_positions(reads, 31, 20, "lex")
    assert hotpath.get_kmer_minimizers(s, 9, 20, ctx) == oracle.get_kmer_minimizers(s, 9, 20)


def test_minimizers_empty_batch(ctx):
    pos, off = ctx.minimizers_batch(np.zeros(0, np.uint8), np.zeros(1, np.int64), 20, 31)
    assert len(pos) == 0 and off.tolist() == [0]


# ------------------------------------------------------------------------------- alignment
def _check_pairs(ctx, seqs, pa, pb, params, tiebreak=None):
    from isonform_b200 import hotpath
    res = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx)
    for p in range(len(pa)):
        sc, eq, er, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], *params, tiebreak=tiebreak)
        got = res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]]
        assert (int(res["score"][p]), int(res["end_q"][p]), int(res["end_r"][p])) == (sc, eq, er), (p, len(seqs[pa[p]]), len(seqs[pb[p]]))
        assert got.tolist() == runs.tolist(), (p, oracle.cigar_runs_to_string(got), oracle.cigar_runs_to_string(runs))


def test_align_golden_decisions(ctx):
    from isonform_b200 import hotpath
    g = load_golden("decisions.json")
    for rec in g["pairs"]:
        a1, a2, cig, tup, score = hotpath.parasail_alignment(rec["s1"], rec["s2"], 2, -8, 12, 1, ctx)
        assert (cig, score) == (rec["bubble"]["cigar"], rec["bubble"]["score"]), rec["name"]
        assert hashlib.sha256(a1.encode()).hexdigest()[:12] == rec["bubble"]["a1_sha"]
        assert hashlib.sha256(a2.encode()).hexdigest()[:12] == rec["bubble"]["a2_sha"]
        assert hotpath.parse_cigar_diversity(tup, 0.20, 5) == rec["bubble"]["pop_d5"]
        a1, a2, cig, tup, score = hotpath.parasail_alignment(rec["s1"], rec["s2"], 2, -2, 12, 1, ctx)
        assert (cig, score) == (rec["iso"]["cigar"], rec["iso"]["score"]), rec["name"]
        assert hotpath.find_first_significant_match(a1, a2, 30, 0.7) == rec["iso"]["start_match"]
        assert hotpath.find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7) == rec["iso"]["end_match"]
        for dl, d in ((5, 0.15), (10, 0.15), (5, 0.1)):
            assert hotpath.align_to_merge(rec["s1"], rec["s2"], d, dl, 30, 50, ctx) == rec["iso"]["merge_dl%d_d%s" % (dl, d)], rec["name"]


def test_align_golden_gates(ctx):
    from isonform_b200 import hotpath
    g = load_golden("decisions.json")
    for gt in g["gates"]:
        pop, aligned = hotpath.bubble_decisions([gt["s1"], gt["s2"]], [0], [1], gt["delta_len"], ctx)
        assert (bool(pop[0]), bool(aligned[0])) == (gt["pop"], gt["aligned"])


def test_align_random_shapes_vs_oracle(ctx):
    rng = np.random.default_rng(23)
    seqs, pa, pb = [], [], []
    shapes = [(1, 1), (1, 40), (40, 1), (7, 9), (8, 8), (31, 33), (255, 256), (256, 255), (257, 300), (300, 257),
              (511, 513), (513, 40), (40, 513), (1000, 1100), (64, 2000), (2000, 64)]
    for (n, m) in shapes:
        for rep in range(3):
            x = _rand_dna(rng, max(n, m))
            if rep == 0:
                a, b = _rand_dna(rng, n), _rand_dna(rng, m)
            else:
                from isonform_b200 import synth
                y = synth.mutate(rng, x, 0.05 * rep)
                a, b = x[:n], (y + _rand_dna(rng, m))[:m]
            seqs += [a, b]
            pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    for params in [(2, -8, 12, 1), (2, -2, 12, 1), (1, -1, 2, 1), (3, -4, 5, 2), (2, -2, 1, 1)]:
        _check_pairs(ctx, seqs, pa, pb, params)


def test_align_tie_heavy_vs_oracle(ctx):
    """Low-complexity sequences and small gap costs maximise ties in every max of the recurrence."""
    rng = np.random.default_rng(29)
    seqs, pa, pb = [], [], []
    for i in range(150):
        alpha = ["AC", "A", "ACG", "AAC"][i % 4]
        seqs += [_rand_dna(rng, int(rng.integers(1, 70)), alpha), _rand_dna(rng, int(rng.integers(1, 70)), alpha)]
        pa.append(2 * i); pb.append(2 * i + 1)
    for params in [(2, -2, 2, 1), (1, -1, 1, 1), (2, -2, 12, 1), (2, -8, 12, 1), (1, -3, 2, 2), (2, -2, 0, 0)]:
        _check_pairs(ctx, seqs, pa, pb, params)


def test_align_wildcards_and_case(ctx):
    rng = np.random.default_rng(31)
    seqs, pa, pb = [], [], []
    for i in range(40):
        a = _rand_dna(rng, int(rng.integers(5, 300)), "ACGTN")
        b = _rand_dna(rng, int(rng.integers(5, 300)), "ACGTNacgtX")
        seqs += [a, b]; pa.append(2 * i); pb.append(2 * i + 1)
    seqs += ["X" * 40, lcg(45, 42)]; pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    _check_pairs(ctx, seqs, pa, pb, (2, -8, 12, 1))
    _check_pairs(ctx, seqs, pa, pb, (2, -2, 12, 1))


def test_align_all_pairs_pool_and_chunking(ctx):
    """An all-pairs job over a small pool with a tiny trace budget: many chunks, same answers."""
    from isonform_b200 import synth
    rng = np.random.default_rng(37)
    base = _rand_dna(rng, 600)
    seqs = [synth.mutate(rng, base, 0.04) for _ in range(14)]
    ia, ib = np.triu_indices(len(seqs), 1)
    ctx.set_trace_budget(2 << 20)
    try:
        _check_pairs(ctx, seqs, ia.tolist(), ib.tolist(), (2, -2, 12, 1))
    finally:
        ctx.set_trace_budget(0)


def test_align_features_vs_oracle(ctx):
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(41)
    seqs, pa, pb = [], [], []
    for i in range(60):
        ln = int(rng.integers(20, 900))
        x = _rand_dna(rng, ln)
        y = synth.mutate(rng, x, [0.0, 0.01, 0.05, 0.12][i % 4])
        if i % 3 == 0 and ln > 100:
            c = ln // 2; y = y[:c] + y[c + int(rng.integers(3, 30)):]
        if i % 4 == 1:
            y = _rand_dna(rng, int(rng.integers(0, 70))) + y
        if i % 5 == 2:
            x = x + _rand_dna(rng, int(rng.integers(0, 70)))
        seqs += [x, y or "A"]; pa.append(2 * i); pb.append(2 * i + 1)
    for dl in (5, 10):
        res = hotpath.align_pairs(seqs, pa, pb, 2, -2, 12, 1, ctx, want_cigar=True, features_delta_len=dl)
        merge = hotpath._merge_from_features(res["features"], 0.15, dl, 30, 50)
        for p in range(len(pa)):
            s1, s2 = seqs[pa[p]], seqs[pb[p]]
            a1, a2, cig, tup, score = oracle.parasail_alignment(s1, s2, 2, -2, 12, 1)
            f = res["features"][p]
            assert oracle.cigar_runs_to_string(res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]]) == cig
            sm = oracle.find_first_significant_match(a1, a2, 30, 0.7)
            em = oracle.find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7)
            alen = sum(l for l, _ in tup)
            assert (int(f[0]), int(f[1]), int(f[4]), int(f[5])) == (alen, len(tup), sm, em)
            assert int(f[2]) == sum(l for l, o in tup if o != "=") and int(f[3]) == max([l for l, o in tup if o != "="] + [0])
            if sm >= 0 and em >= 0:
                brk, miss_runs, bm, bn, am, an, _ = oracle.isoform_features(tup, dl, sm, alen - em)
                assert [int(v) for v in f[6:12]] == [brk, miss_runs, bm, bn, am, an]
            assert bool(merge[p]) == oracle.align_to_merge(s1, s2, 0.15, dl, 30, 50)
        pop = hotpath._pop_from_features(res["features"], 0.20, dl)
        for p in range(len(pa)):
            tup = oracle.cigar_runs_to_tuples(res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]])
            assert bool(pop[p]) == oracle.parse_cigar_diversity(tup, 0.20, dl)


def _tiebreak_cases():
    """Pairs that exercise every tie of the recurrence and every kernel path: tie-heavy low-complexity strings,
    random shapes around the stripe / lane boundaries, related reads (packed path, shared references), a wildcard
    pair (32-bit WILD path), one pair beyond the 16-bit range."""
    from isonform_b200 import synth
    rng = np.random.default_rng(97)
    seqs, pa, pb = [], [], []

    def add(a, b):
        seqs.extend([a, b]); pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    for i in range(60):
        alpha = ["AC", "A", "ACG", "AAC"][i % 4]
        add(_rand_dna(rng, int(rng.integers(1, 70)), alpha), _rand_dna(rng, int(rng.integers(1, 70)), alpha))
    for (n, m) in [(1, 1), (1, 40), (40, 1), (8, 8), (31, 33), (255, 256), (257, 300), (300, 257), (513, 40), (40, 513),
                   (700, 760)]:
        x = _rand_dna(rng, max(n, m))
        y = synth.mutate(rng, x, 0.06)
        add(x[:n], (y + _rand_dna(rng, m))[:m])
        add(_rand_dna(rng, n, "ACGGT"), _rand_dna(rng, m, "ACGGT"))
    base = _rand_dna(rng, 420)
    pool0 = len(seqs)
    seqs.extend(synth.mutate(rng, base, 0.05) for _ in range(7))           # all pairs of related reads: shared references
    for i in range(7):
        for j in range(i + 1, 7):
            pa.append(pool0 + i); pb.append(pool0 + j)
    add(_rand_dna(rng, 150, "ACGTN"), _rand_dna(rng, 170, "ACGTNacgtX"))
    add("X" * 40, lcg(45, 42))
    return seqs, pa, pb


@pytest.mark.parametrize("tb_index", range(16))
def test_align_all_tiebreak_tables_vs_oracle(ctx, tb_index):
    """The tie-break table (H-source order, open vs extend, end-cell order, I/D letters) is a switch of the CUDA
    path, not a constant: under each of the 16 tables the GPU result equals the oracle's, on both arithmetic paths,
    including the decision features and the reference's decisions evaluated from them."""
    from isonform_b200 import hotpath
    tb = oracle.all_tiebreaks()[tb_index]
    seqs, pa, pb = _tiebreak_cases()
    ctx.set_tiebreak(tb.h_e_before_f, tb.gap_open_on_tie, tb.end_col_first, tb.swap_id)
    try:
        for params in [(2, -8, 12, 1), (2, -2, 12, 1), (1, -1, 1, 1), (2, -2, 2, 1), (2, -2, 0, 0)]:
            for packed in (True, False):
                ctx.set_packed(packed)
                _check_pairs(ctx, seqs, pa, pb, params, tiebreak=tb)
        ctx.set_packed(True)
        # decision features under the table (IsoformGeneration.py:251-397, SimplifyGraph.py:175-196)
        for dl in (5, 10):
            res = hotpath.align_pairs(seqs, pa, pb, 2, -2, 12, 1, ctx, want_cigar=True, features_delta_len=dl)
            merge = hotpath._merge_from_features(res["features"], 0.15, dl, 30, 50)
            pop = hotpath._pop_from_features(res["features"], 0.20, dl)
            for p in range(len(pa)):
                s1, s2 = seqs[pa[p]], seqs[pb[p]]
                a1, a2, cig, tup, score = oracle.parasail_alignment(s1, s2, 2, -2, 12, 1, tb)
                f = res["features"][p]
                sm = oracle.find_first_significant_match(a1, a2, 30, 0.7)
                em = oracle.find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7)
                assert (int(f[4]), int(f[5])) == (sm, em), (tb_index, p, cig)
                assert bool(merge[p]) == oracle.align_to_merge(s1, s2, 0.15, dl, 30, 50, tiebreak=tb), (tb_index, p)
                assert bool(pop[p]) == oracle.parse_cigar_diversity(tup, 0.20, dl)
    finally:
        ctx.set_packed(True)
        ctx.set_tiebreak(0, 0, 0, 0)


def test_align_batch_form_and_errors(ctx):
    from isonform_b200 import _lib, hotpath
    s = lcg(120, 42)
    res = ctx.sg_align_batch([s, s[30:90], s], [s, s, s[30:90]])
    cigs = [oracle.cigar_runs_to_string(res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]]) for p in range(3)]
    assert cigs == ["120=", "30D60=30D", "30I60=30I"] and res["score"].tolist() == [240, 120, 120]
    with pytest.raises(ValueError):
        hotpath.parasail_alignment("", "ACGT", ctx=ctx)
    with pytest.raises(_lib.IsofError):
        hotpath.align_pairs(["ACGT", ""], [0], [1], ctx=ctx)
    with pytest.raises(_lib.IsofError):
        hotpath.align_pairs(["ACGT", "AC"], [0], [5], ctx=ctx)


def test_align_score_only_and_capacity_retry(ctx):
    from isonform_b200 import hotpath
    rng = np.random.default_rng(43)
    seqs = [_rand_dna(rng, 50) for _ in range(8)]
    ia, ib = np.triu_indices(8, 1)
    full = hotpath.align_pairs(seqs, ia, ib, ctx=ctx)
    score_only = hotpath.align_pairs(seqs, ia, ib, ctx=ctx, want_cigar=False)
    assert score_only["cigar"] is None and score_only["score"].tolist() == full["score"].tolist()
    assert score_only["cigar_off"].tolist() == full["cigar_off"].tolist()
    cat, off = hotpath.pack_sequences(seqs)
    small = ctx.sg_align_pairs(cat, off, ia, ib, cigar_cap=3)      # forces ISOF_ERR_CAPACITY + retry
    assert small["cigar"].tolist() == full["cigar"].tolist()


# ------------------------------------------------------------------------------- pair database / spans
def _span_case_check(ctx, reads, k, w, xmin, xmax, dl, expect=None):
    from array import array
    from isonform_b200 import hotpath, spans
    M = hotpath.get_minimizers_and_positions(reads, w, k, "lex", ctx)
    assert M == oracle.get_minimizers_and_positions(reads, w, k, "lex")
    idx = spans.SpanIndex(reads, M, k, xmin, xmax, dl, ctx)
    odb = oracle.get_minimizer_combinations_database(M, k, xmin, xmax, reads)
    want_db = [[m1, m2, list(odb[m1][m2])] for m1 in odb for m2 in odb[m1] if len(odb[m1][m2])]
    gdb = idx.database()
    got_db = [[m1, m2, list(gdb[m1][m2])] for m1 in gdb for m2 in gdb[m1]]
    assert got_db == want_db
    if expect is not None:
        assert got_db == expect["db"]
    for r_id in sorted(reads):
        want = oracle.pyside.read_intervals(r_id, M[r_id], odb, k, xmin, xmax, dl)
        got = idx.intervals(r_id)
        assert [(a, b, c, list(d)) for a, b, c, d in got] == [(a, b, c, list(d)) for a, b, c, d in want], r_id
        # the reference-signature seam, anchor by anchor
        seam = []
        for (m1, p1), sp in spans.minimizers_comb_iterator(M[r_id], k, xmin, xmax):
            if sp:
                idx.find_most_supported_span(r_id, m1, p1, sp, None, seam, k, dl)
        assert [(a, b, c, list(d)) for a, b, c, d in seam] == [(a, b, c, list(d)) for a, b, c, d in want]
        if expect is not None:
            e = expect["per_read"][str(r_id)]
            assert [[a, b, c, hashlib.sha256(array("I", d).tobytes()).hexdigest()[:12]] for a, b, c, d in got] == e["intervals"]
    graph_iv, new_reads, skipped = spans.batch_intervals_for_graph(idx, reads)
    gid = 1
    for r_id in sorted(reads):
        want = oracle.pyside.read_intervals(r_id, M[r_id], odb, k, xmin, xmax, dl)
        picks = None
        if want:
            srt = sorted(want, key=lambda x: x[1])
            picks = [srt[j] for j in oracle.solve_WIS(srt)[::-1]]
        if picks:
            assert [(a, b, c, list(d)) for a, b, c, d in graph_iv[gid]] == [(a, b, c, list(d)) for a, b, c, d in picks]
            assert new_reads[gid] == reads[r_id]
            if expect is not None:
                assert [[a, b, c, list(d)] for a, b, c, d in graph_iv[gid]] == expect["per_read"][str(r_id)]["wis"]
            gid += 1
        else:
            assert r_id in skipped
    assert len(graph_iv) == gid - 1


def test_spans_golden(ctx):
    for case in load_golden("spans.json"):
        reads = {int(r): ("acc", s, "") for r, s in case["reads"].items()}
        _span_case_check(ctx, reads, case["k"], case["w"], case["xmin"], case["xmax"], case["delta_len"], expect=case)


def test_spans_random_cluster_vs_oracle(ctx):
    from isonform_b200 import polya, synth
    rd, _iso, _t = synth.make_cluster(seed=21, gene_len=900, n_isoforms=4, n_reads=60, eps=0.03)
    reads = {i + 1: (acc, polya.remove_read_polyA_ends(s, 12, 1), "") for i, (acc, s) in enumerate(rd)}
    reads[61] = ("short", "ACGTACGTAC", "")                  # shorter than one window
    reads[62] = ("withN", rd[0][1][:200] + "NNNN" + rd[0][1][200:400], "")
    reads[63] = ("lower", rd[1][1][:300].lower(), "")
    reads[64] = ("empty", "", "")
    _span_case_check(ctx, reads, 20, 31, 40, 80, 5)
    _span_case_check(ctx, reads, 9, 20, 18, 60, 10)


# ------------------------------------------------------------------------------- packed vs 32-bit DP paths
def test_packed_and_32bit_paths_agree_and_match_oracle(ctx):
    """Tasks pair up consecutive alignments; mix sizes, wildcards (forces the 32-bit path for a task),
    odd counts and over-long sequences so every dispatch branch of sg_dp_kernel runs."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(53)
    seqs, pa, pb = [], [], []
    for i in range(61):                                   # odd number of pairs
        n = int(rng.integers(1, 700))
        x = _rand_dna(rng, n)
        y = synth.mutate(rng, x, 0.06) if i % 3 else _rand_dna(rng, int(rng.integers(1, 700)))
        if i % 9 == 4:
            y = y[:len(y) // 2] + "N" + y[len(y) // 2:]    # wildcard -> 32-bit WILD path for this task
        seqs += [x, y or "C"]; pa.append(2 * i); pb.append(2 * i + 1)
    big = _rand_dna(rng, 4300)                             # min(n,m) > 4093: not eligible for 16 bit at match=2
    seqs += [big, synth.mutate(rng, big, 0.03)]; pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    for params in [(2, -2, 12, 1), (2, -8, 12, 1), (5, -4, 10, 2), (1, -1, 0, 0)]:
        ctx.set_packed(True)
        a = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)
        ctx.set_packed(False)
        try:
            b = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)
        finally:
            ctx.set_packed(True)
        for key in ("score", "cigar_off", "cigar", "features"):
            assert np.array_equal(a[key], b[key]), (params, key)
        for p in range(0, len(pa), 5):
            sc, eq, er, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], *params)
            assert int(a["score"][p]) == sc
            assert a["cigar"][a["cigar_off"][p]:a["cigar_off"][p + 1]].tolist() == runs.tolist(), (params, p)


def test_packed_path_range_limits(ctx):
    """Sequences right at the 16-bit eligibility limits (4*match*min(n,m)+16 <= 32767) and scores that
    use the whole range: identical sequences maximise the score, unrelated ones the negative side."""
    from isonform_b200 import hotpath
    rng = np.random.default_rng(59)
    a = _rand_dna(rng, 4093)
    b = _rand_dna(rng, 4093)
    seqs = [a, a, a, b, a[:4000], a[50:], "A" * 4093, "C" * 4093]
    pa, pb = [0, 2, 4, 6], [1, 3, 5, 7]
    res = hotpath.align_pairs(seqs, pa, pb, 2, -2, 12, 1, ctx=ctx)
    for p in range(4):
        sc, eq, er, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], 2, -2, 12, 1)
        assert (int(res["score"][p]), int(res["end_q"][p]), int(res["end_r"][p])) == (sc, eq, er)
        assert res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]].tolist() == runs.tolist()
    assert int(res["score"][0]) == 2 * 4093


def test_pipelined_and_serial_chunking_agree(ctx):
    """Many small chunks: the two-stream pipeline (default) and the serial path give identical output,
    including the CIGAR offsets that are chained across chunks."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(61)
    base = _rand_dna(rng, 500)
    seqs = [synth.mutate(rng, base, 0.05) for _ in range(24)]
    ia, ib = np.triu_indices(len(seqs), 1)
    ctx.set_trace_budget(3 << 20)
    try:
        ctx.set_overlap(True)
        a = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
        ctx.set_overlap(False)
        b = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
    finally:
        ctx.set_overlap(True)
        ctx.set_trace_budget(0)
    for key in ("score", "end_q", "end_r", "cigar_off", "cigar", "features"):
        if a[key] is not None:
            assert np.array_equal(a[key], b[key]), key
    for p in range(0, len(ia), 11):
        sc, eq, er, runs = oracle.sg_align(seqs[ia[p]], seqs[ib[p]], 2, -2, 12, 1)
        assert a["cigar"][a["cigar_off"][p]:a["cigar_off"][p + 1]].tolist() == runs.tolist()


# ------------------------------------------------------------------------------- full-size properties
def _rescore(s1, s2, runs, match, mismatch, gap_open, gap_ext):
    """(score with interior gaps charged and end gaps free, query consumed, ref consumed)."""
    ops = [(int(r) >> 4, int(r) & 15) for r in runs]
    i = j = sc = 0
    for k_, (ln, op) in enumerate(ops):
        edge = k_ == 0 or k_ == len(ops) - 1
        if op in (7, 8):
            a = np.frombuffer(s1[i:i + ln].encode(), np.uint8)
            b = np.frombuffer(s2[j:j + ln].encode(), np.uint8)
            assert bool((a == b).all()) == (op == 7) or ln == 0
            sc += int((a == b).sum()) * match + int((a != b).sum()) * mismatch
            i += ln; j += ln
        elif op == 1:
            i += ln; sc -= 0 if edge else gap_open + (ln - 1) * gap_ext
        else:
            j += ln; sc -= 0 if edge else gap_open + (ln - 1) * gap_ext
    return sc, i, j


def test_c2_size_all_pairs_properties_and_oracle_sample(ctx):
    """BASELINE config-2 read lengths (2.7 kb, 5 % error), all pairs of 110 reads (5 995 alignments, 4.5e10
    cells): size-independent properties on every pair, bit-exact oracle comparison on a sample, and the
    packed and 32-bit paths against each other on a slice."""
    from isonform_b200 import hotpath, synth, polya
    reads, _iso, _t = synth.make_config("C2", n_reads=110)
    seqs = sorted((polya.remove_read_polyA_ends(s, 12, 1) for _, s in reads), key=len)
    ia, ib = np.triu_indices(len(seqs), 1)
    res = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
    off = res["cigar_off"]
    assert off[0] == 0 and (np.diff(off) > 0).all() and off[-1] == len(res["cigar"])
    lens = np.array([len(s) for s in seqs])
    runs_len = res["cigar"] >> 4
    ops = res["cigar"] & 15
    assert set(np.unique(ops).tolist()) <= {1, 2, 7, 8}
    # per pair: columns consumed == sequence lengths (vectorised over all 5 995 cigars)
    seg = np.repeat(np.arange(len(ia)), np.diff(off))
    q_used = np.bincount(seg, weights=runs_len * np.isin(ops, (1, 7, 8)), minlength=len(ia)).astype(np.int64)
    r_used = np.bincount(seg, weights=runs_len * np.isin(ops, (2, 7, 8)), minlength=len(ia)).astype(np.int64)
    assert np.array_equal(q_used, lens[ia]) and np.array_equal(r_used, lens[ib])
    assert np.array_equal(res["features"][:, 0], np.bincount(seg, weights=runs_len, minlength=len(ia)).astype(np.int64))
    assert np.array_equal(res["features"][:, 1], np.diff(off))
    # no two adjacent runs share an op (run-length encoding is canonical)
    same = (ops[1:] == ops[:-1]) & (seg[1:] == seg[:-1])
    assert not same.any()
    # scores: re-scoring the cigar never undercuts the reported score; identical reads score 2*len
    rng = np.random.default_rng(7)
    sample = rng.choice(len(ia), size=36, replace=False)
    for p in sample:
        runs = res["cigar"][off[p]:off[p + 1]]
        sc, qi, rj = _rescore(seqs[ia[p]], seqs[ib[p]], runs, 2, -2, 12, 1)
        assert sc >= int(res["score"][p])
        osc, oeq, oer, oruns = oracle.sg_align(seqs[ia[p]], seqs[ib[p]], 2, -2, 12, 1)
        assert int(res["score"][p]) == osc
        assert runs.tolist() == oruns.tolist()
    ends = hotpath.align_pairs(seqs, ia[sample], ib[sample], 2, -2, 12, 1, ctx=ctx, want_cigar=False)
    for q, p in enumerate(sample):
        osc, oeq, oer, _ = oracle.sg_align(seqs[ia[p]], seqs[ib[p]], 2, -2, 12, 1)
        assert (int(ends["score"][q]), int(ends["end_q"][q]), int(ends["end_r"][q])) == (osc, oeq, oer)
    merge = hotpath._merge_from_features(res["features"], 0.15, 5, 30, 50)
    for p in sample[:12]:
        assert bool(merge[p]) == oracle.align_to_merge(seqs[ia[p]], seqs[ib[p]], 0.15, 5, 30, 50)
    # packed vs 32-bit on a slice
    sl = slice(0, 400)
    ctx.set_packed(False)
    try:
        ref32 = hotpath.align_pairs(seqs, ia[sl], ib[sl], 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
    finally:
        ctx.set_packed(True)
    assert np.array_equal(ref32["score"], res["score"][sl]) and np.array_equal(ref32["features"], res["features"][sl])
    assert np.array_equal(ref32["cigar"], res["cigar"][:off[400]])
    # self alignment is the identity
    self_res = hotpath.align_pairs(seqs[:8], np.arange(8), np.arange(8), 2, -2, 12, 1, ctx=ctx)
    assert self_res["score"].tolist() == [2 * len(s) for s in seqs[:8]]
    assert [int(r) for r in self_res["cigar"]] == [(len(s) << 4) | 7 for s in seqs[:8]]


def test_c2_size_minimizers_and_spans_properties(ctx):
    """Full C2 batch (1000 reads): minimizer/pair invariants that hold at any size, plus an oracle check on
    a sample of reads."""
    from isonform_b200 import hotpath, polya, spans, synth
    reads_l, _iso, _t = synth.make_config("C2")
    reads = {i + 1: (acc, polya.remove_read_polyA_ends(s, 12, 1), "") for i, (acc, s) in enumerate(reads_l)}
    k, w = 20, 31
    M = hotpath.get_minimizers_and_positions(reads, w, k, "lex", ctx)
    for r in (1, 17, 333, 1000):
        assert M[r] == oracle.get_kmer_minimizers(reads[r][1], k, w)
    for r in list(reads)[::37]:
        pos = np.array([p for _, p in M[r]])
        assert (np.diff(pos) > 0).all() and pos[0] <= w - k and pos[-1] < len(reads[r][1]) - k
        assert (np.diff(pos) <= w - k + 1).all()          # every window holds a minimizer
    idx = spans.SpanIndex(reads, M, k, 40, 80, 5, ctx)
    d = idx.rec_p2 - idx.rec_p1
    assert ((d > 40) & (d <= 80)).all()
    assert (np.diff(idx.rec_read) >= 0).all()             # emission order: read by read
    sizes = np.diff(idx.bucket_off)
    assert sizes.sum() == idx.n_rec and (sizes > 0).all()
    assert np.array_equal(np.sort(idx.sorted_rec), np.arange(idx.n_rec))
    bsz = sizes[idx.rec_bucket]
    assert ((idx.rec_count > 0) == ((bsz >= 3) & (idx.bucket_dropped[idx.rec_bucket] == 0))).all()
    assert (idx.rec_count <= bsz).all()
    # oracle comparison for three reads (the reference loop over the same database)
    odb = oracle.get_minimizer_combinations_database(M, k, 40, 80, reads)
    for r in (2, 500, 999):
        want = oracle.pyside.read_intervals(r, M[r], odb, k, 40, 80, 5)
        got = idx.intervals(r)
        assert [(a, b, c, list(x)) for a, b, c, x in got] == [(a, b, c, list(x)) for a, b, c, x in want]


def test_c4_size_long_reads_vs_oracle(ctx):
    """BASELINE config-4 read lengths (10 kb, 5 % error): beyond the 16-bit range, so the 32-bit path with
    40-stripe traces; all pairs of 6 reads bit-exact against the oracle, plus minimizers/spans of the reads."""
    from isonform_b200 import hotpath, polya, spans, synth
    reads_l, _iso, _t = synth.make_config("C4", n_reads=6)
    seqs = [polya.remove_read_polyA_ends(s, 12, 1) for _, s in reads_l]
    assert min(len(s) for s in seqs) > 6000
    ia, ib = np.triu_indices(len(seqs), 1)
    res = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
    cat = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    osc, _oeq, _oer, ocig, ooff = oracle.sg_align_pairs(cat, off, ia.astype(np.int32), ib.astype(np.int32), 2, -2, 12, 1, threads=0)
    assert res["score"].tolist() == osc.tolist()
    assert res["cigar_off"].tolist() == ooff.tolist() and np.array_equal(res["cigar"], ocig)
    reads = {i + 1: ("r%d" % i, s, "") for i, s in enumerate(seqs)}
    M = hotpath.get_minimizers_and_positions(reads, 31, 20, "lex", ctx)
    assert M == oracle.get_minimizers_and_positions(reads, 31, 20, "lex")
    idx = spans.SpanIndex(reads, M, 20, 40, 80, 5, ctx)
    odb = oracle.get_minimizer_combinations_database(M, 20, 40, 80, reads)
    for r in reads:
        want = oracle.pyside.read_intervals(r, M[r], odb, 20, 40, 80, 5)
        assert [(a, b, c, list(x)) for a, b, c, x in idx.intervals(r)] == [(a, b, c, list(x)) for a, b, c, x in want]


def test_super_batched_front_half_on_gpu(ctx):
    """isof_spans_multi: several clusters in one launch sequence == cluster by cluster == oracle."""
    from isonform_b200 import runner, synth
    batches = []
    for seed, n in ((51, 40), (52, 25), (53, 33), (54, 1)):
        rd, _iso, _t = synth.make_cluster(seed=seed, gene_len=800, n_isoforms=3, n_reads=n, eps=0.02)
        batches.append(runner.read_batch([(acc, s, "") for acc, s in rd]))
    multi = runner.batches_front_half(batches, 20, 31, 40, 80, 5, ctx)
    for reads, (giv, new_reads, skipped, idx) in zip(batches, multi):
        giv1, new1, skip1, idx1 = runner.batch_front_half(reads, 20, 31, 40, 80, 5, ctx)
        assert skipped == skip1 and new_reads == new1
        assert {g: [(a, b, c, list(d)) for a, b, c, d in v] for g, v in giv.items()} == \
               {g: [(a, b, c, list(d)) for a, b, c, d in v] for g, v in giv1.items()}
        M = oracle.get_minimizers_and_positions(reads, 31, 20)
        odb = oracle.get_minimizer_combinations_database(M, 20, 40, 80, reads)
        for r in list(reads)[::5]:
            want = oracle.pyside.read_intervals(r, M[r], odb, 20, 40, 80, 5)
            assert [(a, b, c, list(x)) for a, b, c, x in idx.intervals(r)] == [(a, b, c, list(x)) for a, b, c, x in want]


# ------------------------------------------------------------------------------- bounds assertions (debug build)
def test_debug_build_bounds_assertions_stay_silent():
    """compute-sanitizer is not available on the target pool, so the library has a second build with
    device-side bounds assertions (trace / boundary / sequence / output indices, CIGAR tail vs unread trace).
    Run a spread of shapes through it -- packed and 32-bit paths, tiny and multi-stripe pairs, many chunks
    through the three-stream pipeline, ragged minimizer batches, spans -- and require an empty violation mask
    (and the usual bit-exactness)."""
    from isonform_b200 import _lib, hotpath, spans, synth
    dctx = _lib.Context(0, debug=True)
    try:
        assert dctx.debug_violations() == (0, True)
        dctx.debug_selftest()                                  # the reporting path itself works ...
        assert dctx.debug_violations() == (1 << 30, True)
        assert dctx.debug_violations() == (0, True)            # ... and reading clears it
        rng = np.random.default_rng(71)
        # alignment shapes
        seqs, pa, pb = [], [], []
        for (n, m) in [(1, 1), (2, 70), (70, 2), (63, 65), (64, 64), (65, 63), (129, 200), (192, 193), (255, 257), (256, 256),
                       (257, 255), (300, 1), (511, 700), (700, 511), (1025, 130), (130, 1025), (2047, 2049)]:
            x = _rand_dna(rng, max(n, m))
            y = synth.mutate(rng, x, 0.06) + _rand_dna(rng, m)
            seqs += [x[:n], y[:m]]; pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
        seqs += ["ACGTN" * 30, "ACGT" * 40]; pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)      # wildcard -> 32-bit
        big = _rand_dna(rng, 4500)
        seqs += [big, synth.mutate(rng, big, 0.04)]; pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)   # > 16-bit range
        big2 = _rand_dna(rng, 5100)
        seqs += [big2[:4800], synth.mutate(rng, big2, 0.05)]; pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)   # a re-based task with the one above
        for packed in (True, False):
            dctx.set_packed(packed)
            for params in [(2, -2, 12, 1), (2, -8, 12, 1), (1, -1, 1, 1)]:
                res = hotpath.align_pairs(seqs, pa, pb, *params, ctx=dctx, features_delta_len=5)
                for p in range(0, len(pa), 3):
                    sc, _q, _r, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], *params)
                    assert int(res["score"][p]) == sc
                    assert res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]].tolist() == runs.tolist()
        dctx.set_packed(True)
        assert dctx.debug_violations()[0] == 0
        # all pairs through many small chunks (pipeline of three streams); grouped by reference -> profile path
        base = _rand_dna(rng, 900)
        pool = [synth.mutate(rng, base, 0.05) for _ in range(20)]
        ib, ia = np.tril_indices(len(pool), -1)
        dctx.set_trace_budget(6 << 20)
        a = hotpath.align_pairs(pool, ia, ib, 2, -2, 12, 1, ctx=dctx, features_delta_len=5)
        dctx.set_trace_budget(0)
        b = hotpath.align_pairs(pool, ia, ib, 2, -2, 12, 1, ctx=dctx, features_delta_len=5)
        assert np.array_equal(a["cigar"], b["cigar"]) and np.array_equal(a["features"], b["features"])
        assert dctx.debug_violations()[0] == 0
        # short-pair regime (thread-per-task kernel): ragged groups, wildcards, a partial last group
        # fused tiny-call kernel (trace in shared memory)
        for (n, m) in [(1, 1), (99, 101), (257, 120), (480, 490), (130, 700), (1300, 1250), (600, 2100)]:
            x = _rand_dna(rng, max(n, m)); y = synth.mutate(rng, x, 0.06) + _rand_dna(rng, m)
            res = hotpath.align_pairs([x[:n], y[:m]], [0], [1], 2, -8, 12, 1, ctx=dctx, features_delta_len=5)
            sc, _q, _r, runs = oracle.sg_align(x[:n], y[:m], 2, -8, 12, 1)
            assert int(res["score"][0]) == sc and res["cigar"].tolist() == runs.tolist()
        assert dctx.debug_violations()[0] == 0
        dctx.set_short(2)
        for n_pairs, wild in ((130, False), (77, True)):
            sseqs, spa, spb = _short_cases(np.random.default_rng(5 + n_pairs), n_pairs, wild)
            res = hotpath.align_pairs(sseqs, spa, spb, 2, -8, 12, 1, ctx=dctx, features_delta_len=5)
            for p in range(0, n_pairs, 9):
                sc, _q, _r, runs = oracle.sg_align(sseqs[spa[p]], sseqs[spb[p]], 2, -8, 12, 1)
                assert int(res["score"][p]) == sc
                assert res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]].tolist() == runs.tolist()
        assert dctx.debug_violations()[0] == 0
        # banded trace of long pairs (feature-only call), default and narrow band: stores stay inside the band window
        dctx.set_short(1)
        lx = _rand_dna(rng, 7000)
        lseqs = [lx, synth.mutate(rng, lx, 0.05), lx[:2500] + lx[3000:], lx[1500:6800], synth.mutate(rng, lx[:5200], 0.08),
                 _rand_dna(rng, 4700)]
        lia, lib_ = np.triu_indices(len(lseqs), 1)
        dctx.set_band(0)
        want = hotpath.align_pairs(lseqs, lia, lib_, 2, -2, 12, 1, ctx=dctx, want_cigar=False, features_delta_len=5)
        for hw in (1024, 64):
            dctx.set_band(hw)
            got = hotpath.align_pairs(lseqs, lia, lib_, 2, -2, 12, 1, ctx=dctx, want_cigar=False, features_delta_len=5)
            assert np.array_equal(got["score"], want["score"]) and np.array_equal(got["features"], want["features"])
            assert np.array_equal(got["cigar_off"], want["cigar_off"])
        dctx.set_band(1024)
        assert dctx.debug_violations()[0] == 0
        # minimizers + spans
        reads = {i + 1: ("r", _rand_dna(rng, int(rng.integers(0, 900)), "ACGT" if i % 4 else "AC"), "") for i in range(80)}
        M = hotpath.get_minimizers_and_positions(reads, 31, 20, "lex", dctx)
        assert M == oracle.get_minimizers_and_positions(reads, 31, 20, "lex")
        spans.SpanIndex(reads, M, 20, 40, 80, 5, dctx)
        assert dctx.debug_violations()[0] == 0
    finally:
        dctx.close()


def test_release_build_has_no_assertions(ctx):
    ctx.debug_selftest()
    assert ctx.debug_violations() == (0, False)


def test_bubble_decisions_batch_vs_oracle(ctx):
    """align_bubble_nodes' gates + alignment + parse_cigar_diversity for a batch of bubble-like consensus
    pairs (SimplifyGraph.py:630-664), against the oracle pair by pair."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(83)
    seqs, pa, pb = [], [], []
    for i in range(400):
        ln = int(rng.integers(1, 260)) if i % 9 else int(rng.integers(1, 12))
        x = _rand_dna(rng, ln)
        y = synth.mutate(rng, x, [0.0, 0.02, 0.08, 0.2][i % 4])
        if i % 6 == 0 and len(y) > 30:
            c = len(y) // 2; y = y[:c] + y[c + int(rng.integers(1, 14)):]
        if i % 10 == 3:
            y = y + _rand_dna(rng, int(rng.integers(1, 80)))
        seqs += [x, y or "G"]; pa.append(2 * i); pb.append(2 * i + 1)
    for dl in (5, 10):
        pop, aligned = hotpath.bubble_decisions(seqs, pa, pb, dl, ctx)
        for p in range(len(pa)):
            want = oracle.align_bubble_decision(seqs[pa[p]], seqs[pb[p]], dl)
            assert (bool(pop[p]), bool(aligned[p])) == want, (p, dl, len(seqs[pa[p]]), len(seqs[pb[p]]))
        assert aligned.any() and (~aligned).any() and pop.any() and (~pop).any()


def test_profile_path_same_reference_tasks(ctx):
    """Tasks whose two alignments share the reference take the shared-memory query-profile variant of the
    packed path.  Group all pairs of a pool by reference (column-major), compare with the oracle and with the
    same pairs issued in an order where no task shares a reference (plain packed path) and on the 32-bit path."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(97)
    base = _rand_dna(rng, 700)
    pool = [synth.mutate(rng, base, 0.06)[int(rng.integers(0, 40)):] for _ in range(26)]
    pool += [_rand_dna(rng, int(rng.integers(1, 300))) for _ in range(6)]         # short / unrelated, odd stripe shapes
    pool.sort(key=len)
    ib, ia = np.tril_indices(len(pool), -1)                                       # grouped by reference ib
    for params in [(2, -2, 12, 1), (2, -8, 12, 1), (3, -1, 4, 2)]:
        grouped = hotpath.align_pairs(pool, ia, ib, *params, ctx=ctx, features_delta_len=5)
        perm = rng.permutation(len(ia))                                           # breaks the grouping
        shuffled = hotpath.align_pairs(pool, ia[perm], ib[perm], *params, ctx=ctx, features_delta_len=5)
        inv = np.argsort(perm)
        assert np.array_equal(grouped["score"], shuffled["score"][inv])
        assert np.array_equal(grouped["features"], shuffled["features"][inv])
        ctx.set_packed(False)
        try:
            ref32 = hotpath.align_pairs(pool, ia, ib, *params, ctx=ctx, features_delta_len=5)
        finally:
            ctx.set_packed(True)
        for key in ("score", "cigar_off", "cigar", "features"):
            assert np.array_equal(grouped[key], ref32[key]), (params, key)
        for p in range(0, len(ia), 9):
            sc, _q, _r, runs = oracle.sg_align(pool[ia[p]], pool[ib[p]], *params)
            assert int(grouped["score"][p]) == sc
            assert grouped["cigar"][grouped["cigar_off"][p]:grouped["cigar_off"][p + 1]].tolist() == runs.tolist(), (params, p)
    # identical query AND reference in both halves, and self-alignment
    same = hotpath.align_pairs(pool, [3, 3, 5, 5, 7], [9, 9, 9, 9, 7], 2, -2, 12, 1, ctx=ctx)
    sc, _q, _r, runs = oracle.sg_align(pool[3], pool[9], 2, -2, 12, 1)
    assert same["score"][:2].tolist() == [sc, sc]
    assert same["cigar"][same["cigar_off"][1]:same["cigar_off"][2]].tolist() == runs.tolist()


def test_packed_profile_range_extremes_debug_build():
    """Shared-reference tasks (query-profile variant of the packed path) at the edges of the 16-bit range:
    longest eligible references, all-match and all-mismatch content, queries of very different lengths sharing
    a task (one half runs on padding rows for many stripes).  Debug build: the bounds assertions must stay
    silent; results must equal the oracle's."""
    from isonform_b200 import _lib, hotpath, synth
    rng = np.random.default_rng(131)
    dctx = _lib.Context(0, debug=True)
    try:
        for params, lmax in [((2, -2, 12, 1), 7722), ((2, -8, 12, 1), 7716)]:
            ref_long = _rand_dna(rng, lmax)
            ref_mid = _rand_dna(rng, 4093)
            refs = [ref_long, "C" * lmax, ref_mid, "G" * 4093]
            queries = ["A", "AC", _rand_dna(rng, 100), ref_long[3000:3255], ref_long[100:356], "A" * 257, ref_mid[:511],
                       synth.mutate(rng, ref_mid[200:900], 0.05), "A" * 4093, ref_mid, synth.mutate(rng, ref_mid, 0.03)[:4093],
                       "G" * 4093, "T" * 4000, ref_long[:4093], _rand_dna(rng, 3)]
            seqs = queries + refs
            pa, pb = [], []
            for r in range(len(refs)):                      # grouped by reference, lengths interleaved long/short
                order = list(range(len(queries)))
                rng.shuffle(order)
                pa += order; pb += [len(queries) + r] * len(order)
            dctx.profile_enable(True); dctx.profile_reset()
            res = hotpath.align_pairs(seqs, pa, pb, *params, ctx=dctx)
            prof = dctx.profile()
            assert prof["dp_cells_profile"] > 0 and prof["dp_cells_packed"] >= prof["dp_cells_profile"], prof
            for p in range(len(pa)):
                sc, eq, er, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], *params)
                assert (int(res["score"][p]), int(res["end_q"][p]), int(res["end_r"][p])) == (sc, eq, er), (params, p)
                assert res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]].tolist() == runs.tolist(), (params, p)
            assert dctx.debug_violations()[0] == 0
    finally:
        dctx.close()


def test_small_batch_path_matches_regular_path(ctx):
    """Calls with <= 1024 pairs over <= 1 MiB of bases use the fused set-up / CIGAR-finish kernels; one pair more
    sends the same pairs through the regular plan (encode, plan, scan, radix-sorted task order).  Results must be
    identical, including for out-of-order references, wildcards and pairs that fall to the 32-bit path."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(211)
    base = _rand_dna(rng, 400)
    seqs = [synth.mutate(rng, base, 0.08)[int(rng.integers(0, 30)):int(rng.integers(200, 400))] or "A" for _ in range(60)]
    seqs += [_rand_dna(rng, int(rng.integers(1, 90))) for _ in range(20)]
    seqs += ["ACGTNNACGT" * 7, "acgtacgtac" * 9, _rand_dna(rng, 4500)]          # wildcard, lower case, > 16-bit range
    n = len(seqs)
    pa = rng.integers(0, n, 1025).astype(np.int32)
    pb = rng.integers(0, n, 1025).astype(np.int32)
    for params in [(2, -2, 12, 1), (2, -8, 12, 1)]:
        big = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)             # regular path
        small = hotpath.align_pairs(seqs, pa[:1024], pb[:1024], *params, ctx=ctx, features_delta_len=5)   # fused path
        assert np.array_equal(small["score"], big["score"][:1024])
        assert np.array_equal(small["features"], big["features"][:1024])
        assert np.array_equal(small["cigar_off"], big["cigar_off"][:1025])
        assert np.array_equal(small["cigar"], big["cigar"][:big["cigar_off"][1024]])
        one = hotpath.align_pairs(seqs, pa[7:8], pb[7:8], *params, ctx=ctx)
        sc, eq, er, runs = oracle.sg_align(seqs[pa[7]], seqs[pb[7]], *params)
        assert (int(one["score"][0]), int(one["end_q"][0]), int(one["end_r"][0])) == (sc, eq, er)
        assert one["cigar"].tolist() == runs.tolist()


def test_merge_consensuses_driver_on_gpu_equals_oracle_backed_run(ctx):
    """The all-pairs driver (merge.merge_consensuses: speculative blocks + sequential replay) with the GPU behind
    it against the same driver over the oracle-backed test double -- same merges, same regenerated consensuses."""
    import copy
    from fake_ctx import FakeContext
    from isonform_b200 import merge, synth
    rng = np.random.default_rng(401)
    bases = [_rand_dna(rng, int(rng.integers(500, 900))) for _ in range(4)]
    seqs = []
    for i in range(48):
        b = bases[i % 4]
        s = synth.mutate(rng, b, 0.004 if i % 3 else 0.03)
        seqs.append(s[int(rng.integers(0, 12)):len(s) - int(rng.integers(0, 12))])
    seqs_by_id = {100 + i: s for i, s in enumerate(seqs)}
    support = {100 + i: [100 + i] * (60 if i % 5 == 0 else 1) for i in range(len(seqs))}

    def gen_all(all_c, alt, cbs, reads, work_dir, max_spoa):
        for id_ in cbs:
            all_c[id_] = seqs_by_id[id_]
            alt.append((id_, seqs_by_id[id_]))

    def regen(i1, i2, reads, cbs, wd, ms):
        a, b = seqs_by_id[i1], seqs_by_id[i2]
        return b[:len(b) // 2] + a[len(a) // 2:] if len(cbs[i2]) % 2 else b

    results = []
    for c in (ctx, FakeContext()):
        cbs = copy.deepcopy(support)
        stats = {}
        new = merge.merge_consensuses(cbs, None, 0.15, 5, None, 200, 30, 50, generate_all_consensuses=gen_all,
                                      generate_new_full_consensus=regen, ctx=c, stats=stats)
        results.append((new, cbs, stats))
    assert results[0][0] == results[1][0] and results[0][1] == results[1][1] and results[0][2] == results[1][2]
    assert len(results[0][1]) < len(seqs), "the set must actually merge something"


# ------------------------------------------------------------------------------- drop-in at the seams, on hardware
@pytest.mark.parametrize("cap_index", [0, 1])
def test_reference_seams_replayed_on_gpu(ctx, cap_index):
    """tests/golden/seam_capture.json holds what the UNMODIFIED reference `main()` passed through every hot-path seam on
    a synthetic batch (scripts/make_seam_capture.py, build container).  Here the recorded inputs go through the real GPU
    context and every recorded output must come back: the reference's control flow is a deterministic function of
    these values, so this is the entry-point drop-in claim proven on the B200."""
    from isonform_b200 import hotpath, merge, runner
    cap = load_golden("seam_capture.json")["captures"][cap_index]
    pool, P = cap["pool"], cap["params"]
    reads = {int(r): (acc, pool[si], "I" * len(pool[si])) for r, acc, si in cap["reads"]}
    # (1) get_minimizers_and_positions (main:159)
    M = hotpath.get_minimizers_and_positions(reads, P["w"], P["k"], "lex", ctx)
    assert {str(r): [p for _m, p in M[r]] for r in M} == cap["minimizers"]
    for r in M:
        assert all(m == reads[r][1][p:p + P["k"]] for m, p in M[r])
    # (2) the whole front half: what main() hands to generateGraphfromIntervals (main:531)
    graph_iv, new_reads, _skipped, _index = runner.batch_front_half(reads, P["k"], P["w"], P["xmin"], P["xmax"], P["delta_len"], ctx)
    got_iv = {str(g): [[a, b, c, list(arr)] for a, b, c, arr in v] for g, v in graph_iv.items()}
    assert got_iv == cap["intervals_for_graph"]
    assert {str(g): [v[0], v[1]] for g, v in new_reads.items()} == {g: [v[0], pool[v[1]]] for g, v in cap["new_all_reads"].items()}
    # (3) every consensus.parasail_alignment call of bubble popping and isoform merging: one prefetch, then the per-pair seam
    memo = merge.AlignmentMemo(ctx)
    by_params = {}
    for s1, s2, ma, mi, go, ge, _cig, _sc in cap["alignments"]:
        by_params.setdefault((ma, mi, go, ge), []).append((pool[s1], pool[s2]))
    for params, pairs in by_params.items():
        memo.prefetch_alignments(pairs, *params)
    for s1, s2, ma, mi, go, ge, cig, sc in cap["alignments"]:
        a1, a2, got_cig, tuples, got_sc = memo.parasail_alignment(pool[s1], pool[s2], ma, mi, go, ge)
        assert (got_cig, got_sc) == (cig, sc)
        assert len(a1) == len(a2) == sum(l for l, _ in tuples)
    assert memo.misses == 0
    # (4) IsoformGeneration.align_to_merge decisions
    for c1, c2, delta, dl, d3, d5, want in cap["merge_decisions"]:
        assert hotpath.align_to_merge(pool[c1], pool[c2], delta, dl, d3, d5, ctx) == want
    # (5) merge_consensuses (IsoformGeneration.py:464-507) through the speculative driver, consensus strings as recorded
    mc = cap["merge"]
    cbs = {int(i): list(v) for i, v in mc["curr_best_seqs_in"].items()}
    mreads = {int(r): (v[0], pool[v[1]], "") for r, v in mc["reads"].items()}
    regen = {(a, b): pool[s] for a, b, s in mc["regenerated"]}

    def gen_all(all_c, alt, cur, rd, wd, ms):
        for i, s in mc["all_consensuses"]:
            all_c[i] = pool[s]
            alt.append((i, pool[s]))

    out = merge.merge_consensuses(cbs, None, mc["delta"], mc["delta_len"], mreads, P["max_seqs_to_spoa"], P["delta_iso_len_3"],
                                  P["delta_iso_len_5"], generate_all_consensuses=gen_all,
                                  generate_new_full_consensus=lambda i1, i2, rd, cur, wd, ms: regen[(i1, i2)], ctx=ctx)
    assert {str(i): v for i, v in cbs.items()} == mc["curr_best_seqs_out"]
    assert {str(i): s for i, s in out.items()} == {i: pool[s] for i, s in mc["new_consensuses"].items()}


# ------------------------------------------------------------------------------- short-pair regime (sg_short.cu)
def _short_cases(rng, n_pairs, wild):
    from isonform_b200 import synth
    seqs, pa, pb = [], [], []
    alpha = "ACGTN" if wild else "ACGT"
    for i in range(n_pairs):
        n = int(rng.integers(1, 385)) if i % 7 else int(rng.integers(1, 20))
        x = _rand_dna(rng, n, alpha if i % 5 == 0 else "ACGT")
        kind = i % 4
        if kind == 0:
            y = _rand_dna(rng, int(rng.integers(1, 385)), "ACGT")
        elif kind == 1:
            y = (synth.mutate(rng, x, 0.08) or "A")[:384]
        elif kind == 2:
            y = _rand_dna(rng, int(rng.integers(1, 40)), "AC") + x[:200]
        else:
            y = x
        if wild and i % 11 == 3:
            y = y[:len(y) // 2] + "x" + y[len(y) // 2 + 1:]
        seqs += [x, y[:384] or "C"]
        pa.append(2 * i); pb.append(2 * i + 1)
    return seqs, pa, pb


@pytest.mark.parametrize("n_pairs,wild", [(1, False), (65, False), (700, True), (3001, False), (1500, True)])
def test_short_pair_regime_matches_long_path_and_oracle(ctx, n_pairs, wild):
    """All-short calls take the thread-per-task kernel; the same call forced onto the wavefront kernel and the oracle
    must agree bit for bit (scores, end cells, CIGAR runs, decision features), for the fused (P <= 1024) and the regular
    set-up, odd pair counts, wildcards, both bubble and isoform scores."""
    from isonform_b200 import hotpath
    rng = np.random.default_rng(1000 + n_pairs)
    seqs, pa, pb = _short_cases(rng, n_pairs, wild)
    for params in [(2, -8, 12, 1), (2, -2, 12, 1), (1, -1, 1, 1)]:
        ctx.set_short(2)                                   # force the short kernel whatever the pair count
        try:
            a = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)
            plain = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx)
            ctx.set_short(0)
            b = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)
        finally:
            ctx.set_short(1)
        for key in ("score", "cigar_off", "cigar", "features"):
            assert np.array_equal(a[key], b[key]), (params, key)
        assert np.array_equal(plain["cigar"], a["cigar"]) and np.array_equal(plain["score"], a["score"])
        step = max(1, n_pairs // 60)
        for p in range(0, n_pairs, step):
            sc, eq, er, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], *params)
            assert (int(plain["score"][p]), int(plain["end_q"][p]), int(plain["end_r"][p])) == (sc, eq, er), (params, p)
            assert plain["cigar"][plain["cigar_off"][p]:plain["cigar_off"][p + 1]].tolist() == runs.tolist(), (params, p)


@pytest.mark.parametrize("tb_index", [1, 2, 3, 4, 9, 15])
def test_short_pair_regime_under_tiebreak_tables(ctx, tb_index):
    tb = oracle.all_tiebreaks()[tb_index]
    rng = np.random.default_rng(77)
    seqs, pa, pb = [], [], []
    for i in range(90):
        alpha = ["AC", "A", "ACG", "AAC", "ACGT"][i % 5]
        seqs += [_rand_dna(rng, int(rng.integers(1, 90)), alpha), _rand_dna(rng, int(rng.integers(1, 90)), alpha)]
        pa.append(2 * i); pb.append(2 * i + 1)
    ctx.set_tiebreak(tb.h_e_before_f, tb.gap_open_on_tie, tb.end_col_first, tb.swap_id)
    ctx.set_short(2)
    try:
        for params in [(2, -8, 12, 1), (1, -1, 1, 1), (2, -2, 2, 1), (2, -2, 0, 0)]:
            _check_pairs(ctx, seqs, pa, pb, params, tiebreak=tb)
    finally:
        ctx.set_short(1)
        ctx.set_tiebreak(0, 0, 0, 0)


# ------------------------------------------------------------------------------- re-based 16-bit lanes (long reads)
def test_rebased_16bit_lanes_match_32bit_path_and_oracle(ctx):
    """Pairs beyond the plain 16-bit score range (BASELINE config 4: 10 kb reads) run in 16-bit lanes that follow a
    running base (dp_pair16<., RB>).  66 pairs -- C4 reads, identical sequences (score 2n: the largest value there is),
    unrelated sequences (the most negative ones), exon-sized deletions and insertions, unequal lengths, a low-complexity
    pair -- must equal the 32-bit path on every output and the oracle on a sample; under two more tie-break tables too."""
    from isonform_b200 import hotpath, polya, synth
    rng = np.random.default_rng(4242)
    reads_l, _iso, _t = synth.make_config("C4", n_reads=10)
    seqs = [polya.remove_read_polyA_ends(s, 12, 1) for _, s in reads_l]
    ia, ib = [int(x) for x in np.triu_indices(10, 1)[0]], [int(x) for x in np.triu_indices(10, 1)[1]]       # 45 pairs

    def add(a, b):
        seqs.extend([a, b]); ia.append(len(seqs) - 2); ib.append(len(seqs) - 1)
    x = _rand_dna(rng, 9000)
    add(x, x)                                                    # identical: score 18000 (72000 in tagged units)
    add(x, _rand_dna(rng, 8800))                                 # unrelated
    add(x, x[:3000] + x[3600:])                                  # 600 bp deletion
    add(x[:4000] + x[4500:], x)                                  # 500 bp insertion
    add(x[2000:7000], x)                                         # contained: long free end gaps
    add(x, x[1000:6200])
    add(synth.mutate(rng, x, 0.12), x)
    add(_rand_dna(rng, 5000, "AC"), _rand_dna(rng, 5200, "AC"))  # tie-heavy
    for _ in range(13):
        n = int(rng.integers(4200, 9000))
        y = _rand_dna(rng, n)
        z = synth.mutate(rng, y, float(rng.choice([0.01, 0.05, 0.1])))
        if rng.random() < 0.5:
            c = n // 3; z = z[:c] + z[c + int(rng.integers(50, 900)):]
        add(y, z)
    assert len(ia) == 66
    for tb in (oracle.TieBreak(0, 0, 0, 0), oracle.TieBreak(1, 0, 1, 0), oracle.TieBreak(0, 1, 0, 1)):
        ctx.set_tiebreak(tb.h_e_before_f, tb.gap_open_on_tie, tb.end_col_first, tb.swap_id)
        try:
            ctx.set_packed(True)
            a = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
            ctx.set_packed(False)                                # 32-bit path only
            b = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
        finally:
            ctx.set_packed(True)
            ctx.set_tiebreak(0, 0, 0, 0)
        for key in ("score", "cigar_off", "cigar", "features"):
            assert np.array_equal(a[key], b[key]), (key, tb.h_e_before_f, tb.gap_open_on_tie)
        ea = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx)          # end cells (default table only below)
        if tb.h_e_before_f == 0 and tb.gap_open_on_tie == 0:
            assert int(a["score"][45]) == 18000
            for p in (0, 17, 44, 45, 46, 47, 48, 49, 50, 52, 60):
                sc, eq, er, runs = oracle.sg_align(seqs[ia[p]], seqs[ib[p]], 2, -2, 12, 1)
                assert (int(ea["score"][p]), int(ea["end_q"][p]), int(ea["end_r"][p])) == (sc, eq, er), p
                assert ea["cigar"][ea["cigar_off"][p]:ea["cigar_off"][p + 1]].tolist() == runs.tolist(), p
    prof0 = ctx.profile()
    ctx.profile_enable(True); ctx.profile_reset()
    hotpath.align_pairs(seqs, ia[:44], ib[:44], 2, -2, 12, 1, ctx=ctx, want_cigar=False, features_delta_len=5)
    pr = ctx.profile(); ctx.profile_enable(False)
    assert pr["dp_cells_packed"] == pr["dp_cells"] > 0           # every pair of the call took the 16-bit lanes
    del prof0


# ------------------------------------------------------------------------------- kernel (c): exact k-mer identity
@pytest.mark.parametrize("mask", [0x7fffffffffffffff, 0xF, 0x0])
def test_spans_exact_kmer_identity_under_forced_hash_collisions(ctx, mask):
    """k-mers that do not pack into 2 bits per base (N, lower case, k >= 32) are grouped by a hash and then told apart
    by comparing bytes.  With the hash cut to 4 bits -- or to nothing: every such k-mer collides with every other --
    the pair database, the span search and the WIS picks must still be the reference's (main:174-212, 308-338)."""
    from isonform_b200 import synth
    rng = np.random.default_rng(88)
    base = _rand_dna(rng, 700)
    # lower-case stretches, Ns and a lower-case/upper-case twin of the same region: distinct strings for the reference
    low = base[:200] + base[200:330].lower() + base[330:450] + "N" + base[451:520] + "NN" + base[522:]
    reads = {}
    for i in range(9):
        src = base if i % 3 == 0 else low
        s = synth.mutate(rng, src, 0.01)                              # mutate keeps the case of untouched bases
        reads[i + 1] = ("r%d" % i, s, "")
    reads[10] = ("twin", low, "")
    reads[11] = ("twin2", low, "")
    reads[12] = ("upper", base, "")
    ctx.debug_set_span_hash_mask(mask)
    try:
        _span_case_check(ctx, reads, 20, 31, 40, 80, 5)
        _span_case_check(ctx, reads, 33, 45, 66, 120, 5)             # k >= 32: every k-mer goes through the byte comparison
        _span_case_check(ctx, reads, 9, 20, 18, 60, 10)
    finally:
        ctx.debug_set_span_hash_mask()


# ------------------------------------------------------------------------------- fused tiny-call path
def test_fused_small_call_matches_regular_path_and_oracle(ctx):
    """Host-buffer calls of a few small pairs run as one fused launch (trace in shared memory, mapped I/O).  Same
    scores, end cells, CIGAR runs and features as the regular path and the oracle: single pairs from 1 x 1 to multi-stripe
    shapes, wildcards and lower case, batches of up to 32 pairs, the capacity-retry protocol, a non-default tie-break
    table, pairs whose trace does not fit shared memory (it then lives in device memory, up to 4096 x 4096), and shapes
    just beyond the fused limits (which must fall through to the regular path unnoticed)."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(2024)
    cases = []
    for (n, m) in [(1, 1), (1, 60), (60, 1), (7, 9), (99, 101), (256, 256), (257, 120), (300, 310), (520, 200), (130, 700),
                   (480, 490), (800, 800), (20, 3000), (1500, 1400), (4096, 700), (2500, 2600), (4097, 300)]:
        x = _rand_dna(rng, max(n, m))
        y = synth.mutate(rng, x, 0.06) + _rand_dna(rng, m)
        cases.append(([x[:n], y[:m]], [0], [1]))
    cases.append((["ACGTNNACGTacgtACGTX" * 6, "ACGTACGTACGTacgt" * 7], [0], [1]))
    cases.append((["X" * 40, lcg(45, 42)], [0, 1], [1, 0]))
    pool = [synth.mutate(rng, _rand_dna(rng, 150), 0.05) for _ in range(12)] + ["ACGNNT" * 20]
    ia, ib = np.triu_indices(len(pool), 1)
    cases.append((pool, ia[:32].tolist(), ib[:32].tolist()))
    cases.append((pool, ia[:33].tolist(), ib[:33].tolist()))             # 33 pairs: beyond the fused limit
    for tb in ((0, 0, 0, 0), (1, 1, 1, 1)):
        ctx.set_tiebreak(*tb)
        try:
            for seqs, pa, pb in cases:
                for params in [(2, -8, 12, 1), (2, -2, 12, 1)]:
                    ctx.set_fused_small(True)
                    a = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx)
                    af = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, want_cigar=True, features_delta_len=5)
                    ctx.set_fused_small(False)
                    b = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx)
                    bf = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, want_cigar=True, features_delta_len=5)
                    ctx.set_fused_small(True)
                    for key in ("score", "end_q", "end_r", "cigar_off", "cigar"):
                        assert np.array_equal(a[key], b[key]), (key, len(seqs[pa[0]]), len(seqs[pb[0]]), params, tb)
                    for key in ("score", "cigar_off", "cigar", "features"):
                        assert np.array_equal(af[key], bf[key]), (key, len(seqs[pa[0]]), len(seqs[pb[0]]), params, tb)
        finally:
            ctx.set_fused_small(True)
            ctx.set_tiebreak(0, 0, 0, 0)
    # against the oracle (default table) and through the reference-shaped seam
    for seqs, pa, pb in cases[:19]:
        res = hotpath.align_pairs(seqs, pa, pb, 2, -8, 12, 1, ctx=ctx)
        for p in range(len(pa)):
            sc, eq, er, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], 2, -8, 12, 1)
            assert (int(res["score"][p]), int(res["end_q"][p]), int(res["end_r"][p])) == (sc, eq, er)
            assert res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]].tolist() == runs.tolist()
    cat, off = hotpath.pack_sequences(pool)
    full = ctx.sg_align_pairs(cat, off, ia[:20], ib[:20])
    small = ctx.sg_align_pairs(cat, off, ia[:20], ib[:20], cigar_cap=3)       # ISOF_ERR_CAPACITY + retry on the fused path
    assert small["cigar"].tolist() == full["cigar"].tolist()
    a1, a2, cig, tup, score = hotpath.parasail_alignment(pool[0], pool[1], 2, -8, 12, 1, ctx)
    o = oracle.parasail_alignment(pool[0], pool[1], 2, -8, 12, 1)
    assert (a1, a2, cig, tup, score) == (o[0], o[1], o[2], o[3], o[4])


# ------------------------------------------------------------------------------- randomised differential test
@pytest.mark.parametrize("seed", range(int(os.environ.get("ISOF_FUZZ_SEEDS", "64"))))
def test_randomised_regimes_vs_oracle(ctx, seed):
    """Random batch compositions x random scoring parameters x random tie-break table x random regime switches (packed
    on/off, short-pair kernel forced/never, fused tiny-call path on/off -- incl. its multi-stripe chase mode --, traceback
    per thread / per warp, trace band width): every pair must equal the oracle under the same table.  Catches
    interactions no hand-written case covers.  ISOF_FUZZ_SEEDS=N widens the sweep (development)."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(7000 + seed)
    n_pairs = int(rng.choice([1, 3, 17, 33, 64, 65, 130, 400]))
    max_len = int(rng.choice([12, 60, 150, 384, 700, 1500, 3000]))
    if max_len >= 1500:
        n_pairs = min(n_pairs, 33 if max_len == 1500 else 17)
    alphabet = str(rng.choice(["ACGT", "ACGT", "AC", "ACGTN", "ACGTacgtN"]))
    seqs, pa, pb = [], [], []
    for i in range(n_pairs):
        n = int(rng.integers(1, max_len + 1))
        x = _rand_dna(rng, n, alphabet)
        kind = int(rng.integers(0, 4))
        if kind == 0:
            y = _rand_dna(rng, int(rng.integers(1, max_len + 1)), alphabet)
        elif kind == 1:
            y = synth.mutate(rng, x, float(rng.choice([0.0, 0.02, 0.1]))) or "A"
        elif kind == 2:
            y = _rand_dna(rng, int(rng.integers(0, 30)), alphabet) + x[int(rng.integers(0, n)):]
            y = y or "C"
        else:
            y = x[:int(rng.integers(1, n + 1))]
        seqs += [x, y]
        pa.append(2 * i); pb.append(2 * i + 1)
    if rng.random() < 0.3:                                     # shared references: the query-profile path
        pb = [pb[0]] * n_pairs
    match = int(rng.integers(1, 4))
    mismatch = -int(rng.integers(0, 9))
    gap_open = int(rng.integers(0, 14))
    gap_ext = int(rng.integers(0, 3))
    params = (match, mismatch, gap_open, gap_ext)
    tb = oracle.all_tiebreaks()[int(rng.integers(0, 16))]
    packed, short_mode, fused = bool(rng.integers(0, 2)), int(rng.choice([0, 1, 2])), bool(rng.integers(0, 2))
    tb_mode, band = int(rng.choice([0, 1, 2])), int(rng.choice([0, 64, 1024]))
    ctx.set_tiebreak(tb.h_e_before_f, tb.gap_open_on_tie, tb.end_col_first, tb.swap_id)
    ctx.set_packed(packed); ctx.set_short(short_mode); ctx.set_fused_small(fused)
    ctx.set_traceback_warp(tb_mode); ctx.set_band(band)
    try:
        res = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx)
        feat = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, want_cigar=False, features_delta_len=5)
    finally:
        ctx.set_tiebreak(0, 0, 0, 0); ctx.set_packed(True); ctx.set_short(1); ctx.set_fused_small(True)
        ctx.set_traceback_warp(2); ctx.set_band(1024)
    assert np.array_equal(res["score"], feat["score"])
    for p in range(n_pairs):
        sc, eq, er, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], *params, tiebreak=tb)
        got = res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]]
        ctxt = (seed, p, params, (tb.h_e_before_f, tb.gap_open_on_tie, tb.end_col_first, tb.swap_id), packed, short_mode, fused,
                tb_mode, band, len(seqs[pa[p]]), len(seqs[pb[p]]))
        assert (int(res["score"][p]), int(res["end_q"][p]), int(res["end_r"][p])) == (sc, eq, er), ctxt
        assert got.tolist() == runs.tolist(), ctxt
        tup = oracle.cigar_runs_to_tuples(runs)
        f = feat["features"][p]
        assert int(f[0]) == sum(l for l, _ in tup) and int(f[1]) == len(tup), ctxt
        assert int(f[2]) == sum(l for l, o in tup if o != "=") and int(f[3]) == max([l for l, o in tup if o != "="] + [0]), ctxt


# ------------------------------------------------------------------------------- banded trace of long pairs
@pytest.mark.parametrize("half_width", [1024, 96, 8])
def test_banded_trace_of_long_pairs_never_changes_results(ctx, half_width):
    """Feature-only calls store, for pairs of the re-based class, only a band of trace steps around the diagonals the
    free end gaps make room for; a pair whose optimal path leaves the band is aligned again with the full trace
    (isof_set_band).  Scores, features and run offsets must equal the unbanded call for every band width: 1024 (the
    default: exon-sized indels stay inside), 96 and 8 (most pairs with an indel fail and take the second pass)."""
    from isonform_b200 import hotpath, polya, synth
    rng = np.random.default_rng(977 + half_width)
    reads_l, _iso, _t = synth.make_config("C4", n_reads=8)
    seqs = [polya.remove_read_polyA_ends(s, 12, 1) for _, s in reads_l]
    ia, ib = [int(x) for x in np.triu_indices(8, 1)[0]], [int(x) for x in np.triu_indices(8, 1)[1]]          # 28 pairs

    def add(a, b):
        seqs.extend([a, b]); ia.append(len(seqs) - 2); ib.append(len(seqs) - 1)
    x = _rand_dna(rng, 9000)
    add(x, x)
    add(x, x[:3000] + x[3600:])                                  # 600 bp deletion in the middle: path leaves a narrow band
    add(x[:4000] + x[4500:], x)
    add(x[:2000] + x[2400:6000] + _rand_dna(rng, 400) + x[6000:], x)        # deletion then insertion: ends on the main diagonal
    add(x[2000:7000], x)                                         # contained: the band follows the free end gaps
    add(x, x[1000:6200])
    add(x, _rand_dna(rng, 8800))                                 # unrelated
    add(_rand_dna(rng, 5000, "AC"), _rand_dna(rng, 5200, "AC"))
    add(x[:300], x)                                              # short vs long (plain 16-bit class or re-based, never banded wrongly)
    for _ in range(12):
        n = int(rng.integers(4200, 9000))
        y = _rand_dna(rng, n)
        z = synth.mutate(rng, y, float(rng.choice([0.01, 0.05, 0.1])))
        if rng.random() < 0.6:
            c = int(rng.integers(200, n - 1200)); z = z[:c] + z[c + int(rng.integers(20, 900)):]
        add(y, z) if rng.random() < 0.5 else add(z, y)
    try:
        ctx.set_band(0)
        ref = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, want_cigar=False, features_delta_len=5)
        ref_full = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
        ref_runs = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx)
        ctx.set_band(half_width)
        ctx.profile_enable(True); ctx.profile_reset()
        got = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, want_cigar=False, features_delta_len=5)
        pr = ctx.profile(); ctx.profile_enable(False)
        # a call that returns the runs: the stream is spliced together from the two passes
        full = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
        runs = hotpath.align_pairs(seqs, ia, ib, 2, -2, 12, 1, ctx=ctx)
        cat, off = hotpath.pack_sequences(seqs)
        tight = ctx.sg_align_pairs(cat, off, np.asarray(ia, np.int32), np.asarray(ib, np.int32), 2, -2, 12, 1, cigar_cap=1000)
    finally:
        ctx.set_band(1024)
        ctx.profile_enable(False)
    for key in ("score", "features", "cigar_off"):
        assert np.array_equal(got[key], ref[key]), (key, half_width)
        assert np.array_equal(full[key], ref[key]), (key, half_width)
    for key in ("score", "features", "cigar_off", "cigar"):
        assert np.array_equal(full[key], ref_full[key]), (key, half_width)
    for key in ("score", "end_q", "end_r", "cigar_off", "cigar"):
        assert np.array_equal(runs[key], ref_runs[key]), (key, half_width)
        assert np.array_equal(tight[key], ref_runs[key]), (key, half_width)      # capacity-retry protocol across both passes
    retried = pr["band_pairs_retried"]
    if half_width == 1024:
        assert retried <= 2, retried                             # only the unrelated / low-complexity pairs may wander
    else:
        assert retried >= 3, retried                             # the exon-sized indels do leave a narrow band


def test_trace_lanes_shrink_when_the_device_fills_up_after_first_use():
    """The trace budget is sized from the memory that was free when the context first aligned.  If the host framework
    has since filled the device (here: a torch tensor that leaves ~24 GB), the lanes no longer fit: the call must re-plan
    with smaller chunks -- handing back the lanes it already got -- and return the same results, not ISOF_ERR_NOMEM."""
    import torch
    from isonform_b200 import _lib, hotpath, synth
    rng = np.random.default_rng(31)
    base = _rand_dna(rng, 2400)
    pool = [synth.mutate(rng, base, 0.05) for _ in range(260)]
    ib, ia = np.tril_indices(len(pool), -1)                      # 33 670 pairs, ~97 GB of trace: many chunks
    c = _lib.Context(0)
    hog = None
    try:
        hotpath.align_pairs(pool[:40], ia[:700], ib[:700], 2, -2, 12, 1, ctx=c)              # first use: budget from the free memory now
        free, _total = torch.cuda.mem_get_info()
        if free < (60 << 30):
            pytest.skip("needs a mostly empty device")
        hog = torch.empty(free - (24 << 30), dtype=torch.uint8, device="cuda")
        c.profile_reset()
        got = hotpath.align_pairs(pool, ia, ib, 2, -2, 12, 1, ctx=c, want_cigar=False, features_delta_len=5)
        assert c.profile()["trace_replans"] >= 1
        del hog
        hog = None
        torch.cuda.empty_cache()
        want = hotpath.align_pairs(pool, ia, ib, 2, -2, 12, 1, ctx=c, want_cigar=False, features_delta_len=5)
        assert np.array_equal(got["score"], want["score"]) and np.array_equal(got["features"], want["features"])
        sc, _q, _r, _runs = oracle.sg_align(pool[ia[777]], pool[ib[777]], 2, -2, 12, 1)
        assert int(got["score"][777]) == sc
    finally:
        del hog
        torch.cuda.empty_cache()
        c.close()


def test_warp_traceback_equals_thread_traceback(ctx):
    """The regular path's traceback runs one warp per pair (look-ahead along the diagonal / the gap); the round-1
    thread-per-pair kernel stays selectable.  Same runs, features and scores on a mixed batch: packed, 32-bit (wildcards),
    re-based and banded pairs, compacted last stripes, exon-sized gaps longer than one look-ahead, many chunks."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(99)
    seqs, pa, pb = [], [], []

    def add(a, b):
        seqs.extend([a, b]); pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    for (n, m) in [(1, 1), (3, 70), (70, 3), (255, 257), (257, 255), (700, 1500), (1500, 700), (2047, 2049), (33, 2000)]:
        x = _rand_dna(rng, max(n, m)); y = synth.mutate(rng, x, 0.07) + _rand_dna(rng, m)
        add(x[:n], y[:m])
    x = _rand_dna(rng, 3000)
    add(x, x[:1000] + x[1400:])                                  # 400 bp gap: 13 look-aheads long
    add(x[:500] + x[1500:], x)
    add("ACGTN" * 200, "ACGT" * 260)                             # wildcards: 32-bit path, nibble layout
    add("AC" * 600, "AC" * 630)                                  # tie-heavy
    big = _rand_dna(rng, 6000)
    add(big, synth.mutate(rng, big, 0.05))                       # re-based class
    add(big[:2000] + big[2300:], synth.mutate(rng, big, 0.03))
    base = _rand_dna(rng, 1100)
    pool = [synth.mutate(rng, base, 0.05) for _ in range(24)]
    o = len(seqs)
    seqs += pool
    ib, ia = np.tril_indices(len(pool), -1)
    pa += (ia + o).tolist(); pb += (ib + o).tolist()
    for budget in (0, 8 << 20):
        ctx.set_trace_budget(budget)
        try:
            ctx.set_traceback_warp(True)
            a = hotpath.align_pairs(seqs, pa, pb, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
            af = hotpath.align_pairs(seqs, pa, pb, 2, -2, 12, 1, ctx=ctx, want_cigar=False, features_delta_len=5)
            ctx.set_traceback_warp(False)
            b = hotpath.align_pairs(seqs, pa, pb, 2, -2, 12, 1, ctx=ctx, features_delta_len=5)
            bf = hotpath.align_pairs(seqs, pa, pb, 2, -2, 12, 1, ctx=ctx, want_cigar=False, features_delta_len=5)
        finally:
            ctx.set_traceback_warp(2)
            ctx.set_trace_budget(0)
        for key in ("score", "cigar_off", "cigar", "features"):
            assert np.array_equal(a[key], b[key]), (key, budget)
        for key in ("score", "cigar_off", "features"):
            assert np.array_equal(af[key], bf[key]) and np.array_equal(af[key], a[key]), (key, budget)
    for p in range(15):
        sc, _q, _r, runs = oracle.sg_align(seqs[pa[p]], seqs[pb[p]], 2, -2, 12, 1)
        assert int(a["score"][p]) == sc and a["cigar"][a["cigar_off"][p]:a["cigar_off"][p + 1]].tolist() == runs.tolist(), p
