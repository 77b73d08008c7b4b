"""CPU: the oracle (oracle/) against the golden vectors produced by executing the reference
(scripts/make_golden.py) and against the SURVEY.md section 8c known-answer vectors."""
import hashlib
import os
import subprocess
import sys
from array import array

import numpy as np
import pytest

import oracle
from conftest import lcg, load_golden


def test_pyhash_matches_reference_interpreter_hashes():
    g = load_golden("hashes.json")
    assert g["algorithm"] == "siphash13"
    for c in g["cases"]:
        assert oracle.pyhash(c["s"]) == c["hash"], c["s"]
    assert oracle.pyhash("ACGTACGTACGTACGTACGT") == -5958443965306619372


def test_pyhash_matches_live_interpreter_under_seed0():
    """Pin against the running CPython's hash() in a PYTHONHASHSEED=0 child."""
    rng = np.random.default_rng(5)
    strs = ["".join("ACGTN"[int(x)] for x in rng.integers(0, 5, size=int(n))) for n in rng.integers(0, 70, size=300)]
    code = "import sys,json; print(json.dumps([hash(s) for s in json.loads(sys.stdin.read())]))"
    import json
    env = dict(os.environ, PYTHONHASHSEED="0")
    out = subprocess.run([sys.executable, "-c", code], input=json.dumps(strs), text=True, capture_output=True,
                         env=env, check=True).stdout
    if sys.hash_info.algorithm != "siphash13":
        pytest.skip("interpreter does not use siphash13")
    assert [oracle.pyhash(s) for s in strs] == json.loads(out)


def test_minimizers_golden_cases():
    g = load_golden("minimizers.json")
    for c in g["cases"]:
        got = oracle.minimizers(c["seq"], c["k"], c["w"]).tolist()
        assert got == c["pos"], c["name"]
        assert [p for _, p in oracle.get_kmer_minimizers(c["seq"], c["k"], c["w"])] == c["pos"]


def test_minimizers_survey_kats():
    s = lcg(200, 42)
    assert s.startswith("CCCCACCACCAGGACACTTTCAGAGTTCTC")
    assert oracle.minimizers(s, 20, 31).tolist() == [9, 13, 16, 19, 29, 40, 43, 54, 60, 68, 75, 76, 86, 94, 100, 108,
                                                     118, 121, 131, 139, 142, 152, 157, 158, 170, 171]
    assert oracle.minimizers(s, 9, 20).tolist() == [6, 14, 21, 29, 39, 41, 49, 50, 62, 74, 76, 85, 91, 92, 98, 106,
                                                    109, 111, 117, 119, 121, 123, 135, 136, 140, 151, 163, 173, 183]
    kat4 = "ACGTTGCA" * 10 + lcg(60, 3) + "AC" * 25
    assert oracle.minimizers(kat4, 20, 31).tolist() == [4, 12, 20, 28, 36, 44, 52, 60, 61, 63, 64, 69, 74, 82, 85, 91,
                                                        95, 101, 108, 118, 121, 130, 137, 149, 161]


def test_minimizers_kat3_digest():
    g = load_golden("minimizers.json")["kat3"]
    pos = oracle.minimizers(lcg(g["n"], g["seed"]), g["k"], g["w"]).tolist()
    assert len(pos) == g["count"] == 15398
    assert hashlib.sha256(",".join(map(str, pos)).encode()).hexdigest()[:16] == g["sha256_16"] == "f2fd9a03ebacf568"


def test_minimizers_batch_matches_single():
    rng = np.random.default_rng(3)
    seqs = ["".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=int(n))) for n in rng.integers(0, 400, size=40)]
    cat = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    pos, poff = oracle.minimizers_batch(cat, off, 20, 31, threads=2)
    for i, s in enumerate(seqs):
        assert pos[poff[i]:poff[i + 1]].tolist() == oracle.minimizers(s, 20, 31).tolist()


def test_polya_golden():
    for c in load_golden("polya.json"):
        assert oracle.remove_read_polyA_ends(c["seq"], 12, 1) == c["out"]


def test_pairs_spans_wis_golden():
    for case in load_golden("spans.json"):
        k, w, xmin, xmax, dl = case["k"], case["w"], case["xmin"], case["xmax"], case["delta_len"]
        reads = {int(r): ("acc", s, "") for r, s in case["reads"].items()}
        M = oracle.get_minimizers_and_positions(reads, w, k, "lex")
        assert {str(r): [p for _, p in M[r]] for r in M} == case["minimizers"], case["name"]
        db = oracle.get_minimizer_combinations_database(M, k, xmin, xmax, reads)
        got = [[m1, m2, list(db[m1][m2])] for m1 in db for m2 in db[m1] if len(db[m1][m2])]
        assert got == case["db"], case["name"]
        for r_id in sorted(reads):
            exp = case["per_read"][str(r_id)]
            iv = oracle.pyside.read_intervals(r_id, M[r_id], db, k, xmin, xmax, dl)
            assert [[a, b, c, hashlib.sha256(array("I", d).tobytes()).hexdigest()[:12]] for a, b, c, d in iv] == exp["intervals"]
            if iv:
                srt = sorted(iv, key=lambda x: x[1])
                chosen = [srt[j] for j in oracle.solve_WIS(srt)[::-1]]
                assert [[a, b, c, list(d)] for a, b, c, d in chosen] == exp["wis"]
            else:
                assert exp["wis"] is None


def test_kat5_first_intervals():
    case = [c for c in load_golden("spans.json") if c["name"] == "KAT5"][0]
    iv = case["per_read"]["1"]["intervals"]
    assert len(iv) == 85 and iv[0][:3] == [29, 86, 3] and iv[3][:3] == [29, 68, 4]
    assert [w[:3] for w in case["per_read"]["1"]["wis"]] == [[29, 60, 4], [60, 118, 3], [120, 171, 4]]


def test_decisions_golden():
    g = load_golden("decisions.json")
    for rec in g["pairs"]:
        a, b = rec["s1"], rec["s2"]
        a1, a2, cig, tup, score = oracle.parasail_alignment(a, b, 2, -8, 12, 1)
        assert (cig, score) == (rec["bubble"]["cigar"], rec["bubble"]["score"]), rec["name"]
        assert hashlib.sha256(a1.encode()).hexdigest()[:12] == rec["bubble"]["a1_sha"]
        assert oracle.parse_cigar_diversity(tup, 0.20, 5) == rec["bubble"]["pop_d5"]
        assert oracle.parse_cigar_diversity(tup, 0.20, 10) == rec["bubble"]["pop_d10"]
        a1, a2, cig, tup, score = oracle.parasail_alignment(a, b, 2, -2, 12, 1)
        assert (cig, score) == (rec["iso"]["cigar"], rec["iso"]["score"]), rec["name"]
        assert oracle.find_first_significant_match(a1, a2, 30, 0.7) == rec["iso"]["start_match"]
        assert oracle.find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7) == rec["iso"]["end_match"]
        for dl, d in ((5, 0.15), (10, 0.15), (5, 0.1)):
            assert oracle.align_to_merge(a, b, d, dl, 30, 50) == rec["iso"]["merge_dl%d_d%s" % (dl, d)], rec["name"]
    for gt in g["gates"]:
        assert oracle.align_bubble_decision(gt["s1"], gt["s2"], gt["delta_len"]) == (gt["pop"], gt["aligned"])


# ---------------------------------------------------------------- alignment restatement sanity
def _score_of_cigar(s1, s2, runs, match, mismatch, gap_open, gap_ext):
    """Re-score a cigar under semi-global rules: leading/trailing gap runs are free."""
    tup = oracle.cigar_runs_to_tuples(runs)
    i = j = 0
    sc = 0
    for idx, (ln, op) in enumerate(tup):
        edge = idx == 0 or idx == len(tup) - 1
        if op in "=X":
            for q in range(ln):
                a, b = s1[i + q].upper(), s2[j + q].upper()
                if a not in "ACGT" or b not in "ACGT":
                    pass
                else:
                    sc += match if a == b else mismatch
                assert (a == b) == (op == "=")
            i += ln; j += ln
        elif op == "I":
            i += ln
            sc -= 0 if edge else gap_open + (ln - 1) * gap_ext
        else:
            j += ln
            sc -= 0 if edge else gap_open + (ln - 1) * gap_ext
    assert i == len(s1) and j == len(s2)
    return sc


def test_sg_align_cigar_consistent_and_optimal_small():
    """The cigar covers both sequences, re-scores to the reported score, and the score equals a
    brute-force numpy Gotoh maximum over the last row/column."""
    rng = np.random.default_rng(9)
    for it in range(60):
        n, m = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        s1 = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=n))
        s2 = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=m))
        for (ma, mi) in ((2, -8), (2, -2)):
            score, eq, er, runs = oracle.sg_align(s1, s2, ma, mi, 12, 1)
            assert eq == n - 1 or er == m - 1
            # leading and trailing gap runs may be interior gaps that merged with an end gap: the
            # rescoring treats whole edge runs as free, so it can only over-estimate
            assert _score_of_cigar(s1, s2, runs, ma, mi, 12, 1) >= score
            NEG = -10 ** 9
            H = np.zeros((n + 1, m + 1), dtype=np.int64)
            E = np.full((n + 1, m + 1), NEG, dtype=np.int64)
            F = np.full((n + 1, m + 1), NEG, dtype=np.int64)
            for i in range(1, n + 1):
                for j in range(1, m + 1):
                    E[i, j] = max(H[i, j - 1] - 12, E[i, j - 1] - 1)
                    F[i, j] = max(H[i - 1, j] - 12, F[i - 1, j] - 1)
                    H[i, j] = max(H[i - 1, j - 1] + (ma if s1[i - 1] == s2[j - 1] else mi), E[i, j], F[i, j])
            assert score == max(H[n, 1:].max(), H[1:, m].max())


def test_sg_align_identical_and_edge_cases():
    s = lcg(120, 42)
    score, eq, er, runs = oracle.sg_align(s, s)
    assert oracle.cigar_runs_to_string(runs) == "120=" and score == 240 and (eq, er) == (119, 119)
    # query contained in ref: leading/trailing D
    score, eq, er, runs = oracle.sg_align(s[30:90], s)
    assert oracle.cigar_runs_to_string(runs) == "30D60=30D" and score == 120
    score, eq, er, runs = oracle.sg_align(s, s[30:90])
    assert oracle.cigar_runs_to_string(runs) == "30I60=30I" and score == 120
    with pytest.raises(ValueError):
        oracle.sg_align("", "ACGT")
    # wildcard scores 0 but compares equal letter-wise
    score, _, _, runs = oracle.sg_align("ACGTNACGT", "ACGTNACGT")
    assert oracle.cigar_runs_to_string(runs) == "9=" and score == 16


def test_sg_align_pairs_matches_single():
    rng = np.random.default_rng(4)
    seqs = ["".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=int(n))) for n in rng.integers(1, 200, size=12)]
    cat = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    ia = np.array([0, 1, 2, 3, 4, 5, 11], dtype=np.int32)
    ib = np.array([6, 7, 8, 9, 10, 0, 11], dtype=np.int32)
    score, eq, er, cig, coff = oracle.sg_align_pairs(cat, off, ia, ib, 2, -2, 12, 1, threads=2)
    for p in range(len(ia)):
        s, q, r, runs = oracle.sg_align(seqs[ia[p]], seqs[ib[p]])
        assert (score[p], eq[p], er[p]) == (s, q, r)
        assert cig[coff[p]:coff[p + 1]].tolist() == runs.tolist()


def test_tiebreak_table_is_live():
    """Flipping an entry of the tie-break table changes at least one adversarial cigar, i.e. the
    table really is the single place where ties are decided."""
    tb = oracle.TieBreak(0, 0, 1)
    changed = 0
    rng = np.random.default_rng(1)
    for it in range(200):
        s1 = "".join("AC"[int(x)] for x in rng.integers(0, 2, size=int(rng.integers(2, 14))))
        s2 = "".join("AC"[int(x)] for x in rng.integers(0, 2, size=int(rng.integers(2, 14))))
        a = oracle.sg_align(s1, s2, 2, -2, 2, 1)
        b = oracle.sg_align(s1, s2, 2, -2, 2, 1, tiebreak=tb)
        assert a[0] == b[0]
        changed += int(a[1:3] != b[1:3])
    assert changed > 0


def test_simd_batch_aligner_is_bit_identical_to_the_scalar_checker():
    """The inter-pair SIMD CPU aligner (bench.py's CPU baseline) against the scalar restatement: ragged
    groups, wildcards, tie-heavy alphabets, a pair outside the int16 range (scalar fallback)."""
    rng = np.random.default_rng(12)
    seqs = []
    for i in range(90):
        alpha = ["ACGT", "AC", "ACGTN", "A", "ACGTacgt"][i % 5]
        seqs.append("".join(alpha[int(x)] for x in rng.integers(0, len(alpha), size=int(rng.integers(1, 330)))))
    big = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=16500))
    seqs += [big, big[:16400] + "ACGT"]
    cat = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    ia = rng.integers(0, 90, size=75).astype(np.int32)
    ib = rng.integers(0, 90, size=75).astype(np.int32)
    for params in [(2, -2, 12, 1), (2, -8, 12, 1), (1, -1, 1, 1), (2, -2, 0, 0)]:
        a = oracle.sg_align_pairs(cat, off, ia, ib, *params, threads=2, simd=False)
        b = oracle.sg_align_pairs(cat, off, ia, ib, *params, threads=2, simd=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), params
    ia2 = np.array([90, 3], dtype=np.int32); ib2 = np.array([91, 4], dtype=np.int32)     # 16.5 kb: int16 cannot hold it
    a = oracle.sg_align_pairs(cat, off, ia2, ib2, 2, -2, 12, 1, simd=False)
    b = oracle.sg_align_pairs(cat, off, ia2, ib2, 2, -2, 12, 1, simd=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
