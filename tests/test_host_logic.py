"""CPU: host-side pieces of the product (no GPU): polyA trimming, CIGAR parsing, float decisions
on integer features, synthetic data determinism."""
import numpy as np

import oracle
from conftest import load_golden
from isonform_b200 import hotpath, polya, synth


def test_polya_matches_reference_golden():
    for c in load_golden("polya.json"):
        assert polya.remove_read_polyA_ends(c["seq"], 12, 1) == c["out"]
    rng = np.random.default_rng(0)
    for _ in range(200):
        s = "".join("ACGT"[int(x)] for x in rng.choice(4, size=int(rng.integers(0, 260)), p=[0.55, 0.1, 0.1, 0.25]))
        assert polya.remove_read_polyA_ends(s, 3, 1) == oracle.remove_read_polyA_ends(s, 3, 1)


def test_cigar_to_seq_matches_golden_alignments():
    import hashlib
    g = load_golden("decisions.json")
    for rec in g["pairs"]:
        a1, a2, tup = hotpath.cigar_to_seq(rec["iso"]["cigar"], rec["s1"], rec["s2"])
        assert hashlib.sha256(a1.encode()).hexdigest()[:12] == rec["iso"]["a1_sha"]
        assert hashlib.sha256(a2.encode()).hexdigest()[:12] == rec["iso"]["a2_sha"]
        assert hotpath.find_first_significant_match(a1, a2, 30, 0.7) == rec["iso"]["start_match"]
        assert hotpath.find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7) == rec["iso"]["end_match"]
        assert hotpath.parse_cigar_diversity(tup, 0.2, 5) == oracle.parse_cigar_diversity(tup, 0.2, 5)


def _features_from_oracle(s1, s2, params, delta_len):
    a1, a2, _cig, tup, _score = oracle.parasail_alignment(s1, s2, *params)
    alen = sum(l for l, _ in tup)
    sm = oracle.find_first_significant_match(a1, a2, 30, 0.7)
    em = oracle.find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7)
    nm = [l for l, o in tup if o != "="]
    f = [alen, len(tup), sum(nm), max(nm + [0]), sm, em, 0, 0, 0, 0, 0, 0]
    if sm >= 0 and em >= 0:
        f[6:12] = oracle.isoform_features(tup, delta_len, sm, alen - em)[:6]
    return f


def test_float_decisions_on_integer_features_match_reference_golden():
    """The numpy float64 evaluation of the reference's comparisons (hotpath._merge_from_features /
    _pop_from_features) reproduces the reference's own decisions on the golden pairs."""
    g = load_golden("decisions.json")
    for dl, d in ((5, 0.15), (10, 0.15), (5, 0.1)):
        feats = np.array([_features_from_oracle(r["s1"], r["s2"], (2, -2, 12, 1), dl) for r in g["pairs"]], dtype=np.int32)
        got = hotpath._merge_from_features(feats, d, dl, 30, 50)
        assert got.tolist() == [r["iso"]["merge_dl%d_d%s" % (dl, d)] for r in g["pairs"]]
    for dl in (5, 10):
        feats = np.array([_features_from_oracle(r["s1"], r["s2"], (2, -8, 12, 1), dl) for r in g["pairs"]], dtype=np.int32)
        got = hotpath._pop_from_features(feats, 0.20, dl)
        assert got.tolist() == [r["bubble"]["pop_d%d" % dl] for r in g["pairs"]]


def test_synth_is_deterministic_and_shaped():
    r1, iso1, t1 = synth.make_config("C1")
    r2, iso2, t2 = synth.make_config("C1")
    assert r1 == r2 and iso1 == iso2 and t1 == t2
    assert len(r1) == 200 and len(iso1) == 3
    assert all(set(s) <= set("ACGT") for _, s in r1)
    r3, _, _ = synth.make_config("C1", seed=99)
    assert r3 != r1


def test_wis_mirror_matches_oracle_on_golden_intervals():
    from isonform_b200 import spans
    for case in load_golden("spans.json"):
        reads = {int(r): ("acc", s, "") for r, s in case["reads"].items()}
        M = oracle.get_minimizers_and_positions(reads, case["w"], case["k"], "lex")
        db = oracle.get_minimizer_combinations_database(M, case["k"], case["xmin"], case["xmax"], reads)
        for r_id in sorted(reads):
            iv = oracle.pyside.read_intervals(r_id, M[r_id], db, case["k"], case["xmin"], case["xmax"], case["delta_len"])
            assert list(spans.minimizers_comb_iterator(M[r_id], case["k"], case["xmin"], case["xmax"])) == \
                list(oracle.minimizers_comb_iterator(M[r_id], case["k"], case["xmin"], case["xmax"]))
            if iv:
                srt = sorted(iv, key=lambda x: x[1])
                assert spans.solve_WIS(srt) == oracle.solve_WIS(srt)
