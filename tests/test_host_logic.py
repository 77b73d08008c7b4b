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


def test_generate_combinations_matrix_form_equals_reference_loop():
    """bubbles.generate_combinations (one matrix product) against the reference's nested set-intersection loop
    (SimplifyGraph.py:110-119), restated here: same tuples, same order."""
    import numpy as np
    from isonform_b200 import bubbles

    def reference_loop(starts, ends, topo, combis):
        for s in starts:
            for e in ends:
                if topo[s[0]] < topo[e[0]]:
                    inter = tuple(sorted(set(s[1]).intersection(set(e[1]))))
                    if len(inter) >= 2:
                        combis.append((s[0], e[0], inter))

    rng = np.random.default_rng(7)
    for trial in range(20):
        n_reads = int(rng.integers(2, 40))
        nodes = ["n%d" % i for i in range(int(rng.integers(2, 30)))]
        perm = rng.permutation(len(nodes))
        topo = {n: int(perm[i]) for i, n in enumerate(nodes)}

        def supp():
            size = int(rng.integers(1, n_reads + 1))
            return tuple(int(x) for x in rng.choice(np.arange(1, n_reads + 1), size=size, replace=False))
        starts = [(n, supp()) for n in nodes if rng.random() < 0.6]
        ends = [(n, supp()) for n in nodes if rng.random() < 0.6]
        want, got = [], []
        reference_loop(starts, ends, topo, want)
        bubbles.generate_combinations(starts, ends, topo, got)
        assert got == want, trial
    got = []
    bubbles.generate_combinations([], [("a", (1, 2))], {"a": 0}, got)
    assert got == []


def test_host_wis_and_instances_match_python_path():
    """csrc/host_wis.cpp (isof_wis_picks / isof_span_instances: plain C++, runs without a GPU) against the numpy/Python
    restatement of main:227-272,324-337 on random batches -- same picks, same instance arrays, same skipped reads."""
    import sys as _sys
    import os as _os
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    from fake_ctx import FakeContext
    from isonform_b200 import hotpath, runner, spans, synth
    for seed, n, eps, dl in ((3, 30, 0.01, 5), (4, 40, 0.05, 5), (9, 25, 0.0, 10), (10, 12, 0.1, 3)):
        reads, _i, _t = synth.make_cluster(seed=seed, gene_len=800, n_isoforms=3, n_reads=n, eps=eps)
        rd = runner.read_batch([(a, s, "I" * len(s)) for a, s in reads])
        rd[len(rd) + 1] = ("short", "ACGTACGTAC", "I" * 10)                 # a read without any interval
        ctx = FakeContext()
        M = hotpath.get_minimizers_and_positions(rd, 31, 20, "lex", ctx)
        idx = spans.SpanIndex(rd, M, 20, 40, 80, dl, ctx)
        a = spans.batch_intervals_for_graph(idx, rd)
        b = spans.batch_intervals_for_graph_py(idx, rd)
        assert a[1] == b[1] and a[2] == b[2] and list(a[0]) == list(b[0])
        for key in a[0]:
            assert [(x, y, z, list(w)) for x, y, z, w in a[0][key]] == [(x, y, z, list(w)) for x, y, z, w in b[0][key]]
        # the front half as the runner issues it (positions as arrays, no minimizer dict) gives the same
        c = runner.batch_front_half(rd, 20, 31, 40, 80, dl, FakeContext())
        assert c[2] == a[2] and list(c[0]) == list(a[0])
        for key in a[0]:
            assert [(x, y, z, list(w)) for x, y, z, w in c[0][key]] == [(x, y, z, list(w)) for x, y, z, w in a[0][key]]


def test_seam_capture_replays_over_the_oracle_double():
    """The recorded reference seams (tests/golden/seam_capture.json) replayed through the host logic over the
    oracle-backed test double: pins the fixture and the host half on CPU; the `-m gpu` twin runs the real context."""
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import test_gpu_parity
    from fake_ctx import FakeContext
    for i in (0, 1):
        test_gpu_parity.test_reference_seams_replayed_on_gpu(FakeContext(), i)


def test_host_poa_consensus_properties():
    """csrc/host_poa.cpp: deterministic; a single sequence is its own consensus; a majority base wins a column; the
    consensus of noisy copies is far closer to the truth than any copy (30 copies at 5 % error: exact)."""
    import numpy as np
    import oracle
    from isonform_b200 import poa, synth
    assert poa.consensus(["ACGTACGT"]) == "ACGTACGT"
    assert poa.consensus(["acgtacgt"]) == "ACGTACGT"
    assert poa.consensus(["ACGTACGT", "ACGTTCGT", "ACGTACGT"]) == "ACGTACGT"
    assert poa.consensus([]) == "" and poa.consensus(["", "ACGT"]) == "ACGT"
    rng = np.random.default_rng(11)
    truth = synth.random_dna(rng, 400)
    copies = [synth.mutate(rng, truth, 0.05) for _ in range(30)]
    con = poa.consensus(copies)
    assert con == poa.consensus(list(copies))                       # deterministic
    assert oracle.edit_distance(con, truth) <= 1 < min(oracle.edit_distance(c, truth) for c in copies)


def test_edlib_shaped_host_functions_over_the_oracle_double():
    """hotpath.edlib_align / edit_distance_pairs (edlib's interface; setup.py:81 declares it, nothing calls it): host-side
    formatting and argument checks over the oracle-backed context double.  The CUDA kernel behind them is tested against
    the same oracle in tests/test_edit_distance_gpu.py."""
    import pytest
    from fake_ctx import FakeContext
    from isonform_b200 import hotpath
    ctx = FakeContext()
    assert hotpath.edlib_align("kitten", "sitting", ctx=ctx) == {"editDistance": 3, "cigar": None}
    r = hotpath.edlib_align("ACGTACGT", "ACGACGTT", task="path", ctx=ctx)
    assert r["editDistance"] == 2 and sum(int(x) for x in __import__("re").findall(r"(\d+)[=XI]", r["cigar"])) == 8
    assert hotpath.edlib_align("AAAA", "TTTTTTTT", k=3, task="path", ctx=ctx) == {"editDistance": -1, "cigar": None}
    out = hotpath.edit_distance_pairs(["ACGT", "ACGA", ""], [0, 0, 2], [1, 2, 1], task="path", ctx=ctx)
    assert [o["editDistance"] for o in out] == [1, 4, 4] and out[0]["cigar"] == "3=1X" and out[1]["cigar"] == "4I" and out[2]["cigar"] == "4D"
    with pytest.raises(ValueError):
        hotpath.edlib_align("A", "C", mode="HW", ctx=ctx)
    with pytest.raises(ValueError):
        hotpath.edit_distance_pairs(["A"], [0], [0], task="locations", ctx=ctx)
