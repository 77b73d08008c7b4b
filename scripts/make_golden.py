#!/usr/bin/env python
"""Generate tests/golden/*.json by EXECUTING the unmodified reference (/root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONHASHSEED=0 python scripts/make_golden.py

What is pinned by execution of the reference's own code:
  * get_kmer_minimizers / get_minimizers_and_positions        (main:81-110,159-171)
  * remove_read_polyA_ends                                    (main:47-67)
  * minimizers_comb_iterator, get_minimizer_combinations_database, find_most_supported_span,
    solve_WIS                                                 (main:174-263,308-338)
  * cigar_to_seq, parse_cigar_diversity, find_first_significant_match,
    parse_cigar_diversity_isoform_level_new, align_to_merge, the align_bubble_nodes length gates
                                                              (consensus.py, SimplifyGraph.py,
                                                               IsoformGeneration.py)
What is NOT pinned: parasail itself (absent, third-party).  The reference's
`consensus.parasail_alignment` is executed with a stand-in `parasail` module backed by
oracle/isof_oracle.c, so the decision goldens pin everything *around* the aligner given the
oracle's CIGARs.
"""
import hashlib
import importlib.machinery
import importlib.util
import json
import os
import sys
import types

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from isonform_b200 import synth  # noqa: E402


def install_parasail_stub():
    """A `parasail` module exposing exactly what consensus.py:55-67 touches."""
    mod = types.ModuleType("parasail")

    class _Matrix:
        def __init__(self, alphabet, match, mismatch):
            assert alphabet == "ACGT"
            self.match, self.mismatch = match, mismatch

    class _Cigar:
        def __init__(self, runs):
            self.decode = oracle.cigar_runs_to_string(runs).encode()

    class _Result:
        def __init__(self, s1, s2, gap_open, gap_ext, matrix):
            self.score, self.end_query, self.end_ref, runs = oracle.sg_align(
                s1, s2, matrix.match, matrix.mismatch, gap_open, gap_ext)
            self.saturated = False
            self.cigar = _Cigar(runs)

    mod.matrix_create = lambda alphabet, match, mismatch: _Matrix(alphabet, match, mismatch)
    mod.sg_trace_scan_16 = lambda s1, s2, o, e, m: _Result(s1, s2, o, e, m)
    mod.sg_trace_scan_32 = mod.sg_trace_scan_16
    sys.modules["parasail"] = mod


def load_reference():
    install_parasail_stub()
    loader = importlib.machinery.SourceFileLoader("isonform_ref_main", os.path.join(REF, "main"))
    spec = importlib.util.spec_from_loader("isonform_ref_main", loader)
    mod = importlib.util.module_from_spec(spec)
    _stdout = sys.stdout
    loader.exec_module(mod)
    sys.stdout = _stdout
    from modules import IsoformGeneration, SimplifyGraph, consensus
    return mod, consensus, SimplifyGraph, IsoformGeneration


def lcg(n, seed):
    x = seed
    out = []
    for _ in range(n):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        out.append("ACGT"[(x >> 16) & 3])
    return "".join(out)


def main():
    assert sys.flags.hash_randomization == 0 or os.environ.get("PYTHONHASHSEED") == "0"
    ref, consensus, SG, IG = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(20261019)

    # ------------------------------------------------------------------ hashes
    hcases = ["", "A", "ACGT", "ACGTACG", "ACGTACGT", "ACGTACGTA", "ACGTACGTACGTACGTACGT", "N" * 20,
              "acgtacgtacgtacgtacgt", lcg(31, 5), lcg(64, 6)]
    for ln in (1, 2, 3, 7, 8, 9, 15, 16, 17, 19, 20, 21, 23, 24, 25, 31, 32, 33, 40):
        hcases.append(synth.random_dna(rng, ln))
    hashes = [{"s": s, "hash": hash(s)} for s in hcases]

    # ------------------------------------------------------------------ minimizers
    mcases = []

    def add_min(name, seq, k, w):
        res = ref.get_kmer_minimizers(seq, k, w)
        for kmer, p in res:
            assert kmer == seq[p:p + k]
        mcases.append({"name": name, "seq": seq, "k": k, "w": w, "pos": [p for _, p in res]})

    s42 = lcg(200, 42)
    add_min("KAT1", s42, 20, 31)
    add_min("KAT2", s42, 9, 20)
    kat4 = "ACGTTGCA" * 10 + lcg(60, 3) + "AC" * 25
    add_min("KAT4_dup_windows", kat4, 20, 31)
    add_min("survey_k5", "ACGTTGCAAGGCTTAACCGGTTAGCATCGATCGGATCGATTTAGCGAGCTAGCTAGGGATCGATCAGCTTT", 5, 9)
    add_min("homopolymer", "A" * 120, 20, 31)
    add_min("dinuc", "AC" * 80, 20, 31)
    add_min("trinuc_k9", "ACG" * 50, 9, 20)
    add_min("with_N", s42[:60] + "NNNNN" + s42[60:140] + "acgtn" + s42[140:], 20, 31)
    for ln in (0, 1, 5, 11, 12, 19, 20, 21, 30, 31, 32, 33, 40, 50, 51, 52, 53, 60):
        add_min("short_%d" % ln, synth.random_dna(rng, ln), 20, 31)
    for ln in (3, 8, 13, 14, 15, 22):
        add_min("short_k5_%d" % ln, synth.random_dna(rng, ln), 5, 9)
    for i in range(12):
        add_min("rand_%d" % i, synth.random_dna(rng, int(rng.integers(100, 3000))), 20, 31)
    for i, (k, w) in enumerate([(20, 20), (20, 21), (15, 50), (31, 40), (32, 64), (8, 8), (13, 27), (24, 100)]):
        add_min("kw_%d_%d" % (k, w), synth.random_dna(rng, 700), k, w)
    for i in range(8):          # tie-heavy: low-complexity alphabets and periodic strings
        alpha = "AC" if i % 2 else "AAC"
        seq = "".join(alpha[int(x)] for x in rng.integers(0, len(alpha), size=400))
        add_min("lowcomplex_%d" % i, seq, 9 if i < 4 else 20, 20 if i < 4 else 31)
        unit = synth.random_dna(rng, int(rng.integers(3, 30)))
        add_min("periodic_%d" % i, (unit * 60)[:600], 20, 31)
    # KAT3: long sequence, digest only
    s7 = lcg(100000, 7)
    p7 = [p for _, p in ref.get_kmer_minimizers(s7, 20, 31)]
    kat3 = {"seed": 7, "n": 100000, "k": 20, "w": 31, "count": len(p7),
            "sha256_16": hashlib.sha256(",".join(map(str, p7)).encode()).hexdigest()[:16]}

    # ------------------------------------------------------------------ polyA
    pcases = []
    for s in ["", "A", "ACGT", "A" * 30, "ACGT" * 10 + "A" * 13, "ACGT" * 10 + "A" * 12, "ACGT" * 40 + "T" * 20 + "ACG",
              "C" * 50 + "G" * 14 + "A" * 14 + "T" * 14, synth.random_dna(rng, 300) + "A" * 25,
              "A" * 15 + synth.random_dna(rng, 150) + "A" * 15 + "C"]:
        pcases.append({"seq": s, "out": ref.remove_read_polyA_ends(s, 12, 1)})

    # ------------------------------------------------------------------ pairs / spans / WIS
    span_cases = []

    def add_span_case(name, reads, k, w, xmin, xmax, delta_len):
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            M = ref.get_minimizers_and_positions(reads, w, k, "lex")
            db = ref.get_minimizer_combinations_database(M, k, xmin, xmax, reads)
        dbl = []
        for m1 in db:
            for m2 in db[m1]:
                if len(db[m1][m2]):
                    dbl.append([m1, m2, list(db[m1][m2])])
        per_read = {}
        for r_id in sorted(reads):
            intervals = []
            for (m1, p1), spans in ref.minimizers_comb_iterator(M[r_id], k, xmin, xmax):
                if spans:       # == not_prev_corrected_spans non-empty under --exact (main:466-480)
                    ref.find_most_supported_span(r_id, m1, p1, spans, db, intervals, k, delta_len)
            wis = None
            if intervals:
                srt = sorted(intervals, key=lambda x: x[1])
                opt = ref.solve_WIS(srt)
                chosen = ref.get_intervals_to_correct(opt[::-1], srt)
                wis = [[a, b, c, list(d)] for a, b, c, d in chosen]
            # full instance arrays only for the WIS picks; a digest for the (many) candidates
            per_read[str(r_id)] = {"intervals": [[a, b, c, hashlib.sha256(d.tobytes()).hexdigest()[:12]]
                                                 for a, b, c, d in intervals], "wis": wis}
        span_cases.append({"name": name, "k": k, "w": w, "xmin": xmin, "xmax": xmax, "delta_len": delta_len,
                           "reads": {str(r): reads[r][1] for r in reads},
                           "minimizers": {str(r): [p for _, p in M[r]] for r in M},
                           "db": dbl, "per_read": per_read})

    s = s42
    kat5 = {1: ("a", s, "I"), 2: ("b", s[:90] + s[95:], "I"), 3: ("c", s, "I"), 4: ("d", s[:150] + "G" + s[150:], "I")}
    add_span_case("KAT5", kat5, 20, 31, 40, 80, 5)
    rd, _iso, _t = synth.make_cluster(seed=11, gene_len=600, n_isoforms=3, n_reads=24, eps=0.01)
    reads = {i + 1: (acc, ref.remove_read_polyA_ends(sq, 12, 1), "I" * len(sq)) for i, (acc, sq) in enumerate(rd)}
    add_span_case("synth24_e01", reads, 20, 31, 40, 80, 5)
    rd, _iso, _t = synth.make_cluster(seed=12, gene_len=500, n_isoforms=2, n_reads=16, eps=0.04)
    reads = {i + 1: (acc, ref.remove_read_polyA_ends(sq, 12, 1), "I" * len(sq)) for i, (acc, sq) in enumerate(rd)}
    add_span_case("synth16_e04_k9", reads, 9, 20, 18, 60, 10)
    # over-abundance (> 3*len(reads) occurrences) and the polyA-pair exclusion
    rep = "ACGTTGCATG" * 40
    reads = {1: ("r1", rep, ""), 2: ("r2", rep[3:] + "A" * 150, ""), 3: ("r3", "A" * 150 + rep[:200], "")}
    add_span_case("repeats_polyA", reads, 9, 20, 18, 60, 5)

    # ------------------------------------------------------------------ alignment decisions
    dec = []

    def add_pair(name, a, b):
        rec = {"name": name, "s1": a, "s2": b}
        for tag, (mm, mis) in (("bubble", (2, -8)), ("iso", (2, -2))):
            a1, a2, cig, tup, score = consensus.parasail_alignment(a, b, mm, mis, 12, 1)
            rec[tag] = {"cigar": cig, "score": score, "a1_sha": hashlib.sha256(a1.encode()).hexdigest()[:12],
                        "a2_sha": hashlib.sha256(a2.encode()).hexdigest()[:12]}
            if tag == "bubble":
                rec[tag]["pop_d5"] = SG.parse_cigar_diversity(tup, 0.20, 5)
                rec[tag]["pop_d10"] = SG.parse_cigar_diversity(tup, 0.20, 10)
            else:
                rec[tag]["start_match"] = IG.find_first_significant_match(a1, a2, 30, 0.7)
                rec[tag]["end_match"] = IG.find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7)
                for dl, d in ((5, 0.15), (10, 0.15), (5, 0.1)):
                    rec[tag]["merge_dl%d_d%s" % (dl, d)] = IG.align_to_merge(a, b, d, dl, 30, 50)
        dec.append(rec)

    for i in range(40):
        ln = int(rng.integers(12, 400))
        x = synth.random_dna(rng, ln)
        y = synth.mutate(rng, x, [0.0, 0.01, 0.05, 0.15][i % 4])
        if i % 5 == 0 and ln > 60:      # exon-sized indel
            c = int(rng.integers(10, ln - 40)); y = y[:c] + y[c + int(rng.integers(6, 40)):]
        if i % 7 == 0:
            y = synth.random_dna(rng, int(rng.integers(0, 60))) + y
        if i % 11 == 0:
            x = x + synth.random_dna(rng, int(rng.integers(0, 60)))
        if len(y) == 0:
            y = "A"
        add_pair("rand_%d" % i, x, y)
    for i in range(12):
        ln = int(rng.integers(300, 1500))
        x = synth.random_dna(rng, ln)
        y = synth.mutate(rng, x, [0.005, 0.02, 0.05][i % 3])
        if i % 2:
            y = y[int(rng.integers(0, 80)):len(y) - int(rng.integers(0, 60))]
        if i % 4 == 3:
            c = len(y) // 2; y = y[:c] + y[c + 7:]
        add_pair("iso_%d" % i, y, x) if i % 3 == 0 else add_pair("iso_%d" % i, x, y)
    add_pair("identical", s42, s42)
    add_pair("unrelated", synth.random_dna(rng, 150), synth.random_dna(rng, 170))
    add_pair("wild_N", s42[:50] + "NNNN" + s42[54:120], s42[:120])
    add_pair("placeholder_X", "X" * 40, s42[:45])
    add_pair("lower", s42[:80].lower(), s42[:80])
    add_pair("single_vs_long", "A", s42[:30])
    add_pair("long_vs_single", s42[:30], "C")
    add_pair("homopolymer", "A" * 60, "A" * 45)
    add_pair("repeat_shift", "ACGT" * 20, "CGTA" * 20)

    # align_bubble_nodes length gates (SimplifyGraph.py:630-664): delta = 0.20
    gates = []
    for (l1, l2, dl) in [(3, 4, 5), (5, 5, 5), (5, 9, 5), (5, 10, 5), (4, 12, 5), (6, 8, 5), (100, 79, 5), (100, 80, 5),
                         (100, 81, 5), (10, 12, 10), (10, 21, 10), (11, 13, 10), (50, 63, 5)]:
        a, b = synth.random_dna(rng, l1), synth.random_dna(rng, l2)
        longer, shorter = max(l1, l2), min(l1, l2)
        if shorter > dl and longer > dl:
            if (longer - shorter) / longer < 0.20:
                _a1, _a2, _c, tup, _s = consensus.parasail_alignment(a, b, 2, -8, 12, 1)
                res, aligned = SG.parse_cigar_diversity(tup, 0.20, dl), True
            else:
                res, aligned = False, False
        elif shorter < dl and longer < dl:
            res, aligned = True, False
        else:
            res, aligned = (longer - shorter) < dl, False
        gates.append({"s1": a, "s2": b, "delta_len": dl, "pop": res, "aligned": aligned})

    def dump(name, obj):
        with open(os.path.join(out_dir, name), "w") as fh:
            json.dump(obj, fh, separators=(",", ":"))
        print(name, os.path.getsize(os.path.join(out_dir, name)), "bytes")

    dump("hashes.json", {"python": sys.version.split()[0], "algorithm": sys.hash_info.algorithm, "cases": hashes})
    dump("minimizers.json", {"cases": mcases, "kat3": kat3})
    dump("polya.json", pcases)
    dump("spans.json", span_cases)
    dump("decisions.json", {"pairs": dec, "gates": gates,
                            "note": "cigars come from oracle/isof_oracle.c (parasail absent: PARITY UNPINNED); "
                                    "decisions are the reference's own functions applied to them"})


if __name__ == "__main__":
    main()
