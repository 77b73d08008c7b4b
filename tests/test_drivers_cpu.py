"""CPU: the host drivers (speculation + sequential replay, span index, patching) on a test double of
the GPU context (tests/fake_ctx.py, oracle-backed).  Where /root/reference exists (build container)
the reference's own functions are executed as the expected value."""
import copy
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

import oracle
from fake_ctx import FakeContext
from isonform_b200 import hotpath, merge, polya, runner, spans, synth

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "modules")), reason="reference tree not present")


def _consensus_set(seed, n, eps, frag=True):
    rng = np.random.default_rng(seed)
    iso = [synth.random_dna(rng, 420), synth.random_dna(rng, 380)]
    iso.append(iso[0][:200] + iso[0][260:])                 # exon skip of isoform 0
    out = []
    for i in range(n):
        s = synth.mutate(rng, iso[int(rng.integers(0, 3))], eps)
        if frag and i % 3 == 0:
            s = s[int(rng.integers(0, 25)):len(s) - int(rng.integers(0, 20))]
        out.append(s)
    return out


def _sequential_merge(seqs_by_id, support, delta, delta_len, d3, d5, regen):
    """The reference's loop (IsoformGeneration.py:464-507) written out plainly on top of the oracle."""
    cbs = copy.deepcopy(support)
    alt = sorted(seqs_by_id.items(), key=lambda x: len(x[1]))
    new = {}
    n_calls = 0
    for i, (id1, orig1) in enumerate(alt[:-1]):
        seq1 = new.get(id1, orig1)
        for id2, seq2 in alt[i + 1:]:
            c1, c2 = (seq2, seq1) if len(seq2) < len(seq1) else (seq1, seq2)
            n_calls += 1
            if oracle.align_to_merge(c1, c2, delta, delta_len, d3, d5):
                if len(cbs[id2]) > 50:
                    cbs[id2] = cbs[id2] + cbs[id1]; cbs.pop(id1)
                    new[id2] = seqs_by_id[id2]
                else:
                    new[id2] = regen(id1, id2, cbs)
                    cbs[id2] = cbs[id2] + cbs[id1]; cbs.pop(id1)
                break
    for id_ in cbs:
        new.setdefault(id_, seqs_by_id[id_])
    return new, cbs, n_calls


@pytest.mark.parametrize("seed,eps,block", [(1, 0.004, 8192), (2, 0.004, 7), (3, 0.02, 30), (4, 0.0, 5)])
def test_merge_consensuses_speculative_equals_sequential(seed, eps, block, monkeypatch):
    monkeypatch.setattr(merge, "_MIN_PAIRS_PER_CALL", block)
    seqs = _consensus_set(seed, 26, eps)
    seqs_by_id = {100 + i: s for i, s in enumerate(seqs)}
    support = {100 + i: [100 + i] * (60 if i % 5 == 0 else 1) for i in range(len(seqs))}

    def regen(id1, id2, cbs):                       # deterministic stand-in for the spoa re-consensus
        a, b = seqs_by_id[id1], seqs_by_id[id2]
        return b[:len(b) // 2] + a[len(a) // 2:] if len(cbs[id2]) % 2 else b

    want_new, want_cbs, n_calls = _sequential_merge(seqs_by_id, support, 0.15, 5, 30, 50, regen)
    assert len(want_cbs) < len(seqs), "the test set must actually merge something"

    def gen_all(all_c, alt, cbs, reads, work_dir, max_spoa):
        for id_ in cbs:
            all_c[id_] = seqs_by_id[id_]
            alt.append((id_, seqs_by_id[id_]))

    cbs = copy.deepcopy(support)
    stats = {}
    ctx = FakeContext()
    got = merge.merge_consensuses(cbs, None, 0.15, 5, None, 200, 30, 50, generate_all_consensuses=gen_all,
                                  generate_new_full_consensus=lambda i1, i2, reads, c, wd, ms: regen(i1, i2, c), ctx=ctx,
                                  stats=stats)
    assert got == want_new and cbs == want_cbs
    assert stats["pairs_used"] == n_calls and stats["pairs_aligned"] >= n_calls


class _Info:
    def __init__(self, sequence, reads):
        self.sequence, self.reads, self.merged = sequence, reads, False


def _sequential_cross_batch(all_infos, delta, delta_len, d3, d5, regen):
    """batch_merging_parallel.py:160-212 written out on top of the oracle."""
    lst = sorted(all_infos.items(), key=lambda x: x[0], reverse=True)
    for b_i, (bid, d1) in enumerate(lst[:-1]):
        l1 = sorted(d1.items(), key=lambda x: len(x[1].sequence))
        for bid2, d2 in lst[b_i + 1:]:
            l2 = sorted(d2.items(), key=lambda x: len(x[1].sequence))
            if not bid2 <= bid:
                for id_, infos in l1:
                    if not infos.merged:
                        for id2, infos2 in l2:
                            if len(infos2.sequence) >= len(infos.sequence):
                                long_, short_ = infos2, infos
                            else:
                                long_, short_ = infos, infos2
                            if long_.merged:
                                continue
                            if not short_.merged:
                                if oracle.align_to_merge(long_.sequence, short_.sequence, delta, delta_len, d3, d5):
                                    if len(long_.reads) > 50:
                                        long_.reads = long_.reads + short_.reads
                                        short_.merged = True
                                    else:
                                        long_.sequence = regen(long_, short_)
                                        long_.reads = long_.reads + short_.reads
                                        short_.merged = True


def test_cross_batch_merging_equals_sequential_and_is_unreachable_in_the_reference():
    """batch_merging_parallel.py:162-167 sorts the batches by id DESCENDING and then only compares a
    batch with later list entries whose id is LARGER -- which never happens for a total order, so the
    reference performs no cross-batch alignment at all (call site 3 of SURVEY 8a is dead).  The
    mirror must reproduce exactly that: same (unchanged) state and zero alignment launches."""
    seqs = _consensus_set(7, 18, 0.004)

    def build():
        return {2: {i: _Info(seqs[i], [i] * (60 if i % 4 == 0 else 1)) for i in range(0, 6)},
                1: {i: _Info(seqs[i], [i]) for i in range(6, 12)},
                0: {i: _Info(seqs[i], [i]) for i in range(12, 18)}}
    regen = lambda long_, short_: long_.sequence[:-3]
    want = build()
    _sequential_cross_batch(want, 0.1, 5, 30, 50, regen)
    got = build()
    ctx = FakeContext()
    merge.actual_merging_process(got, 0.1, 5, 30, 50, 200, None, None,
                                 generate_consensus_path=lambda wd, m1, m2, abs_, ms: "X", ctx=ctx)
    flat = lambda d: {(b, i): (inf.sequence, inf.reads, inf.merged) for b, dd in d.items() for i, inf in dd.items()}
    assert flat(got) == flat(want) == flat(build())
    assert ctx.calls["align"] == 0


def test_cross_batch_replay_logic_when_made_reachable(monkeypatch):
    """The replay itself (memo prefetch + loop) is still exercised: feed batch ids through a key
    wrapper whose ordering makes the comparison reachable, on both sides."""
    class K:                                       # sorts descending by v, but "<=" is inverted
        def __init__(self, v): self.v = v
        def __lt__(self, o): return self.v < o.v
        def __le__(self, o): return self.v > o.v
        def __hash__(self): return hash(self.v)
        def __eq__(self, o): return self.v == o.v
    seqs = _consensus_set(7, 18, 0.004)
    keys = [K(2), K(1), K(0)]

    def build():
        return {keys[0]: {i: _Info(seqs[i], [i] * (60 if i % 4 == 0 else 1)) for i in range(0, 6)},
                keys[1]: {i: _Info(seqs[i], [i]) for i in range(6, 12)},
                keys[2]: {i: _Info(seqs[i], [i]) for i in range(12, 18)}}
    regen = lambda long_, short_: long_.sequence[:-3] if len(long_.reads) % 2 else long_.sequence
    want = build()
    _sequential_cross_batch(want, 0.1, 5, 30, 50, regen)
    got = build()

    def gcp(work_dir, mappings1, mappings2, all_batch_sequences, max_spoa):
        owner = [inf for d in got.values() for inf in d.values() if inf.reads is mappings1][0]
        return regen(owner, None)
    ctx = FakeContext()
    memo = merge.actual_merging_process(got, 0.1, 5, 30, 50, 200, None, None, generate_consensus_path=gcp, ctx=ctx)
    flat = lambda d: {(b.v, i): (inf.sequence, inf.reads, inf.merged) for b, dd in d.items() for i, inf in dd.items()}
    assert flat(got) == flat(want)
    assert any(m for _, _, m in flat(want).values()), "the test set must actually merge something"
    assert memo.hits > 0 and ctx.calls["align"] < memo.hits


def test_alignment_memo_prefetch_and_fallback():
    ctx = FakeContext()
    memo = merge.AlignmentMemo(ctx)
    seqs = _consensus_set(5, 6, 0.02, frag=False)
    pairs = [(seqs[i], seqs[j]) for i in range(6) for j in range(i + 1, 6)]
    assert memo.prefetch_alignments(pairs, 2, -8, 12, 1) == len(pairs)
    assert memo.prefetch_alignments(pairs, 2, -8, 12, 1) == 0
    calls = ctx.calls["align"]
    for a, b in pairs:
        assert memo.parasail_alignment(a, b, 2, -8, 12, 1) == oracle.parasail_alignment(a, b, 2, -8, 12, 1)
    assert ctx.calls["align"] == calls and memo.hits == len(pairs)
    # argument order is part of the key; a miss falls back to a synchronous single-pair call
    assert memo.parasail_alignment(seqs[1], seqs[0], 2, -8, 12, 1) == oracle.parasail_alignment(seqs[1], seqs[0], 2, -8, 12, 1)
    assert memo.misses == 1
    with pytest.raises(ValueError):
        memo.parasail_alignment("", "ACGT")
    assert memo.align_to_merge(seqs[0], seqs[1], 0.15, 5, 30, 50) == oracle.align_to_merge(seqs[0], seqs[1], 0.15, 5, 30, 50)


def test_front_half_runner_matches_oracle_chain():
    rd, _iso, _t = synth.make_cluster(seed=33, gene_len=700, n_isoforms=3, n_reads=20, eps=0.02)
    reads = runner.read_batch([(acc, s, "I" * len(s)) for acc, s in rd])
    ctx = FakeContext()
    graph_iv, new_reads, skipped, index = runner.batch_front_half(reads, 20, 31, 14, 80, 5, ctx)   # xmin raised to 40
    M = oracle.get_minimizers_and_positions(reads, 31, 20)
    db = oracle.get_minimizer_combinations_database(M, 20, 40, 80, reads)
    gid = 1
    for r_id in sorted(reads):
        iv = oracle.pyside.read_intervals(r_id, M[r_id], db, 20, 40, 80, 5)
        picks = []
        if iv:
            srt = sorted(iv, key=lambda x: x[1])
            picks = [srt[j] for j in oracle.solve_WIS(srt)[::-1]]
        if picks:
            assert [(a, b, c, list(d)) for a, b, c, d in graph_iv[gid]] == [(a, b, c, list(d)) for a, b, c, d in picks]
            gid += 1
        else:
            assert r_id in skipped
    assert ctx.calls["minimizers"] == 1 and ctx.calls["spans"] == 1


# ------------------------------------------------------------------------------------------------
# against the executed reference (build container only)
# ------------------------------------------------------------------------------------------------
def _load_reference():
    sys.path.insert(0, REF)
    stub = types.ModuleType("parasail")

    class _R:
        def __init__(self, s1, s2, o, e, m):
            self.score, self.end_query, self.end_ref, runs = oracle.sg_align(s1, s2, m[0], m[1], o, e)
            self.saturated = False
            self.cigar = types.SimpleNamespace(decode=oracle.cigar_runs_to_string(runs).encode())

    stub.matrix_create = lambda a, m, x: (m, x)
    stub.sg_trace_scan_16 = lambda s1, s2, o, e, m: _R(s1, s2, o, e, m)
    stub.sg_trace_scan_32 = stub.sg_trace_scan_16
    sys.modules["parasail"] = stub
    loader = importlib.machinery.SourceFileLoader("isonform_ref_main_t", os.path.join(REF, "main"))
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader("isonform_ref_main_t", loader))
    out = sys.stdout
    loader.exec_module(mod)
    sys.stdout = out
    return mod


@needs_ref
def test_patch_install_rebinds_reference_seams_and_reference_merge_agrees(tmp_path, capsys):
    import subprocess
    ref_main = _load_reference()
    from modules import IsoformGeneration as IG, batch_merging_parallel as BMP, consensus
    seqs = _consensus_set(11, 16, 0.004)
    reads = {i: ("acc%d" % i, s, "") for i, s in enumerate(seqs)}
    support = {i: [i] for i in range(len(seqs))}
    # deterministic spoa stand-in, identical on both sides: the longest member read
    def fake_spoa(ref_file, outfile):
        recs = [l.strip() for l in open(ref_file) if not l.startswith(">")]
        return max(recs, key=len)
    saved = (IG.run_spoa, consensus.parasail_alignment, IG.align_to_merge, IG.merge_consensuses, BMP.actual_merging_process)
    IG.run_spoa = fake_spoa
    try:
        want_cbs = copy.deepcopy(support)
        want = IG.merge_consensuses(want_cbs, str(tmp_path), 0.15, 5, reads, 200, 30, 50)      # the reference itself
        from isonform_b200 import patch
        g = {}
        memo = patch.install(g, ctx=FakeContext())
        assert consensus.parasail_alignment == memo.parasail_alignment and IG.align_to_merge == memo.align_to_merge
        assert set(g) == {"get_minimizers_and_positions", "get_kmer_minimizers", "get_minimizer_combinations_database",
                          "find_most_supported_span"}
        got_cbs = copy.deepcopy(support)
        got = IG.merge_consensuses(got_cbs, str(tmp_path), 0.15, 5, reads, 200, 30, 50)        # patched
        assert got == want and got_cbs == want_cbs and len(got_cbs) < len(seqs)
        # front-half seams with the reference's own per-read loop
        rd, _i, _t = synth.make_cluster(seed=35, gene_len=600, n_isoforms=2, n_reads=12, eps=0.01)
        rds = {i + 1: (a, polya.remove_read_polyA_ends(s, 12, 1), "") for i, (a, s) in enumerate(rd)}
        M_ref = ref_main.get_minimizers_and_positions(rds, 31, 20, "lex") if os.environ.get("PYTHONHASHSEED") == "0" else None
        M = g["get_minimizers_and_positions"](rds, 31, 20, "lex")
        if M_ref is not None:
            assert M == M_ref
        db = g["get_minimizer_combinations_database"](M, 20, 40, 80, rds)
        db_ref = ref_main.get_minimizer_combinations_database(M, 20, 40, 80, rds)
        for r_id in sorted(rds):
            a, b = [], []
            for (m1, p1), sp in ref_main.minimizers_comb_iterator(M[r_id], 20, 40, 80):
                if sp:
                    g["find_most_supported_span"](r_id, m1, p1, sp, db, a, 20, 5)
                    ref_main.find_most_supported_span(r_id, m1, p1, sp, db_ref, b, 20, 5)
            assert [(w, x, y, list(z)) for w, x, y, z in a] == [(w, x, y, list(z)) for w, x, y, z in b]
        assert [[m1, m2, list(db[m1][m2])] for m1 in db for m2 in db[m1]] == \
               [[m1, m2, list(db_ref[m1][m2])] for m1 in db_ref for m2 in db_ref[m1] if len(db_ref[m1][m2])]
    finally:
        IG.run_spoa, consensus.parasail_alignment, IG.align_to_merge, IG.merge_consensuses, BMP.actual_merging_process = saved


def test_super_batched_front_half_equals_per_batch():
    """runner.batches_front_half: one minimizer call + one span call for several clusters, same result as
    batch by batch (incl. an empty batch and a batch whose reads yield no interval)."""
    batches = []
    for seed, n in ((41, 14), (42, 9), (43, 11)):
        rd, _iso, _t = synth.make_cluster(seed=seed, gene_len=600, n_isoforms=2, n_reads=n, eps=0.02)
        batches.append(runner.read_batch([(acc, s, "") for acc, s in rd]))
    batches.insert(1, runner.read_batch([("lonely", synth.random_dna(np.random.default_rng(3), 300), "")]))
    ctx = FakeContext()
    multi = runner.batches_front_half(batches, 20, 31, 40, 80, 5, ctx)
    assert ctx.calls["minimizers"] == 1 and ctx.calls["spans"] == 1
    for reads, (giv, new_reads, skipped, idx) in zip(batches, multi):
        giv1, new1, skip1, idx1 = runner.batch_front_half(reads, 20, 31, 40, 80, 5, FakeContext())
        assert skipped == skip1 and new_reads == new1
        assert {g: [(a, b, c, list(d)) for a, b, c, d in v] for g, v in giv.items()} == \
               {g: [(a, b, c, list(d)) for a, b, c, d in v] for g, v in giv1.items()}
        gdb, gdb1 = idx.database(), idx1.database()
        assert [[m1, m2, list(gdb[m1][m2])] for m1 in gdb for m2 in gdb[m1]] == \
               [[m1, m2, list(gdb1[m1][m2])] for m1 in gdb1 for m2 in gdb1[m1]]


def test_merge_consensuses_block_grows_while_nothing_merges(monkeypatch):
    """Rows that never merge consume every speculated pair, so the speculation block grows x4 per launch (a C2 batch
    is issued in a handful of launches, not one per 8192 pairs); a merge resets it."""
    calls = []

    def no_merge(seqs, pa, pb, delta, delta_len, d3, d5, ctx=None):
        calls.append(len(pa))
        return np.zeros(len(pa), dtype=bool)

    monkeypatch.setattr(merge.hotpath, "align_to_merge_pairs", no_merge)
    monkeypatch.setattr(merge, "_MIN_PAIRS_PER_CALL", 50)
    rng = np.random.default_rng(9)
    seqs_by_id = {i: synth.random_dna(rng, 100 + i) for i in range(120)}
    cbs = {i: [i] for i in seqs_by_id}

    def gen_all(all_c, alt, cbs_, reads, work_dir, max_spoa):
        for id_ in cbs_:
            all_c[id_] = seqs_by_id[id_]
            alt.append((id_, seqs_by_id[id_]))

    stats = {}
    out = merge.merge_consensuses(cbs, None, 0.15, 5, None, 200, 30, 50, generate_all_consensuses=gen_all,
                                  generate_new_full_consensus=None, stats=stats)
    total = 120 * 119 // 2
    assert sum(calls) == total == stats["pairs_aligned"] == stats["pairs_used"]
    assert len(calls) <= 6 and calls == sorted(calls[:-1]) + calls[-1:]        # 50-ish, x4, x4, ... then the remainder
    assert out == seqs_by_id and len(cbs) == 120
