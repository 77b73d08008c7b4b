#!/usr/bin/env python
"""Capture the hot-path seams of the UNMODIFIED reference `main()` on synthetic batches -> tests/golden/seam_capture.json.

Build container only (the reference does not travel to the GPU box):

    PYTHONHASHSEED=0 python scripts/make_seam_capture.py

The reference's per-batch entry point (main:341) is executed as is, with the oracle standing in for the absent
`parasail` module and a deterministic stand-in for the absent `spoa` binary, while thin recorders sit on the seams
the B200 library replaces:

    get_minimizers_and_positions (main:159)            -> minimizer positions per read
    GraphGeneration.generateGraphfromIntervals (:251)  -> its inputs = the output of the whole front half
                                                          (minimizer pairs, span search, WIS: main:411-514)
    consensus.parasail_alignment (consensus.py:55)     -> every call of bubble popping and isoform merging
    IsoformGeneration.align_to_merge (:380)            -> arguments and decision
    IsoformGeneration.merge_consensuses (:464)         -> inputs, consensus strings it generated, outputs

tests/test_gpu_parity.py::test_reference_seams_replayed_on_gpu feeds the recorded inputs through the real GPU
context and requires the recorded outputs: since the reference's control flow is a deterministic function of
what these seams return, equal seam outputs on the GPU are equal output files.
"""
import copy
import importlib.machinery
import importlib.util
import json
import os
import sys
import tempfile
import types

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))      # oracle-backed `parasail`
sys.path.insert(0, REF)

from isonform_b200 import synth  # noqa: E402


class Pool:
    """String pool: the capture stores every sequence once."""

    def __init__(self):
        self.index, self.items = {}, []

    def add(self, s):
        i = self.index.get(s)
        if i is None:
            i = self.index[s] = len(self.items)
            self.items.append(s)
        return i


def capture(seed, n_reads, eps, gene_len):
    loader = importlib.machinery.SourceFileLoader("isonform_ref_main_cap", os.path.join(REF, "main"))
    ref = importlib.util.module_from_spec(importlib.util.spec_from_loader("isonform_ref_main_cap", loader))
    loader.exec_module(ref)
    from modules import GraphGeneration, IsoformGeneration, consensus

    def fake_spoa(reads_file, out_file):
        seqs = [l.strip() for l in open(reads_file) if not l.startswith(">")]
        seqs.sort(key=lambda s: (len(s), s))
        return seqs[len(seqs) // 2]
    IsoformGeneration.run_spoa = fake_spoa

    pool = Pool()
    cap = {"params": dict(k=20, w=31, xmin=40, xmax=80, delta_len=5, delta=0.15, delta_iso_len_3=30, delta_iso_len_5=50,
                          max_seqs_to_spoa=200), "seed": seed, "n_reads": n_reads, "eps": eps,
           "alignments": [], "merge_decisions": []}

    orig_min = ref.get_minimizers_and_positions

    def rec_min(reads, w, k, hash_fcn):
        out = orig_min(reads, w, k, hash_fcn)
        cap["reads"] = [[r, reads[r][0], pool.add(reads[r][1])] for r in reads]
        cap["minimizers"] = {str(r): [p for _m, p in out[r]] for r in out}
        return out
    ref.get_minimizers_and_positions = rec_min

    orig_graph = GraphGeneration.generateGraphfromIntervals

    def rec_graph(all_intervals_for_graph, k, delta_len, read_len_dict, new_all_reads):
        cap["intervals_for_graph"] = {str(g): [[a, b, c, list(arr)] for a, b, c, arr in v]
                                      for g, v in all_intervals_for_graph.items()}
        cap["new_all_reads"] = {str(g): [v[0], pool.add(v[1])] for g, v in new_all_reads.items()}
        return orig_graph(all_intervals_for_graph, k, delta_len, read_len_dict, new_all_reads)
    GraphGeneration.generateGraphfromIntervals = rec_graph

    orig_align = consensus.parasail_alignment

    def rec_align(s1, s2, match_score=2, mismatch_penalty=-2, opening_penalty=12, gap_ext=1):
        out = orig_align(s1, s2, match_score, mismatch_penalty, opening_penalty, gap_ext)
        cap["alignments"].append([pool.add(s1), pool.add(s2), match_score, mismatch_penalty, opening_penalty, gap_ext,
                                  out[2], out[4]])
        return out
    consensus.parasail_alignment = rec_align

    orig_atm = IsoformGeneration.align_to_merge

    def rec_atm(c1, c2, delta, delta_len, d3, d5):
        out = orig_atm(c1, c2, delta, delta_len, d3, d5)
        cap["merge_decisions"].append([pool.add(c1), pool.add(c2), delta, delta_len, d3, d5, bool(out)])
        return out
    IsoformGeneration.align_to_merge = rec_atm

    orig_gen_all = IsoformGeneration.generate_all_consensuses
    orig_regen = IsoformGeneration.generate_new_full_consensus
    orig_mc = IsoformGeneration.merge_consensuses

    def rec_gen_all(all_consensuses, alternative, curr_best_seqs, reads, work_dir, max_seqs_to_spoa):
        orig_gen_all(all_consensuses, alternative, curr_best_seqs, reads, work_dir, max_seqs_to_spoa)
        cap["merge"]["all_consensuses"] = [[i, pool.add(s)] for i, s in alternative]

    def rec_regen(id1, id2, reads, curr_best_seqs, work_dir, max_seqs_to_spoa):
        out = orig_regen(id1, id2, reads, curr_best_seqs, work_dir, max_seqs_to_spoa)
        cap["merge"]["regenerated"].append([id1, id2, pool.add(out)])
        return out

    def rec_mc(curr_best_seqs, work_dir, delta, delta_len, reads, max_seqs_to_spoa, d3, d5):
        cap["merge"] = {"curr_best_seqs_in": {str(i): list(v) for i, v in curr_best_seqs.items()}, "delta": delta,
                        "delta_len": delta_len, "regenerated": [],
                        "reads": {str(r): [v[0], pool.add(v[1])] for r, v in reads.items()}}
        out = orig_mc(curr_best_seqs, work_dir, delta, delta_len, reads, max_seqs_to_spoa, d3, d5)
        cap["merge"]["curr_best_seqs_out"] = {str(i): list(v) for i, v in curr_best_seqs.items()}
        cap["merge"]["new_consensuses"] = {str(i): pool.add(s) for i, s in out.items()}
        return out
    IsoformGeneration.generate_all_consensuses = rec_gen_all
    IsoformGeneration.generate_new_full_consensus = rec_regen
    IsoformGeneration.merge_consensuses = rec_mc

    originals = [(GraphGeneration, "generateGraphfromIntervals", orig_graph), (consensus, "parasail_alignment", orig_align),
                 (IsoformGeneration, "align_to_merge", orig_atm), (IsoformGeneration, "generate_all_consensuses", orig_gen_all),
                 (IsoformGeneration, "generate_new_full_consensus", orig_regen), (IsoformGeneration, "merge_consensuses", orig_mc)]
    try:
        return _run(ref, cap, pool, seed, n_reads, eps, gene_len)
    finally:
        for mod, name, fn in originals:            # `modules.*` are shared between captures: unhook the recorders
            setattr(mod, name, fn)


def _run(ref, cap, pool, seed, n_reads, eps, gene_len):
    reads, _iso, _t = synth.make_cluster(seed=seed, gene_len=gene_len, n_isoforms=3, n_reads=n_reads, eps=eps)
    with tempfile.TemporaryDirectory() as tmp:
        fastq = os.path.join(tmp, "7_0.fastq")
        synth.write_fastq(fastq, reads)
        args = types.SimpleNamespace(fastq=fastq, outfolder=os.path.join(tmp, "out"), k=20, w=31, xmin=40, xmax=80, T=0.1,
                                     exact=True, disable_numpy=False, delta_len=5, max_seqs_to_spoa=200, max_seqs=1000,
                                     exact_instance_limit=0, set_w_dynamically=False, verbose=False, compression=False,
                                     delta_iso_len_3=30, delta_iso_len_5=50, parallel=True, slow=False)
        os.makedirs(args.outfolder)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            ref.main(args)
        finally:
            os.chdir(cwd)
    cap["pool"] = pool.items
    return cap


def main():
    caps = [capture(1, 36, 0.01, 700), capture(5, 28, 0.03, 650)]
    out = os.path.join(ROOT, "tests", "golden", "seam_capture.json")
    with open(out, "w") as fh:
        json.dump({"generator": "scripts/make_seam_capture.py (reference v0.3.9 executed, PYTHONHASHSEED=0, CPython %d.%d; "
                                "parasail = oracle stand-in, spoa = median-sequence stand-in)" % sys.version_info[:2],
                   "captures": caps}, fh, separators=(",", ":"))
    for c in caps:
        print("seed %d: %d reads, %d alignments, %d merge decisions, %d graph reads, %d regenerated, pool %d strings" % (
            c["seed"], c["n_reads"], len(c["alignments"]), len(c["merge_decisions"]), len(c["intervals_for_graph"]),
            len(c["merge"]["regenerated"]), len(c["pool"])))
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
