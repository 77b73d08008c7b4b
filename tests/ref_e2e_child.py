"""Child process of tests/test_reference_e2e.py: runs the UNMODIFIED reference `main(args)` on one
FASTQ batch, either as is (with the oracle standing in for the absent parasail) or with the B200 seams
patched in over a test double of the GPU context.  Build container only (/root/reference)."""
import argparse
import importlib.machinery
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)

import oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fastq")
    ap.add_argument("--outfolder")
    ap.add_argument("--patched", type=int, default=0)
    a = ap.parse_args()

    stub = types.ModuleType("parasail")

    class _R:
        def __init__(self, s1, s2, o, e, m):
            self.score, self.end_query, self.end_ref, runs = oracle.sg_align(s1, s2, m[0], m[1], o, e)
            self.saturated = False
            self.cigar = types.SimpleNamespace(decode=oracle.cigar_runs_to_string(runs).encode())

    stub.matrix_create = lambda alpha, m, x: (m, x)
    stub.sg_trace_scan_16 = lambda s1, s2, o, e, m: _R(s1, s2, o, e, m)
    stub.sg_trace_scan_32 = stub.sg_trace_scan_16
    sys.modules["parasail"] = stub

    loader = importlib.machinery.SourceFileLoader("isonform_ref_main_e2e", os.path.join(REF, "main"))
    ref = importlib.util.module_from_spec(importlib.util.spec_from_loader("isonform_ref_main_e2e", loader))
    loader.exec_module(ref)
    from modules import IsoformGeneration

    spoa_calls = [0]

    def fake_spoa(reads_file, out_file):           # deterministic stand-in for the absent spoa binary
        spoa_calls[0] += 1
        seqs = [l.strip() for l in open(reads_file) if not l.startswith(">")]
        seqs.sort(key=lambda s: (len(s), s))
        return seqs[len(seqs) // 2]
    IsoformGeneration.run_spoa = fake_spoa

    stats = {}
    if a.patched:
        from fake_ctx import FakeContext
        from isonform_b200 import patch
        ctx = FakeContext()
        memo = patch.install(vars(ref), ctx=ctx)
        memo.bubbles.check_paths = True           # every find_paths call is cross-checked against the reference's own
        stats = {"ctx": ctx, "memo": memo}

    args = types.SimpleNamespace(fastq=a.fastq, outfolder=a.outfolder, k=20, w=31, xmin=40, xmax=80, T=0.1, exact=True,
                                 disable_numpy=False, delta_len=5, max_seqs_to_spoa=200, max_seqs=1000,
                                 exact_instance_limit=0, set_w_dynamically=False, verbose=False, compression=False,
                                 delta_iso_len_3=30, delta_iso_len_5=50, parallel=True, slow=False)
    ref.main(args)
    sys.stderr.write("SPOA_SPAWNS %d\n" % spoa_calls[0])
    if a.patched:
        c = stats["ctx"].calls
        sys.stderr.write("B200_SEAM_CALLS minimizers=%d spans=%d align_calls=%d pairs=%d memo_hits=%d memo_misses=%d\n"
                         % (c["minimizers"], c["spans"], c["align"], c["pairs"], stats["memo"].hits, stats["memo"].misses))
        b = stats["memo"].bubbles.stats
        sys.stderr.write("B200_BUBBLES iterations=%d speculated=%d prefetch_calls=%d prefetched=%d spoa_calls=%d spoa_hits=%d\n"
                         % (b["iterations"], b["speculated_pairs"], b["prefetch_calls"], b["prefetched"], b["spoa_calls"],
                            b["spoa_hits"]))
        sys.stderr.write("B200_FIND_PATHS checked=%d\n" % b.get("find_paths_checked", 0))


if __name__ == "__main__":
    main()
