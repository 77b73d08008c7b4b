"""TEST INFRASTRUCTURE: stand-in for the absent `parasail` module, backed by the CPU oracle, so that the UNMODIFIED
reference tree can be imported and run in the build container (consensus.py:7 imports parasail at module load)."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402


class _Result:
    def __init__(self, s1, s2, open_, ext, matrix):
        self.score, self.end_query, self.end_ref, runs = oracle.sg_align(s1, s2, matrix[0], matrix[1], open_, ext)
        self.saturated = False
        self.cigar = types.SimpleNamespace(decode=oracle.cigar_runs_to_string(runs).encode())


def matrix_create(alphabet, match, mismatch):
    return (match, mismatch)


def sg_trace_scan_16(s1, s2, open_, ext, matrix):
    return _Result(s1, s2, open_, ext, matrix)


sg_trace_scan_32 = sg_trace_scan_16
