"""In-process consensus in place of the `spoa` subprocess.

`IsoformGeneration.run_spoa(reads, spoa_out_file)` (IsoformGeneration.py:143-156) writes nothing itself: its callers
have already written the member sequences to `reads` (a FASTA file); it spawns `spoa reads -l 0 -r 0 -g -2`, sends the
output to `spoa_out_file` and returns line 2 of it.  `run_spoa` below has the same signature and result shape, reads the
same FASTA file, and computes the consensus with the library's host POA (csrc/host_poa.cpp) -- no process spawn (887-1503
per batch in the reference).  It is an independent POA with spoa's scoring for that command line, not a bit-exact spoa
(spoa is not vendored by the reference): `scripts/isof_spoa` exposes the same code as an executable so that the
unmodified reference can be run with the identical consensus on the other side of a comparison.
"""
from . import _lib


def consensus(seqs, match=5, mismatch=-4, gap=-2):
    return _lib.poa_consensus([s for s in seqs if s], match, mismatch, gap)


def read_fasta_sequences(path):
    seqs, cur = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if cur:
                    seqs.append("".join(cur))
                cur = []
            else:
                cur.append(line.strip())
    if cur:
        seqs.append("".join(cur))
    return seqs


def run_spoa(reads, spoa_out_file):
    """Drop-in for IsoformGeneration.run_spoa: same arguments, returns the consensus string; the output file is written
    in spoa's format (header line, consensus line) because the reference reads it back in other places."""
    con = consensus(read_fasta_sequences(reads))
    with open(spoa_out_file, "w") as out:
        out.write(">Consensus LN:i:%d\n%s\n" % (len(con), con))
    return con
