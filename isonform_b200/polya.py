"""M0 (host side): polyA/polyT tail collapsing that runs before minimizer extraction.

Mirror of remove_read_polyA_ends (reference main:47-67, called with (12, 1) at main:351): within the
last min(len//2, 100) characters every run of 'A' or 'T' longer than `threshold_len` is replaced
by `to_len` copies of that character.  Stays on the host: it is O(100) per read.
"""
import re

_RUN = re.compile(r"A+|T+")


def remove_read_polyA_ends(seq, threshold_len, to_len):
    window = min(len(seq) // 2, 100)
    # Python's seq[:-0] is '' and seq[-0:] the whole string: reads of length 0/1 are scanned entirely
    head, tail = (seq[:-window], seq[-window:]) if window else ("", seq)

    def collapse(m):
        run = m.group(0)
        return run[0] * to_len if len(run) > threshold_len else run

    return head + _RUN.sub(collapse, tail)
