"""ctypes binding of oracle/libisof_oracle.so (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libisof_oracle.so")

OPS = {0: "M", 1: "I", 2: "D", 7: "=", 8: "X"}


def build(force=False):
    """Compile the C restatement (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "isof_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _SO


class TieBreak(C.Structure):
    _fields_ = [("h_e_before_f", C.c_int32), ("gap_open_on_tie", C.c_int32),
                ("end_col_first", C.c_int32), ("swap_id", C.c_int32)]


def all_tiebreaks():
    """The 16 settings of the tie-break table, default first."""
    return [TieBreak(a, b, c, d) for d in (0, 1) for c in (0, 1) for b in (0, 1) for a in (0, 1)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p = C.POINTER(C.c_uint8)
        L.isof_oracle_pyhash.restype = C.c_int64
        L.isof_oracle_pyhash.argtypes = [C.c_char_p, C.c_int64]
        L.isof_oracle_minimizers.restype = C.c_int64
        L.isof_oracle_minimizers.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.isof_oracle_minimizers_batch.restype = C.c_int64
        L.isof_oracle_minimizers_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.isof_oracle_sg_align_tb.restype = C.c_int32
        L.isof_oracle_sg_align_tb.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_int32, C.c_int32, C.POINTER(TieBreak),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                              C.POINTER(C.c_int32), C.c_void_p, C.c_int32,
                                              C.POINTER(C.c_int32)]
        L.isof_oracle_sg_align_pairs.restype = C.c_int64
        L.isof_oracle_sg_align_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                 C.c_int]
        L.isof_oracle_sg_align_pairs_simd.restype = C.c_int64
        L.isof_oracle_sg_align_pairs_simd.argtypes = L.isof_oracle_sg_align_pairs.argtypes
        L.isof_oracle_edit_distance.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
        L.isof_oracle_edit_distance.restype = C.c_int32
        L.isof_oracle_edit_distance_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                      C.c_void_p, C.c_int]
        L.isof_oracle_edit_distance_pairs.restype = C.c_int64
        L.isof_oracle_edit_path.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_void_p, C.c_int64,
                                            C.POINTER(C.c_int32)]
        L.isof_oracle_edit_path.restype = C.c_int64
        L.isof_oracle_num_threads.restype = C.c_int32
        L.isof_oracle_simd_lanes.restype = C.c_int32
        L.isof_oracle_simd_isa.restype = C.c_char_p
        del u8p
        _lib = L
    return _lib


def _b(s):
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


def pyhash(s):
    b = _b(s)
    return int(lib().isof_oracle_pyhash(b, len(b)))


def minimizers(seq, k, w_size):
    """Positions of get_kmer_minimizers(seq, k, w_size) (main:81-110)."""
    b = _b(seq)
    cap = max(len(b), 1)
    out = np.empty(cap, dtype=np.int32)
    n = lib().isof_oracle_minimizers(b, len(b), k, w_size, out.ctypes.data, cap)
    if n < 0:
        raise ValueError("w_size < k")
    return out[:n].copy()


def minimizers_batch(cat, off, k, w_size, threads=0):
    cat = np.ascontiguousarray(cat, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    R = len(off) - 1
    cap = int(cat.size) + R
    out = np.empty(cap, dtype=np.int32)
    out_off = np.empty(R + 1, dtype=np.int64)
    tot = lib().isof_oracle_minimizers_batch(cat.ctypes.data, off.ctypes.data, R, k, w_size,
                                             out.ctypes.data, out_off.ctypes.data, cap, threads)
    return out[:tot].copy(), out_off


def sg_align(s1, s2, match=2, mismatch=-2, gap_open=12, gap_ext=1, tiebreak=None):
    """Restated parasail.sg_trace_scan + cigar.  Returns (score, end_query, end_ref, runs[uint32])."""
    a, b = _b(s1), _b(s2)
    if len(a) == 0 or len(b) == 0:
        raise ValueError("empty sequence")
    cap = len(a) + len(b) + 2
    runs = np.empty(cap, dtype=np.uint32)
    sc, eq, er, nr = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    tb = C.byref(tiebreak) if tiebreak is not None else None
    rc = lib().isof_oracle_sg_align_tb(a, len(a), b, len(b), match, mismatch, gap_open, gap_ext, tb,
                                       C.byref(sc), C.byref(eq), C.byref(er), runs.ctypes.data, cap,
                                       C.byref(nr))
    if rc < 0:
        raise RuntimeError("oracle sg_align failed rc=%d" % rc)
    return sc.value, eq.value, er.value, runs[:nr.value].copy()


def sg_align_pairs(cat, off, ia, ib, match, mismatch, gap_open, gap_ext, threads=0, want_cigar=True, simd=False):
    """Batch alignment on the CPU.  simd=False: the scalar checker, one pair per thread.  simd=True: 16 pairs
    per thread in int16 lanes (the CPU baseline of bench.py); bit-identical results."""
    cat = np.ascontiguousarray(cat, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ia = np.ascontiguousarray(ia, dtype=np.int32)
    ib = np.ascontiguousarray(ib, dtype=np.int32)
    P = len(ia)
    lens = np.diff(off)
    cap = int((lens[ia] + lens[ib] + 2).sum()) if want_cigar else 0
    score = np.empty(P, np.int32); eq = np.empty(P, np.int32); er = np.empty(P, np.int32)
    cig = np.empty(max(cap, 1), np.uint32)
    coff = np.empty(P + 1, np.int64)
    fn = lib().isof_oracle_sg_align_pairs_simd if simd else lib().isof_oracle_sg_align_pairs
    tot = fn(cat.ctypes.data, off.ctypes.data, ia.ctypes.data, ib.ctypes.data,
                                           P, match, mismatch, gap_open, gap_ext, score.ctypes.data,
                                           eq.ctypes.data, er.ctypes.data,
                                           cig.ctypes.data if want_cigar else None, coff.ctypes.data, cap,
                                           threads)
    if tot < 0:
        raise RuntimeError("oracle sg_align_pairs failed")
    return score, eq, er, (cig[:tot].copy() if want_cigar else None), coff


def edit_distance(s1, s2):
    """Global unit-cost edit distance (Myers/Hyyro bit-vector); config 5's "edlib CPU" stand-in."""
    a, b = _b(s1), _b(s2)
    return int(lib().isof_oracle_edit_distance(a, len(a), b, len(b)))


def edit_path(s1, s2):
    """(distance, uint32 runs) of one optimal global unit-cost alignment: plain O(nm) matrix, walked back preferring
    the diagonal, then 'I' (consumes s1), then 'D' -- the checker of isof_edit_distance_pairs' path."""
    a, b = _b(s1), _b(s2)
    runs = np.empty(len(a) + len(b) + 2, np.uint32)
    d = C.c_int32(0)
    nr = lib().isof_oracle_edit_path(a, len(a), b, len(b), runs.ctypes.data, len(runs), C.byref(d))
    if nr < 0:
        raise MemoryError("oracle edit_path")
    return int(d.value), runs[:nr].copy()


def edit_distance_pairs(cat, off, ia, ib, threads=0):
    cat = np.ascontiguousarray(cat, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ia = np.ascontiguousarray(ia, dtype=np.int32)
    ib = np.ascontiguousarray(ib, dtype=np.int32)
    dist = np.empty(len(ia), np.int32)
    if lib().isof_oracle_edit_distance_pairs(cat.ctypes.data, off.ctypes.data, ia.ctypes.data, ib.ctypes.data, len(ia),
                                             dist.ctypes.data, threads) < 0:
        raise RuntimeError("oracle edit_distance_pairs failed")
    return dist


def num_threads():
    return int(lib().isof_oracle_num_threads())


def simd_lanes():
    return int(lib().isof_oracle_simd_lanes())


def simd_isa():
    return lib().isof_oracle_simd_isa().decode()


def cigar_runs_to_string(runs):
    return "".join("%d%s" % (int(r) >> 4, OPS[int(r) & 15]) for r in runs)


def cigar_runs_to_tuples(runs):
    return [(int(r) >> 4, OPS[int(r) & 15]) for r in runs]
