"""ctypes binding of libisonform_b200.so (include/isonform_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
product raises.  `Context` is one (process, GPU) handle; host-array methods take numpy arrays,
`*_dev` methods take raw device pointers (e.g. ``torch.Tensor.data_ptr()``).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libisonform_b200.so")
SO_PATH_DEBUG = os.path.join(HERE, "libisonform_b200_dbg.so")     # same sources with -DISOF_CHECK

N_FEATURES = 12
OPS = {0: "M", 1: "I", 2: "D", 7: "=", 8: "X"}
ISOF_OK, ERR_CUDA, ERR_ARG, ERR_CAPACITY, ERR_NOMEM, ERR_RANGE = 0, -1, -2, -3, -4, -5


class IsofError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("isonform_b200 error %d: %s" % (code, msg))
        self.code = code


class Profile(C.Structure):
    _fields_ = [("launches", C.c_int64), ("dp_launches", C.c_int64), ("dp_ms", C.c_double),
                ("traceback_ms", C.c_double), ("minimizer_ms", C.c_double), ("other_ms", C.c_double),
                ("dp_cells", C.c_int64), ("dp_cells_padded", C.c_int64), ("dp_cells_packed", C.c_int64), ("dp_cells_profile", C.c_int64),
                ("trace_bytes", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("band_pairs_retried", C.c_int64), ("trace_replans", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["isof_version", "isof_init", "isof_destroy", "isof_last_error", "isof_release_scratch", "isof_set_stream", "isof_set_trace_budget", "isof_set_packed", "isof_set_overlap", "isof_set_tiebreak", "isof_set_short", "isof_set_fused_small", "isof_debug_fused_phases", "isof_set_band", "isof_set_traceback_warp", "isof_debug_violations", "isof_debug_selftest", "isof_debug_set_span_hash_mask",
           "isof_profile_enable", "isof_profile_reset", "isof_profile_get", "isof_int32_peak", "isof_minimizers_batch",
           "isof_minimizers_batch_dev", "isof_spans_batch", "isof_spans_multi", "isof_spans_batch_dev", "isof_sg_align_pairs", "isof_sg_align_pairs_dev", "isof_sg_align_batch",
           "isof_sg_features_pairs", "isof_edit_distance_pairs", "isof_edit_distance_pairs_dev", "isof_wis_picks", "isof_span_instances", "isof_poa_consensus"]

_lib = None
_lib_debug = None


def load(debug=False):
    """Load the shared library (no CUDA call yet).  Raises if it has not been built.
    debug=True loads the assertion-carrying build (tests only)."""
    global _lib, _lib_debug
    if debug and _lib_debug is not None:
        return _lib_debug
    if not debug and _lib is not None:
        return _lib
    path = SO_PATH_DEBUG if debug else SO_PATH
    if not os.path.exists(path):
        raise ImportError("%s is not built: run `python -m isonform_b200.build` "
                          "(needs nvcc; there is no CPU fallback)" % os.path.basename(path))
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.isof_version.restype = C.c_int
    L.isof_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.isof_destroy.argtypes = [vp]
    L.isof_last_error.argtypes = [vp]
    L.isof_last_error.restype = C.c_char_p
    L.isof_release_scratch.argtypes = [vp]
    L.isof_set_stream.argtypes = [vp, vp]
    L.isof_set_trace_budget.argtypes = [vp, C.c_size_t]
    L.isof_set_packed.argtypes = [vp, C.c_int]
    L.isof_set_overlap.argtypes = [vp, C.c_int]
    L.isof_set_tiebreak.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.isof_set_short.argtypes = [vp, C.c_int]
    L.isof_set_fused_small.argtypes = [vp, C.c_int]
    L.isof_set_band.argtypes = [vp, C.c_int]
    L.isof_set_traceback_warp.argtypes = [vp, C.c_int]
    L.isof_debug_fused_phases.argtypes = [vp, vp]
    L.isof_debug_violations.argtypes = [vp, C.POINTER(C.c_uint)]
    L.isof_debug_selftest.argtypes = [vp]
    L.isof_debug_set_span_hash_mask.argtypes = [vp, C.c_uint64]
    L.isof_profile_enable.argtypes = [vp, C.c_int]
    L.isof_profile_reset.argtypes = [vp]
    L.isof_profile_get.argtypes = [vp, C.POINTER(Profile)]
    L.isof_int32_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.isof_minimizers_batch.argtypes = [vp, vp, vp, i32, i32, i32, vp, i64, vp, C.POINTER(i64)]
    L.isof_minimizers_batch_dev.argtypes = [vp, vp, vp, i32, i64, i32, i32, vp, i64, vp, C.POINTER(i64)]
    L.isof_spans_batch.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp,
                                   C.POINTER(i64), C.POINTER(i64)]
    L.isof_spans_multi.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp,
                                   C.POINTER(i64), C.POINTER(i64)]
    L.isof_spans_batch_dev.argtypes = [vp, vp, vp, i32, vp, vp, i64, i32, i32, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp,
                                       C.POINTER(i64), C.POINTER(i64)]
    L.isof_sg_align_pairs.argtypes = [vp, vp, vp, i32, vp, vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, i64, vp]
    L.isof_sg_align_pairs_dev.argtypes = [vp, vp, vp, i32, i64, vp, vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, i64,
                                          vp, vp, i32, C.POINTER(i64)]
    L.isof_sg_align_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, i64, vp]
    L.isof_sg_features_pairs.argtypes = [vp, vp, vp, i32, vp, vp, i64, i32, i32, i32, i32, i32, vp, vp, vp, i64, vp]
    L.isof_edit_distance_pairs.argtypes = [vp, vp, vp, i32, vp, vp, i64, i32, vp, vp, i64, vp]
    L.isof_edit_distance_pairs_dev.argtypes = [vp, vp, vp, i32, i64, vp, vp, i64, i32, vp, vp, i64, vp, C.POINTER(i64)]
    L.isof_wis_picks.argtypes = [i32, vp, vp, vp, vp, i32, vp, i64, vp, C.POINTER(i64)]
    L.isof_span_instances.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    L.isof_poa_consensus.argtypes = [vp, vp, i32, i32, i32, i32, vp, i64, C.POINTER(i64)]
    for name in EXPORTS:
        getattr(L, name)
        if name not in ("isof_last_error",):
            getattr(L, name).restype = C.c_int
    if debug:
        _lib_debug = L
    else:
        _lib = L
    return L


def pack_sequences(seqs):
    """list[str|bytes] -> (uint8 cat, int64 off)."""
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    cat = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(cat), off


def runs_to_string(runs):
    return "".join("%d%s" % (int(r) >> 4, OPS[int(r) & 15]) for r in runs)


def runs_to_tuples(runs):
    return [(int(r) >> 4, OPS[int(r) & 15]) for r in runs]


class Context:
    """One GPU context.  Not thread-safe; calls are stream-ordered."""

    def __init__(self, device=0, debug=False):
        self._L = load(debug)
        h = C.c_void_p()
        rc = self._L.isof_init(int(device), C.byref(h))
        if rc != ISOF_OK:
            raise IsofError(rc, "isof_init(device=%d) failed: no usable CUDA device (there is no CPU fallback)" % device)
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.isof_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != ISOF_OK and rc not in allow:
            raise IsofError(rc, (self._L.isof_last_error(self._h) or b"").decode())
        return rc

    # ------------------------------------------------------------------ plumbing
    def release_scratch(self):
        """Return the ctx's grow-only device scratch to the device (synchronises first)."""
        self._check(self._L.isof_release_scratch(self._h))

    def set_stream(self, cuda_stream):
        self._check(self._L.isof_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_trace_budget(self, nbytes):
        self._check(self._L.isof_set_trace_budget(self._h, int(nbytes)))

    def set_packed(self, on=True):
        """Enable (default) / disable the packed 16-bit DP path; results are identical either way."""
        self._check(self._L.isof_set_packed(self._h, 1 if on else 0))

    def set_traceback_warp(self, mode=2):
        """Regular path's traceback: True/1 = one warp per pair with look-ahead, False/0 = one thread per pair, 2 = by chunk
        size (default).  Same results."""
        self._check(self._L.isof_set_traceback_warp(self._h, int(mode)))

    def set_band(self, half_width=1024):
        """Half-width of the banded trace of long pairs in feature-only calls (0 = off); results never depend on it."""
        self._check(self._L.isof_set_band(self._h, int(half_width)))

    def debug_fused_phases(self):
        """Last fused small call, pair 0: SM cycles of (load+encode, DP, traceback, features+stores), then the list of
        (begin, end) ns of every stripe's DP when the pair ran in chase mode."""
        out = np.zeros(56, np.int64)
        self._check(self._L.isof_debug_fused_phases(self._h, out.ctypes.data))
        return tuple(int(x) for x in out[:4]) + ([(int(out[8 + 2 * g]), int(out[9 + 2 * g]), int(out[40 + g]) >> 40,
                                                   int(out[40 + g]) & ((1 << 40) - 1)) for g in range(16)],)

    def set_fused_small(self, on=True):
        """One fused launch for tiny host-buffer calls (default on); results are identical either way."""
        self._check(self._L.isof_set_fused_small(self._h, 1 if on else 0))

    def set_short(self, mode=1):
        """Short-pair regime for all-short calls: 0/False never, 1/True when the call is large enough (default),
        2 always (tests).  Results are identical either way."""
        self._check(self._L.isof_set_short(self._h, int(mode)))

    def set_tiebreak(self, h_e_before_f=0, gap_open_on_tie=0, end_col_first=0, swap_id=0):
        """Tie-break table of the aligner (isof_set_tiebreak); all zero = the library as restated."""
        self._check(self._L.isof_set_tiebreak(self._h, int(h_e_before_f), int(gap_open_on_tie), int(end_col_first),
                                              int(swap_id)))

    def debug_violations(self):
        """(mask, compiled_in): violated assertion classes since the last call; see isof_debug_violations."""
        m = C.c_uint(0)
        self._check(self._L.isof_debug_violations(self._h, C.byref(m)))
        return int(m.value) & 0x7fffffff, bool(m.value & 0x80000000)

    def debug_set_span_hash_mask(self, mask=0x7fffffffffffffff):
        """Test knob: hash bits kernel (c) keeps for k-mers that do not 2-bit pack (forces collisions)."""
        self._check(self._L.isof_debug_set_span_hash_mask(self._h, int(mask)))

    def debug_selftest(self):
        self._check(self._L.isof_debug_selftest(self._h))

    def set_overlap(self, on=True):
        """Two-stream chunk pipelining (default on); off gives exact per-kernel event timings."""
        self._check(self._L.isof_set_overlap(self._h, 1 if on else 0))

    def profile_enable(self, on=True):
        self._check(self._L.isof_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        self._check(self._L.isof_profile_reset(self._h))

    def profile(self):
        p = Profile()
        self._check(self._L.isof_profile_get(self._h, C.byref(p)))
        return p.as_dict()

    def int32_peak(self):
        """Measured INT32 ALU-pipe peak (lane-ops/s) of this GPU -- the integer roofline denominator."""
        v = C.c_double(0)
        self._check(self._L.isof_int32_peak(self._h, C.byref(v)))
        return float(v.value)

    # ------------------------------------------------------------------ (a) minimizers
    def minimizers_batch(self, cat, off, k, w_size):
        """(positions int32, offsets int64[R+1]) for concatenated reads."""
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        R = len(off) - 1
        # ~0.15 minimizers per base on random sequence; grow on ISOF_ERR_CAPACITY
        cap = max(int(cat.size) // 3 + 4 * R + 64, 64)
        out_off = np.empty(R + 1, dtype=np.int64)
        n_out = C.c_int64(0)
        while True:
            out = np.empty(cap, dtype=np.int32)
            rc = self._L.isof_minimizers_batch(self._h, cat.ctypes.data, off.ctypes.data, R, k, w_size, out.ctypes.data,
                                               cap, out_off.ctypes.data, C.byref(n_out))
            if rc == ERR_CAPACITY:
                cap = int(n_out.value)
                continue
            self._check(rc)
            return out[:n_out.value], out_off

    def minimizers_batch_dev(self, d_bases, d_off, n_reads, n_bases, k, w_size, d_out_pos, out_cap, d_out_off):
        n_out = C.c_int64(0)
        self._check(self._L.isof_minimizers_batch_dev(self._h, d_bases, d_off, n_reads, n_bases, k, w_size, d_out_pos,
                                                      out_cap, d_out_off, C.byref(n_out)))
        return int(n_out.value)

    # ------------------------------------------------------------------ (c) pair database + span search
    def spans_batch(self, cat, off, min_pos, min_off, k, x_low, x_high, delta_len, batch_read_off=None):
        """dict of numpy arrays, see isof_spans_batch / isof_spans_multi in include/isonform_b200.h.
        batch_read_off (int32[n_batches+1]) turns the call into the super-batched form."""
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        min_pos = np.ascontiguousarray(min_pos, dtype=np.int32)
        min_off = np.ascontiguousarray(min_off, dtype=np.int64)
        R = len(off) - 1
        cap = max(8 * len(min_pos) + 64, 64)
        n_rec, n_b = C.c_int64(0), C.c_int64(0)
        while True:
            a = {name: np.empty(cap, np.int32) for name in ("rec_read", "rec_p1", "rec_p2", "rec_count", "rec_bucket",
                                                            "sorted_rec")}
            boff = np.zeros(cap + 1, np.int32)
            bdrop = np.zeros(cap, np.uint8)
            if batch_read_off is not None:
                bro = np.ascontiguousarray(batch_read_off, dtype=np.int32)
                bro_ptr, nbt = bro.ctypes.data, len(bro) - 1
            else:
                bro_ptr, nbt = None, 1
            rc = self._L.isof_spans_multi(self._h, cat.ctypes.data, off.ctypes.data, R, bro_ptr, nbt, min_pos.ctypes.data,
                                          min_off.ctypes.data, k, x_low, x_high, delta_len, cap, a["rec_read"].ctypes.data,
                                          a["rec_p1"].ctypes.data, a["rec_p2"].ctypes.data, a["rec_count"].ctypes.data,
                                          a["rec_bucket"].ctypes.data, a["sorted_rec"].ctypes.data, boff.ctypes.data,
                                          bdrop.ctypes.data, C.byref(n_rec), C.byref(n_b))
            if rc == ERR_CAPACITY:
                cap = int(n_rec.value)
                continue
            self._check(rc)
            break
        G, B = int(n_rec.value), int(n_b.value)
        out = {k_: v[:G] for k_, v in a.items()}
        out["bucket_off"] = boff[:B + 1]
        out["bucket_dropped"] = bdrop[:B]
        return out

    def spans_batch_dev(self, d_bases, d_off, n_reads, d_min_pos, d_min_off, n_min, k, x_low, x_high, delta_len, rec_cap,
                        d_rec_read, d_rec_p1, d_rec_p2, d_rec_count, d_rec_bucket, d_sorted_rec, d_bucket_off, d_bucket_dropped):
        n_rec, n_b = C.c_int64(0), C.c_int64(0)
        self._check(self._L.isof_spans_batch_dev(self._h, d_bases, d_off, n_reads, d_min_pos, d_min_off, n_min, k, x_low,
                                                 x_high, delta_len, rec_cap, d_rec_read, d_rec_p1, d_rec_p2, d_rec_count,
                                                 d_rec_bucket, d_sorted_rec, d_bucket_off, d_bucket_dropped, C.byref(n_rec),
                                                 C.byref(n_b)))
        return int(n_rec.value), int(n_b.value)

    # ------------------------------------------------------------------ (b) alignment
    def sg_align_pairs(self, cat, off, pair_a, pair_b, match=2, mismatch=-2, gap_open=12, gap_ext=1, want_cigar=True,
                       features_delta_len=None, cigar_cap=None):
        """Returns dict(score, end_q, end_r, cigar_off, cigar (uint32 runs or None), features or None)."""
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        pa = np.ascontiguousarray(pair_a, dtype=np.int32)
        pb = np.ascontiguousarray(pair_b, dtype=np.int32)
        P = len(pa)
        if len(pb) != P:
            raise ValueError("pair_a and pair_b differ in length")
        score = np.empty(P, np.int32)
        coff = np.zeros(P + 1, np.int64)
        feats = None
        if cigar_cap is None:
            if want_cigar and P:
                lens = np.diff(off)
                ns = max(len(lens), 1)
                # an alignment has at most n+m+1 runs; reads at 5 % error produce about (n+m)/12.  Start at a
                # fifth of the worst case (a too small buffer costs a second full run: ISOF_ERR_CAPACITY);
                # out-of-range indices are reported by the library, not here
                if not len(lens):
                    worst = 1
                else:
                    # exact sum of n+m+2 over the pairs; int32 gathers (4x cheaper than fancy-indexing int64) and a
                    # clip only when an index is out of range (the library then reports ISOF_ERR_ARG)
                    l32 = lens.astype(np.int32) if lens.max(initial=0) < (1 << 31) else lens
                    ia = pa if (pa.min() >= 0 and pa.max() < ns) else np.clip(pa, 0, ns - 1)
                    ib = pb if (pb.min() >= 0 and pb.max() < ns) else np.clip(pb, 0, ns - 1)
                    worst = int(l32.take(ia).sum(dtype=np.int64) + l32.take(ib).sum(dtype=np.int64)) + 2 * P
                cap = int(min(worst, worst // 5 + 64 * P + (1 << 16)))
            else:
                cap = 0
        else:
            cap = int(cigar_cap)
        while True:
            cig = np.empty(max(cap, 1), np.uint32) if want_cigar else None
            cig_ptr = cig.ctypes.data if want_cigar else None
            if features_delta_len is None:
                eq = np.empty(P, np.int32)
                er = np.empty(P, np.int32)
                rc = self._L.isof_sg_align_pairs(self._h, cat.ctypes.data, off.ctypes.data, len(off) - 1, pa.ctypes.data,
                                                 pb.ctypes.data, P, match, mismatch, gap_open, gap_ext, score.ctypes.data,
                                                 eq.ctypes.data, er.ctypes.data, cig_ptr, cap if want_cigar else 0,
                                                 coff.ctypes.data)
            else:
                eq = er = None
                feats = np.empty((P, N_FEATURES), np.int32)
                rc = self._L.isof_sg_features_pairs(self._h, cat.ctypes.data, off.ctypes.data, len(off) - 1,
                                                    pa.ctypes.data, pb.ctypes.data, P, match, mismatch, gap_open, gap_ext,
                                                    int(features_delta_len), score.ctypes.data, feats.ctypes.data,
                                                    cig_ptr, cap if want_cigar else 0, coff.ctypes.data)
            if rc == ERR_CAPACITY and want_cigar:
                cap = int(coff[P])
                continue
            self._check(rc)
            break
        return {"score": score, "end_q": eq, "end_r": er, "cigar_off": coff,
                "cigar": (cig[:coff[P]] if want_cigar else None), "features": feats}

    def sg_align_one(self, s1, s2, match=2, mismatch=-2, gap_open=12, gap_ext=1):
        """One pair, lean: (score, end_q, end_r, runs uint32[]) with pre-allocated buffers -- the per-pair seam
        (consensus.parasail_alignment on a memo miss).  Goes through isof_sg_align_pairs, i.e. the fused tiny-call
        kernel when the pair is small enough."""
        b1 = s1.encode("latin-1") if isinstance(s1, str) else bytes(s1)
        b2 = s2.encode("latin-1") if isinstance(s2, str) else bytes(s2)
        n, m = len(b1), len(b2)
        st = getattr(self, "_one", None)
        if st is None or st["cap"] < n + m + 2:
            cap = max(2 * (n + m + 2), 4096)
            st = self._one = {"cap": cap, "off": np.zeros(3, np.int64), "pa": np.zeros(1, np.int32), "pb": np.ones(1, np.int32),
                              "out": np.zeros(3, np.int32), "coff": np.zeros(2, np.int64), "runs": np.empty(cap, np.uint32)}
            st["ptr"] = (st["off"].ctypes.data, st["pa"].ctypes.data, st["pb"].ctypes.data, st["out"].ctypes.data,
                         st["out"].ctypes.data + 4, st["out"].ctypes.data + 8, st["runs"].ctypes.data, st["coff"].ctypes.data)
        st["off"][1] = n
        st["off"][2] = n + m
        off_p, pa_p, pb_p, sc_p, eq_p, er_p, runs_p, coff_p = st["ptr"]
        self._check(self._L.isof_sg_align_pairs(self._h, b1 + b2, off_p, 2, pa_p, pb_p, 1, match, mismatch, gap_open, gap_ext,
                                                sc_p, eq_p, er_p, runs_p, st["cap"], coff_p))
        out = st["out"]
        return int(out[0]), int(out[1]), int(out[2]), st["runs"][:int(st["coff"][1])].copy()

    def sg_align_batch(self, s1_list, s2_list, match=2, mismatch=-2, gap_open=12, gap_ext=1):
        """The SURVEY 8b concatenated form (isof_sg_align_batch)."""
        c1, o1 = pack_sequences(s1_list)
        c2, o2 = pack_sequences(s2_list)
        P = len(s1_list)
        score = np.empty(P, np.int32); eq = np.empty(P, np.int32); er = np.empty(P, np.int32)
        coff = np.zeros(P + 1, np.int64)
        cap = int((np.diff(o1) + np.diff(o2) + 2).sum()) if P else 1
        cig = np.empty(max(cap, 1), np.uint32)
        self._check(self._L.isof_sg_align_batch(self._h, c1.ctypes.data, o1.ctypes.data, c2.ctypes.data, o2.ctypes.data, P,
                                                match, mismatch, gap_open, gap_ext, score.ctypes.data, eq.ctypes.data,
                                                er.ctypes.data, cig.ctypes.data, cap, coff.ctypes.data))
        return {"score": score, "end_q": eq, "end_r": er, "cigar_off": coff, "cigar": cig[:coff[P]]}

    def sg_align_pairs_dev(self, d_cat, d_off, n_seqs, n_bases, d_pa, d_pb, P, match, mismatch, gap_open, gap_ext, d_score,
                           d_endq, d_endr, d_cigar, cigar_cap, d_cigar_off, d_features=None, delta_len=0):
        total = C.c_int64(0)
        self._check(self._L.isof_sg_align_pairs_dev(self._h, d_cat, d_off, n_seqs, n_bases, d_pa, d_pb, P, match, mismatch,
                                                    gap_open, gap_ext, d_score, d_endq, d_endr, d_cigar, cigar_cap,
                                                    d_cigar_off, d_features, delta_len, C.byref(total)))
        return int(total.value)


    # ------------------------------------------------------------------ (b') unit-cost edit distance
    def edit_distance_pairs(self, cat, off, pair_a, pair_b, k=-1, want_cigar=False):
        """Global unit-cost edit distance of pool[pair_a[p]] (query) against pool[pair_b[p]] (target): what
        edlib.align(q, t, mode="NW", task="distance"|"path", k=k) computes.  Returns dict(dist int32[P] (-1 above k),
        cigar_off, cigar) -- the last two None unless want_cigar."""
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        pa = np.ascontiguousarray(pair_a, dtype=np.int32)
        pb = np.ascontiguousarray(pair_b, dtype=np.int32)
        P = len(pa)
        if len(pb) != P:
            raise ValueError("pair_a and pair_b differ in length")
        dist = np.empty(P, np.int32)
        if not want_cigar:
            self._check(self._L.isof_edit_distance_pairs(self._h, cat.ctypes.data, off.ctypes.data, len(off) - 1, pa.ctypes.data,
                                                         pb.ctypes.data, P, int(k), dist.ctypes.data, None, 0, None))
            return {"dist": dist, "cigar_off": None, "cigar": None}
        coff = np.zeros(P + 1, np.int64)
        # a path has at most n+m runs; unrelated sequences come close to half of that, noisy copies stay far below.  A
        # too small buffer costs a second full run (ISOF_ERR_CAPACITY), an untouched tail of np.empty costs nothing
        lens = np.diff(off).astype(np.int64)
        ok = P and len(lens) and pa.min() >= 0 and pb.min() >= 0 and max(pa.max(), pb.max()) < len(lens)
        worst = int(lens.take(pa).sum() + lens.take(pb).sum()) + 2 * P if ok else 1
        cap = int(min(worst, worst // 2 + 64 * P + (1 << 16)))
        while True:
            cig = np.empty(max(cap, 1), np.uint32)
            rc = self._L.isof_edit_distance_pairs(self._h, cat.ctypes.data, off.ctypes.data, len(off) - 1, pa.ctypes.data,
                                                  pb.ctypes.data, P, int(k), dist.ctypes.data, cig.ctypes.data, cap,
                                                  coff.ctypes.data)
            if rc == ERR_CAPACITY:
                cap = int(coff[P])
                continue
            self._check(rc)
            break
        return {"dist": dist, "cigar_off": coff, "cigar": cig[:coff[P]]}

    def edit_distance_pairs_dev(self, d_cat, d_off, n_seqs, n_bases, d_pa, d_pb, P, k, d_dist, d_cigar=None, cigar_cap=0,
                                d_cigar_off=None):
        total = C.c_int64(0)
        self._check(self._L.isof_edit_distance_pairs_dev(self._h, d_cat, d_off, n_seqs, n_bases, d_pa, d_pb, P, int(k), d_dist,
                                                         d_cigar, cigar_cap, d_cigar_off, C.byref(total)))
        return int(total.value)


def poa_consensus(seqs, match=5, mismatch=-4, gap=-2):
    """Consensus string of a list of sequences (isof_poa_consensus, host helper; spoa's -l 0 -r 0 -g -2 scoring)."""
    L = load()
    cat, off = pack_sequences(seqs)
    cap = int(2 * max((len(s) for s in seqs), default=0) + 64)
    n = C.c_int64(0)
    while True:
        out = np.empty(max(cap, 1), np.uint8)
        rc = L.isof_poa_consensus(cat.ctypes.data, off.ctypes.data, len(seqs), match, mismatch, gap, out.ctypes.data, cap,
                                  C.byref(n))
        if rc == ERR_CAPACITY:
            cap = int(n.value)
            continue
        if rc != ISOF_OK:
            raise IsofError(rc, "isof_poa_consensus failed")
        return out[:n.value].tobytes().decode("latin-1")


def wis_picks(read_rec_off, rec_p1, rec_p2, rec_count, k):
    """(pick_rec int32[], pick_off int64[n_reads+1]): isof_wis_picks over a whole batch (host helper)."""
    L = load()
    read_rec_off = np.ascontiguousarray(read_rec_off, dtype=np.int64)
    rec_p1 = np.ascontiguousarray(rec_p1, dtype=np.int32)
    rec_p2 = np.ascontiguousarray(rec_p2, dtype=np.int32)
    rec_count = np.ascontiguousarray(rec_count, dtype=np.int32)
    n_reads = len(read_rec_off) - 1
    cap = max(int(np.count_nonzero(rec_count > 0)), 1)
    picks = np.empty(cap, np.int32)
    off = np.zeros(n_reads + 1, np.int64)
    n = C.c_int64(0)
    rc = L.isof_wis_picks(n_reads, read_rec_off.ctypes.data, rec_p1.ctypes.data, rec_p2.ctypes.data, rec_count.ctypes.data,
                          int(k), picks.ctypes.data, cap, off.ctypes.data, C.byref(n))
    if rc == ERR_RANGE:
        raise IndexError("list index out of range")          # what the reference's fill_p2 does on such input
    if rc != ISOF_OK:
        raise IsofError(rc, "isof_wis_picks failed")
    return picks[:n.value], off


def span_instances(pick_rec, arrays, read_id, delta_len):
    """(inst uint32[], inst_off int64[n_picks+1]) for the picked records (isof_span_instances, host helper)."""
    L = load()
    pick_rec = np.ascontiguousarray(pick_rec, dtype=np.int32)
    a = {k_: np.ascontiguousarray(arrays[k_], dtype=np.int32) for k_ in ("rec_read", "rec_p1", "rec_p2", "rec_count", "rec_bucket",
                                                                        "sorted_rec", "bucket_off")}
    read_id = np.ascontiguousarray(read_id, dtype=np.int64)
    inst_off = np.zeros(len(pick_rec) + 1, np.int64)
    np.cumsum(3 * a["rec_count"][pick_rec].astype(np.int64), out=inst_off[1:])
    inst = np.empty(max(int(inst_off[-1]), 1), np.uint32)
    rc = L.isof_span_instances(len(pick_rec), pick_rec.ctypes.data, a["rec_read"].ctypes.data, a["rec_p1"].ctypes.data,
                               a["rec_p2"].ctypes.data, a["rec_bucket"].ctypes.data, a["sorted_rec"].ctypes.data,
                               a["bucket_off"].ctypes.data, read_id.ctypes.data, int(delta_len), inst_off.ctypes.data,
                               inst.ctypes.data)
    if rc != ISOF_OK:
        raise IsofError(rc, "isof_span_instances failed")
    return inst[:inst_off[-1]], inst_off


_default_ctx = {}


def default_context(device=None):
    """Process-wide context per device (created on first use)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("ISOF_DEVICE", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
