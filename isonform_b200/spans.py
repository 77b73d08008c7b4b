"""Host-side mirror of the minimizer-pair database / span-search seams, backed by kernel (c).

    get_minimizer_combinations_database(M, k, x_low, x_high, reads)            main:174-212
    minimizers_comb_iterator(minimizers, k, x_low, x_high)                     main:215-224
    find_most_supported_span(r_id, m1, p1, m1_curr_spans, db, all_intervals,
                             k_size, delta_len)                                main:308-338
    solve_WIS / get_intervals_to_correct                                       main:227-272

`SpanIndex` is built once per batch (one library call: sort-based group-by + all-pairs span
comparison on the GPU).  It serves the per-read interval lists in the reference's order, materialises
instance arrays (``array('I')`` of (r_id, p1, p2) triples) only when asked, and exposes a
dict-compatible view of the pair database for callers that still index ``db[m1][m2]``.
"""
from array import array

import numpy as np

from ._lib import default_context, pack_sequences


def minimizers_comb_iterator(minimizers, k, x_low, x_high):
    """main:215-224 (host form; the GPU emits the same pairs in the same order)."""
    n = len(minimizers)
    for i in range(n - 1):
        m1, p1 = minimizers[i]
        out = []
        for j in range(i + 1, n):
            m2, p2 = minimizers[j]
            d = p2 - p1
            if x_low < d <= x_high:
                out.append((m2, p2))
            elif d > x_high:
                break
        yield (m1, p1), out[::-1]


class SpanIndex:
    def __init__(self, reads, M, k, x_low, x_high, delta_len, ctx=None, _arrays=None, ids=None):
        """reads: {r_id: (acc, seq, qual)} in batch order; M: {r_id: [(kmer, pos)]} as returned by
        get_minimizers_and_positions (same key order).  Internal callers that already hold the library's span arrays
        pass them as `_arrays` with the read ids in `ids` (M is then unused)."""
        self.k, self.x_low, self.x_high, self.delta_len = k, x_low, x_high, delta_len
        self.ids = list(M) if ids is None else list(ids)
        self.id_of = np.array(self.ids, dtype=np.int64)
        self.row_of = {r: i for i, r in enumerate(self.ids)}
        self.reads = reads
        if _arrays is None:
            ctx = ctx or default_context()
            seqs = [reads[r][1] for r in self.ids]
            cat, off = pack_sequences(seqs)
            counts = [len(M[r]) for r in self.ids]
            min_off = np.zeros(len(self.ids) + 1, dtype=np.int64)
            np.cumsum(counts, out=min_off[1:])
            min_pos = np.fromiter((p for r in self.ids for _, p in M[r]), dtype=np.int32, count=int(min_off[-1]))
            _arrays = ctx.spans_batch(cat, off, min_pos, min_off, k, x_low, x_high, delta_len)
        self.__dict__.update(_arrays)
        self.n_rec = len(self.rec_read)
        # records are emitted in read order: per-read slices
        self.read_rec_off = np.searchsorted(self.rec_read, np.arange(len(self.ids) + 1), side="left")
        self.rec_span = self.rec_p2 - self.rec_p1
        self._by_anchor = None

    # ------------------------------------------------------------------ instances / intervals
    def instances(self, g):
        """The `seqs` array find_most_supported_span builds for record g (main:324-337)."""
        b = self.rec_bucket[g]
        members = self.sorted_rec[self.bucket_off[b]:self.bucket_off[b + 1]]
        keep = (self.rec_read[members] != self.rec_read[g]) & (np.abs(self.rec_span[members] - self.rec_span[g]) < self.delta_len)
        m = members[keep]
        out = np.empty((len(m) + 1, 3), dtype=np.uint32)
        out[0] = (self.id_of[self.rec_read[g]], self.rec_p1[g], self.rec_p2[g])
        out[1:, 0] = self.id_of[self.rec_read[m]]
        out[1:, 1] = self.rec_p1[m]
        out[1:, 2] = self.rec_p2[m]
        res = array("I")
        if res.itemsize == 4:
            res.frombytes(out.tobytes())                 # same values, without the per-element Python ints
        else:                                            # pragma: no cover  (platforms where C unsigned int is not 32 bit)
            res.extend(out.ravel().tolist())
        return res

    def interval_records(self, r_id):
        """Record indices of read r_id that yield an interval, in the reference's order."""
        i = self.row_of[r_id]
        g = np.arange(self.read_rec_off[i], self.read_rec_off[i + 1])
        return g[self.rec_count[g] > 0]

    def intervals(self, r_id, with_instances=True):
        """all_intervals of main:417-480 for one read: [(start, stop, count, instances)]."""
        out = []
        for g in self.interval_records(r_id):
            inst = self.instances(g) if with_instances else None
            out.append((int(self.rec_p1[g]) + self.k, int(self.rec_p2[g]), int(self.rec_count[g]), inst))
        return out

    def find_most_supported_span(self, r_id, m1, p1, m1_curr_spans, minimizer_combinations_database, all_intervals,
                                 k_size, delta_len):
        """Reference signature (main:308); `minimizer_combinations_database` is ignored -- the GPU
        result already holds every bucket."""
        if self._by_anchor is None:
            self._by_anchor = {}
            for g in range(self.n_rec):
                self._by_anchor.setdefault((int(self.rec_read[g]), int(self.rec_p1[g])), []).append(g)
        assert delta_len == self.delta_len and k_size == self.k
        wanted = {p2 for _, p2 in m1_curr_spans}
        for g in self._by_anchor.get((self.row_of[r_id], p1), []):
            if self.rec_count[g] > 0 and int(self.rec_p2[g]) in wanted:
                all_intervals.append((p1 + k_size, int(self.rec_p2[g]), int(self.rec_count[g]), self.instances(g)))

    # ------------------------------------------------------------------ database view
    def database(self):
        """{m1: {m2: array('I')}} with the reference's content and key order (main:174-212)."""
        nb = len(self.bucket_off) - 1
        if nb == 0:
            return {}
        first_rec = self.sorted_rec[self.bucket_off[:-1]]            # smallest record index of each bucket
        # The reference's outer dict gets key m1 at the FIRST record with that anchor k-mer, even if that
        # record's bucket is deleted later; the inner dict gets m2 at the bucket's first record.
        m1_of = []
        m1_first = {}
        for b in range(nb):
            g0 = first_rec[b]
            seq = self.reads[self.ids[self.rec_read[g0]]][1]
            m1 = seq[self.rec_p1[g0]:self.rec_p1[g0] + self.k]
            m1_of.append(m1)
            if m1 not in m1_first or g0 < m1_first[m1]:
                m1_first[m1] = g0
        order = sorted(range(nb), key=lambda b: (m1_first[m1_of[b]], first_rec[b]))
        db = {}
        for b in order:
            lo, hi = self.bucket_off[b], self.bucket_off[b + 1]
            if hi - lo < 2 or self.bucket_dropped[b]:
                continue
            members = self.sorted_rec[lo:hi]
            g0 = members[0]
            seq = self.reads[self.ids[self.rec_read[g0]]][1]
            m2 = seq[self.rec_p2[g0]:self.rec_p2[g0] + self.k]
            flat = np.stack([self.id_of[self.rec_read[members]], self.rec_p1[members], self.rec_p2[members]], axis=1)
            db.setdefault(m1_of[b], {})[m2] = array("I", flat.ravel().tolist())
        return db


def span_indexes_multi(batches, Ms, k, x_low, x_high, delta_len, ctx=None):
    """Super-batched form: one library call (isof_spans_multi) for many independent batches.
    batches: list of {r_id: (acc, seq, qual)}; Ms: the matching minimizer dicts.  Returns one SpanIndex per
    batch, identical to building them one by one."""
    ctx = ctx or default_context()
    seqs, min_pos_parts, counts, batch_read_off = [], [], [], [0]
    for reads, M in zip(batches, Ms):
        for r in M:
            seqs.append(reads[r][1])
            counts.append(len(M[r]))
            min_pos_parts.extend(p for _, p in M[r])
        batch_read_off.append(len(seqs))
    cat, off = pack_sequences(seqs)
    min_off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(counts, out=min_off[1:])
    min_pos = np.asarray(min_pos_parts, dtype=np.int32)
    res = ctx.spans_batch(cat, off, min_pos, min_off, k, x_low, x_high, delta_len,
                          batch_read_off=np.asarray(batch_read_off, dtype=np.int32) if len(batches) > 1 else None)
    out = []
    rec_lo_all = np.searchsorted(res["rec_read"], batch_read_off, side="left")
    # buckets are grouped by batch (most significant sort key): bucket b belongs to the batch of its first record
    first_rec = res["sorted_rec"][res["bucket_off"][:-1]] if len(res["bucket_off"]) > 1 else np.zeros(0, np.int32)
    bucket_batch = np.searchsorted(np.asarray(batch_read_off), res["rec_read"][first_rec], side="right") - 1
    for b, (reads, M) in enumerate(zip(batches, Ms)):
        lo, hi = int(rec_lo_all[b]), int(rec_lo_all[b + 1])
        bsel = np.nonzero(bucket_batch == b)[0]
        if len(bsel):
            b_lo, b_hi = int(bsel[0]), int(bsel[-1]) + 1
            assert np.array_equal(bsel, np.arange(b_lo, b_hi))
            s_lo, s_hi = int(res["bucket_off"][b_lo]), int(res["bucket_off"][b_hi])
        else:
            b_lo = b_hi = 0
            s_lo = s_hi = 0
        arrays = {"rec_read": res["rec_read"][lo:hi] - batch_read_off[b], "rec_p1": res["rec_p1"][lo:hi],
                  "rec_p2": res["rec_p2"][lo:hi], "rec_count": res["rec_count"][lo:hi],
                  "rec_bucket": res["rec_bucket"][lo:hi] - b_lo, "sorted_rec": res["sorted_rec"][s_lo:s_hi] - lo,
                  "bucket_off": res["bucket_off"][b_lo:b_hi + 1] - s_lo, "bucket_dropped": res["bucket_dropped"][b_lo:b_hi]}
        out.append(SpanIndex(reads, M, k, x_low, x_high, delta_len, _arrays=arrays))
    return out


class LazySpanDB:
    """What the drop-in get_minimizer_combinations_database returns: the reference only ever hands
    this object back to find_most_supported_span (main:479), which is where delta_len becomes known,
    so the GPU index is built on first use.  Indexing it like the reference's dict-of-dicts
    (``db[m1][m2]`` -> array('I')) materialises the view."""

    def __init__(self, M, k, x_low, x_high, reads, ctx=None):
        self.M, self.k, self.x_low, self.x_high, self.reads, self.ctx = M, k, x_low, x_high, reads, ctx
        self._index = {}
        self._view = None

    def index(self, delta_len):
        if delta_len not in self._index:
            self._index[delta_len] = SpanIndex(self.reads, self.M, self.k, self.x_low, self.x_high, delta_len, self.ctx)
        return self._index[delta_len]

    def view(self):
        if self._view is None:
            idx = next(iter(self._index.values())) if self._index else self.index(5)
            self._view = _DB(idx.database())
        return self._view

    def __getitem__(self, m1):
        return self.view()[m1]

    def __iter__(self):
        return iter(self.view())

    def keys(self):
        return self.view().keys()

    def __len__(self):
        return len(self.view())


def get_minimizer_combinations_database(M, k, x_low, x_high, reads, ctx=None):
    """Drop-in for main:174-212 (same positional arguments)."""
    return LazySpanDB(M, k, x_low, x_high, reads, ctx)


def find_most_supported_span(r_id, m1, p1, m1_curr_spans, minimizer_combinations_database, all_intervals, k_size, delta_len):
    """Drop-in for main:308-338 when `minimizer_combinations_database` is a LazySpanDB."""
    minimizer_combinations_database.index(delta_len).find_most_supported_span(
        r_id, m1, p1, m1_curr_spans, minimizer_combinations_database, all_intervals, k_size, delta_len)


class _DB(dict):
    """dict with defaultdict-like misses (the reference indexes absent keys, main:321)."""

    def __missing__(self, key):
        return _Inner()


class _Inner(dict):
    def __missing__(self, key):
        return array("I")


# ----------------------------------------------------------------------------------------------
# weighted interval scheduling (host; float tie-breaks are the reference's, main:227-272)
# ----------------------------------------------------------------------------------------------
def solve_WIS(all_intervals_sorted_by_finish):
    n = len(all_intervals_sorted_by_finish)
    starts = np.fromiter((iv[0] for iv in all_intervals_sorted_by_finish), dtype=np.int64, count=n)
    stops = np.fromiter((iv[1] for iv in all_intervals_sorted_by_finish), dtype=np.int64, count=n)
    ws = np.fromiter((iv[2] for iv in all_intervals_sorted_by_finish), dtype=np.int64, count=n)
    return _solve_wis_arrays(starts, stops, ws)


def _solve_wis_arrays(starts, stops, ws):
    n = len(starts)
    valid = np.nonzero(starts < stops)[0]
    if int(starts.max(initial=0)) > int(stops[-1]):
        raise IndexError("list index out of range")          # the reference indexes past its coordinate table
    # p[j] = index of the last valid interval with stop <= start_j (0 when none) -- main:227-239
    pos = np.searchsorted(stops[valid], starts, side="right") - 1
    p = np.where(pos >= 0, valid[np.maximum(pos, 0)], 0).tolist()
    v = ((ws - 1) * (stops - starts + 0.0001)).tolist()
    opt = [0] * (n + 1)
    for j in range(1, n + 1):
        a = v[j - 1] + opt[p[j - 1]]
        b = opt[j - 1]
        opt[j] = a if a > b else b          # max(): first argument wins ties, same value either way
    chosen = []
    j = n
    while j > 0:
        if v[j - 1] + opt[p[j - 1]] > opt[j - 1]:
            chosen.append(j - 1)
            j = p[j - 1]
        else:
            j -= 1
    return chosen


def batch_intervals_for_graph(index, reads):
    """The per-read loop of main:417-514 (with --exact): returns
    (all_intervals_for_graph {graph_id: [(start, stop, count, instances)]}, new_all_reads, skipped r_ids).
    WIS for all reads and the instance arrays of the picked intervals come from two calls into the library's host
    helpers (csrc/host_wis.cpp); what is left here is building the Python objects the reference's graph code takes."""
    from . import _lib
    picks, pick_off = _lib.wis_picks(index.read_rec_off, index.rec_p1, index.rec_p2, index.rec_count, index.k)
    inst, inst_off = _lib.span_instances(picks, index.__dict__, index.id_of, index.delta_len)
    raw = memoryview(inst.tobytes())
    starts = (index.rec_p1[picks].astype(np.int64) + index.k).tolist()
    stops = index.rec_p2[picks].tolist()
    counts = index.rec_count[picks].tolist()
    ioff = (4 * inst_off).tolist()
    poff = pick_off.tolist()
    all_intervals_for_graph, new_all_reads, skipped = {}, {}, []
    graph_id = 1
    for r_id in sorted(reads):
        i = index.row_of.get(r_id)
        lo, hi = (poff[i], poff[i + 1]) if i is not None else (0, 0)
        if hi == lo:
            skipped.append(r_id)
            continue
        out = []
        for x in range(lo, hi):
            arr = array("I")
            arr.frombytes(raw[ioff[x]:ioff[x + 1]])
            out.append((starts[x], stops[x], counts[x], arr))
        all_intervals_for_graph[graph_id] = out
        new_all_reads[graph_id] = reads[r_id]
        graph_id += 1
    return all_intervals_for_graph, new_all_reads, skipped


def batch_intervals_for_graph_py(index, reads):
    """The same in numpy / Python (kept as the cross-check of the C++ helpers in the CPU tests)."""
    all_intervals_for_graph, new_all_reads, skipped = {}, {}, []
    graph_id = 1
    for r_id in sorted(reads):
        g = index.interval_records(r_id)
        if len(g) == 0:
            skipped.append(r_id)
            continue
        starts = index.rec_p1[g].astype(np.int64) + index.k
        stops = index.rec_p2[g].astype(np.int64)
        order = np.argsort(stops, kind="stable")               # list.sort(key=stop) is stable (main:498)
        chosen = _solve_wis_arrays(starts[order], stops[order], index.rec_count[g][order].astype(np.int64))[::-1]
        picks = [(int(starts[order[j]]), int(stops[order[j]]), int(index.rec_count[g[order[j]]]), index.instances(g[order[j]]))
                 for j in chosen]
        if picks:
            all_intervals_for_graph[graph_id] = picks
            new_all_reads[graph_id] = reads[r_id]
            graph_id += 1
        else:
            skipped.append(r_id)
    return all_intervals_for_graph, new_all_reads, skipped
