"""Test double for isonform_b200._lib.Context backed by the CPU oracle.

TEST INFRASTRUCTURE ONLY: lets the host-side drivers (merge.py, spans.py, patch.py, runner.py) be
exercised on a machine without a GPU.  The product never imports this (nor the oracle)."""
import numpy as np

import oracle


class FakeContext:
    def __init__(self):
        self.calls = {"minimizers": 0, "spans": 0, "align": 0, "pairs": 0}

    # ---- (a)
    def minimizers_batch(self, cat, off, k, w_size):
        self.calls["minimizers"] += 1
        return oracle.minimizers_batch(cat, off, k, w_size, threads=1)

    # ---- (c): same outputs as isof_spans_batch, computed the slow way
    def spans_batch(self, cat, off, min_pos, min_off, k, x_low, x_high, delta_len, batch_read_off=None):
        self.calls["spans"] += 1
        if batch_read_off is not None and len(batch_read_off) > 2:
            # super-batched form: per batch, then concatenated with global record / bucket numbering
            parts, rec_base, sort_base = [], 0, 0
            off = np.asarray(off); min_off = np.asarray(min_off)
            for b in range(len(batch_read_off) - 1):
                r0, r1 = int(batch_read_off[b]), int(batch_read_off[b + 1])
                sub = self._spans_one(cat[off[r0]:off[r1]], off[r0:r1 + 1] - off[r0], min_pos[min_off[r0]:min_off[r1]],
                                      min_off[r0:r1 + 1] - min_off[r0], k, x_low, x_high, delta_len)
                sub["rec_read"] = sub["rec_read"] + r0
                sub["rec_bucket"] = sub["rec_bucket"] + sum(len(p["bucket_dropped"]) for p in parts)
                sub["sorted_rec"] = sub["sorted_rec"] + rec_base
                sub["bucket_off"] = sub["bucket_off"] + sort_base
                rec_base += len(sub["rec_read"]); sort_base += len(sub["sorted_rec"])
                parts.append(sub)
            out = {key: np.concatenate([p[key] for p in parts]) for key in ("rec_read", "rec_p1", "rec_p2", "rec_count",
                                                                          "rec_bucket", "sorted_rec", "bucket_dropped")}
            out["bucket_off"] = np.concatenate([p["bucket_off"][:-1] for p in parts] + [np.array([sort_base], np.int32)]).astype(np.int32)
            return out
        return self._spans_one(cat, off, min_pos, min_off, k, x_low, x_high, delta_len)

    def _spans_one(self, cat, off, min_pos, min_off, k, x_low, x_high, delta_len):
        cat = np.asarray(cat, dtype=np.uint8)
        R = len(off) - 1
        rec = []                                   # (read, p1, p2, key)
        for r in range(R):
            seq = bytes(cat[off[r]:off[r + 1]])
            pos = [int(p) for p in min_pos[min_off[r]:min_off[r + 1]]]
            for i in range(len(pos) - 1):
                part = []
                for j in range(i + 1, len(pos)):
                    d = pos[j] - pos[i]
                    if x_low < d <= x_high:
                        part.append(j)
                    elif d > x_high:
                        break
                for j in reversed(part):
                    rec.append((r, pos[i], pos[j], (seq[pos[i]:pos[i] + k], seq[pos[j]:pos[j] + k])))
        G = len(rec)
        buckets = {}
        for g, (_r, _p1, _p2, key) in enumerate(rec):
            buckets.setdefault(key, []).append(g)
        keys = sorted(buckets, key=lambda kk: kk)          # any deterministic bucket order will do
        rec_bucket = np.zeros(G, np.int32)
        sorted_rec, bucket_off, dropped = [], [0], []
        polya = b"A" * k
        for b, key in enumerate(keys):
            for g in buckets[key]:
                rec_bucket[g] = b
            sorted_rec += buckets[key]
            bucket_off.append(len(sorted_rec))
            dropped.append(1 if (key == (polya, polya) or len(buckets[key]) > 3 * R) else 0)
        rec_read = np.array([x[0] for x in rec], np.int32)
        rec_p1 = np.array([x[1] for x in rec], np.int32)
        rec_p2 = np.array([x[2] for x in rec], np.int32)
        rec_count = np.zeros(G, np.int32)
        for g in range(G):
            mem = buckets[rec[g][3]]
            if dropped[rec_bucket[g]] or len(mem) < 3:
                continue
            span = rec[g][2] - rec[g][1]
            rec_count[g] = 1 + sum(1 for h in mem if rec[h][0] != rec[g][0] and abs(span - (rec[h][2] - rec[h][1])) < delta_len)
        return {"rec_read": rec_read, "rec_p1": rec_p1, "rec_p2": rec_p2, "rec_count": rec_count, "rec_bucket": rec_bucket,
                "sorted_rec": np.array(sorted_rec, np.int32), "bucket_off": np.array(bucket_off, np.int32),
                "bucket_dropped": np.array(dropped, np.uint8)}

    # ---- (b)
    def sg_align_pairs(self, cat, off, pair_a, pair_b, match=2, mismatch=-2, gap_open=12, gap_ext=1, want_cigar=True,
                       features_delta_len=None, cigar_cap=None):
        self.calls["align"] += 1
        self.calls["pairs"] += len(pair_a)
        cat = np.asarray(cat, dtype=np.uint8)
        P = len(pair_a)
        score = np.zeros(P, np.int32); eq = np.zeros(P, np.int32); er = np.zeros(P, np.int32)
        coff = np.zeros(P + 1, np.int64)
        runs_all, feats = [], (np.zeros((P, 12), np.int32) if features_delta_len is not None else None)
        for p in range(P):
            s1 = bytes(cat[off[pair_a[p]]:off[pair_a[p] + 1]]).decode("latin-1")
            s2 = bytes(cat[off[pair_b[p]]:off[pair_b[p] + 1]]).decode("latin-1")
            sc, q, r, runs = oracle.sg_align(s1, s2, match, mismatch, gap_open, gap_ext)
            score[p], eq[p], er[p] = sc, q, r
            runs_all.append(runs)
            coff[p + 1] = coff[p] + len(runs)
            if feats is not None:
                tup = oracle.cigar_runs_to_tuples(runs)
                a1, a2 = oracle.cigar_to_seq(tup, s1, s2)
                alen = sum(l for l, _ in tup)
                sm = oracle.find_first_significant_match(a1, a2, 30, 0.7)
                em = oracle.find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7)
                nm = [l for l, o in tup if o != "="]
                f = [alen, len(tup), sum(nm), max(nm + [0]), sm, em, 0, 0, 0, 0, 0, 0]
                if sm >= 0 and em >= 0:
                    f[6:12] = oracle.isoform_features(tup, features_delta_len, sm, alen - em)[:6]
                feats[p] = f
        cig = np.concatenate(runs_all).astype(np.uint32) if (want_cigar and runs_all) else (np.zeros(0, np.uint32) if want_cigar else None)
        return {"score": score, "end_q": eq, "end_r": er, "cigar_off": coff, "cigar": cig, "features": feats}

    # ---- (b') unit-cost edit distance: the oracle's Myers restatement and plain-matrix path
    def edit_distance_pairs(self, cat, off, pair_a, pair_b, k=-1, want_cigar=False):
        cat = np.asarray(cat, dtype=np.uint8)
        P = len(pair_a)
        dist = np.zeros(P, np.int32)
        coff = np.zeros(P + 1, np.int64)
        runs_all = []
        for p in range(P):
            s1 = bytes(cat[off[pair_a[p]]:off[pair_a[p] + 1]])
            s2 = bytes(cat[off[pair_b[p]]:off[pair_b[p] + 1]])
            d, runs = oracle.edit_path(s1, s2)
            if k >= 0 and d > k:
                d, runs = -1, runs[:0]
            dist[p] = d
            runs_all.append(runs)
            coff[p + 1] = coff[p] + len(runs)
        if not want_cigar:
            return {"dist": dist, "cigar_off": None, "cigar": None}
        return {"dist": dist, "cigar_off": coff, "cigar": np.concatenate(runs_all).astype(np.uint32) if runs_all else np.zeros(0, np.uint32)}
