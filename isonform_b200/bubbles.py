"""Prefetch driver for bubble popping -- alignment call site 1.

The reference's `new_bubble_popping_routine` (modules/SimplifyGraph.py:883-1151) walks the sorted
(start, end) combinations of an iteration one by one and, between graph mutations, aligns the
consensus strings of two bubble paths with `consensus.parasail_alignment` (`align_bubble_nodes`,
SimplifyGraph.py:584-664, aligner at :644).  The aligner is pure in `(s1, s2, scores)`; what the loop
asks for is decided by the graph at the START of the iteration plus the `marked` set, which only ever
REMOVES candidates (find_paths stops at a marked node, :149-151; filter_out_if_marked, :761-790).  So
the candidates seen with an empty `marked` are a superset of what the iteration will align.

`BubblePrefetcher` hooks the one function the routine calls exactly once per iteration before the
combination loop (`generate_combinations`, :938).  After the reference has built its combinations it

  1. enumerates, with the reference's own `find_paths` / `get_consensus_positions` /
     `align_bubble_nodes` and an empty `marked`, the path pairs the loop will reach (see
     `_speculate` for which ones are certain), with the aligner seam swapped for a recorder (so
     consensus strings, spoa calls and the length gates :630-664 are the reference's own code, not
     a restatement);
  2. sends all recorded `(consensus1, consensus2)` pairs to the GPU in ONE batched call;
  3. lets the reference's loop run unmodified: every `parasail_alignment` it issues is a memo hit, a
     pair the speculation did not foresee falls back to a synchronous one-pair call (counted).

`run_spoa` (IsoformGeneration.py:143-156) is memoised by the content of its input file: the
speculation and the replay ask for the same consensus, and so do later iterations, so each distinct
read set spawns `spoa` once.  Nothing in the reference tree is modified; `uninstall()` restores the
rebound attributes.  The two bookkeeping hot spots of the same loop are replaced by equivalent, faster forms:
`generate_combinations` (one matrix product) and `find_paths` (`PathFinder`: per-node lookup tables).
"""
import collections
import itertools
import sys

import numpy as np


def generate_combinations(possible_starts, possible_ends, topo_nodes_dict, combis):
    """Drop-in for SimplifyGraph.generate_combinations (:110-119), one of the two bookkeeping hot spots SURVEY.md
    section 8f names (4 s per 1000-read batch: |starts| x |ends| Python set intersections per iteration).  Same
    appended tuples in the same order; the read-support sets become rows of two 0/1 matrices and ONE integer matrix
    product counts every intersection, so only the (start, end) pairs that really share >= 2 reads and respect the
    topological order are touched in Python."""
    if not possible_starts or not possible_ends:
        return
    reads = sorted({r for _n, supp in possible_starts for r in supp} | {r for _n, supp in possible_ends for r in supp})
    col = {r: i for i, r in enumerate(reads)}
    S = np.zeros((len(possible_starts), len(reads)), dtype=np.float32)
    E = np.zeros((len(possible_ends), len(reads)), dtype=np.float32)
    for i, (_n, supp) in enumerate(possible_starts):
        S[i, [col[r] for r in supp]] = 1.0
    for j, (_n, supp) in enumerate(possible_ends):
        E[j, [col[r] for r in supp]] = 1.0
    shared = S @ E.T                                       # exact: counts <= #reads << 2^24
    ts = np.array([topo_nodes_dict[n] for n, _s in possible_starts])
    te = np.array([topo_nodes_dict[n] for n, _s in possible_ends])
    ok = (shared >= 2.0) & (ts[:, None] < te[None, :])
    read_arr = np.array(reads, dtype=object)
    for i, j in zip(*np.nonzero(ok)):                      # row-major: starts outer, ends inner, as the reference loops
        inter = tuple(read_arr[np.nonzero((S[i] * E[j]) > 0)[0]].tolist())          # `reads` is sorted, so is this
        combis.append((possible_starts[i][0], possible_ends[j][0], inter))


class PathFinder:
    """Drop-in for SimplifyGraph.find_paths (:132-172), the second bookkeeping hot spot SURVEY.md section 8f names
    (432 k calls per 1000-read batch).  Same walk, same results -- the set operations that fix the order in which reads
    are popped are kept as they are -- but the step "first successor whose edge_supp LIST contains the read" (a scan
    over the out-edges with a linear list search each) becomes one dictionary lookup in a per-node table
    {read: (first such successor, frozenset(edge_supp))}, built on first use from DG's adjacency order.  The tables
    are dropped whenever the graph changes (`invalidate`, called after every linearize_bubble)."""

    def __init__(self):
        self._tables = {}
        self._graph = None

    def invalidate(self):
        self._tables.clear()

    def _table(self, DG, node):
        t = self._tables.get(node)
        if t is None:
            t = {}
            for neighbor in DG.successors(node):
                supp = DG[node][neighbor]['edge_supp']
                entry = (neighbor, frozenset(supp))
                for r in supp:
                    if r not in t:
                        t[r] = entry
            self._tables[node] = t
        return t

    def find_paths(self, DG, startnode, endnode, support, all_paths, marked):
        if DG is not self._graph:
            self._graph = DG
            self._tables.clear()
        node_support_left = set(support)
        all_supp = set(support)
        while node_support_left:
            node = startnode
            read = node_support_left.pop()
            current_node_support = node_support_left        # the reference aliases the set here (:141-142)
            current_node_support.add(read)
            visited_nodes = []
            next_found = False
            while node != endnode:
                visited_nodes.append(node)
                next_found = False
                if node in marked:
                    break
                hit = self._table(DG, node).get(read)
                if hit is None:
                    break
                node = hit[0]
                current_node_support = current_node_support.intersection(hit[1])
                next_found = True
            if current_node_support and next_found:
                node_support_left -= current_node_support
                all_paths.append((visited_nodes, tuple(sorted(current_node_support)), all_supp - current_node_support))
            else:
                break


class _NeedsSpoa(Exception):
    """Raised inside a speculative consensus whose spoa result is not memoised yet."""


class BubblePrefetcher:
    def __init__(self, memo, stats=None):
        self.memo = memo
        self.stats = stats if stats is not None else {}
        self.stats.update({"iterations": 0, "speculated_pairs": 0, "prefetch_calls": 0, "prefetched": 0,
                           "spoa_calls": 0, "spoa_hits": 0, "certain": 0, "uncertain": 0, "skipped_for_spoa": 0})
        self._spoa_cache = {}
        self._orig = {}
        self._in_speculation = False
        self._no_new_spoa = False
        self.paths = PathFinder()
        self.check_paths = False

    # ------------------------------------------------------------------ wiring
    def install(self):
        from modules import IsoformGeneration, SimplifyGraph        # the reference tree
        self._sg, self._ig = SimplifyGraph, IsoformGeneration
        self._orig = {"generate_combinations": SimplifyGraph.generate_combinations,
                      "run_spoa": IsoformGeneration.run_spoa, "find_paths": SimplifyGraph.find_paths,
                      "linearize_bubble": SimplifyGraph.linearize_bubble}
        SimplifyGraph.generate_combinations = self._generate_combinations
        IsoformGeneration.run_spoa = self._run_spoa
        SimplifyGraph.find_paths = self._find_paths
        SimplifyGraph.linearize_bubble = self._linearize_bubble
        return self

    def uninstall(self):
        if self._orig:
            self._sg.generate_combinations = self._orig["generate_combinations"]
            self._ig.run_spoa = self._orig["run_spoa"]
            self._sg.find_paths = self._orig["find_paths"]
            self._sg.linearize_bubble = self._orig["linearize_bubble"]
            self._orig = {}

    # ------------------------------------------------------------------ SimplifyGraph.find_paths / linearize_bubble
    def _find_paths(self, DG, startnode, endnode, support, all_paths, marked):
        before = len(all_paths)
        self.paths.find_paths(DG, startnode, endnode, support, all_paths, marked)
        if self.check_paths:                      # test mode: the reference's own function must agree, call by call
            want = []
            self._orig["find_paths"](DG, startnode, endnode, support, want, marked)
            assert all_paths[before:] == want, (startnode, endnode)
            self.stats["find_paths_checked"] = self.stats.get("find_paths_checked", 0) + 1

    def _linearize_bubble(self, *args, **kwargs):
        try:
            return self._orig["linearize_bubble"](*args, **kwargs)
        finally:
            self.paths.invalidate()               # the only place the routine rewires edges (:291, :507, :514)

    # ------------------------------------------------------------------ IsoformGeneration.run_spoa, memoised
    def _run_spoa(self, reads, spoa_out_file):
        with open(reads, "r") as fh:
            key = fh.read()
        hit = self._spoa_cache.get(key)
        if hit is not None:
            self.stats["spoa_hits"] += 1
            return hit
        if self._no_new_spoa:
            raise _NeedsSpoa()
        self.stats["spoa_calls"] += 1
        con = self._orig["run_spoa"](reads, spoa_out_file)
        self._spoa_cache[key] = con
        return con

    # ------------------------------------------------------------------ the per-iteration hook
    def _generate_combinations(self, possible_starts, possible_ends, topo_nodes_dict, combis):
        generate_combinations(possible_starts, possible_ends, topo_nodes_dict, combis)
        frame = sys._getframe(1)
        if frame.f_code.co_name != "new_bubble_popping_routine" or self._in_speculation:
            return
        loc = frame.f_locals
        try:
            DG, all_reads, work_dir = loc["DG"], loc["all_reads"], loc["work_dir"]
            k_size, delta_len = loc["k_size"], loc["delta_len"]
        except KeyError:                 # a reference version with other local names: no speculation, still correct
            return
        self.stats["iterations"] += 1
        self._in_speculation = True
        try:
            pairs = self._speculate(DG, all_reads, work_dir, k_size, delta_len, combis,
                                    loc.get("not_viable_global", ()), loc.get("not_viable_multibubble", ()),
                                    loc.get("multi_consensuses", {}), topo_nodes_dict)
        finally:
            self._in_speculation = False
        self.stats["speculated_pairs"] += len(pairs)
        if pairs:
            n = self.memo.prefetch_alignments(pairs, 2, -8, 12, 1)        # scores of SimplifyGraph.py:644-649
            if n:
                self.stats["prefetch_calls"] += 1
                self.stats["prefetched"] += n

    def _speculate(self, DG, all_reads, work_dir, k_size, delta_len, combis, not_viable_global, not_viable_multibubble,
                   multi_consensuses, topo_nodes_dict):
        """The (consensus1, consensus2) pairs the combination loop (:951-1146) will hand to the aligner this iteration.

        A combination is evaluated by the loop exactly as it looks now iff none of its nodes is touched by an earlier
        combination of the sorted order (:946): `marked` holds inner nodes of bubbles popped earlier in the
        iteration, `direct_combis` their (start, end), and linearize_bubble rewires edges among the popped bubble's
        own nodes only.  Such CERTAIN combinations are speculated in full -- consensus strings included, so the
        `spoa` spawns they cost are the ones the replay needs anyway (memoised).  A combination that shares a node
        with an earlier candidate may be cut short by `marked`; it is speculated only when its consensus strings need
        no NEW spoa run (paths of <= 2 reads, generate_consensus_path :562-573, or read sets whose consensus is
        already memoised), so speculation never adds a subprocess spawn.  Whatever the replay asks beyond that is a synchronous one-pair call."""
        sg = self._sg
        recorded = []

        def recorder(s1, s2, match_score=2, mismatch_penalty=-2, opening_penalty=12, gap_ext=1):
            recorded.append((s1, s2))
            return "", "", "", [(1 << 30, "X")], 0          # never poppable; the result of a speculative call is dropped

        real_seam = sg.consensus.parasail_alignment
        sg.consensus.parasail_alignment = recorder
        no_marks = frozenset()
        touched = set()
        order = sorted((c for c in combis if c not in not_viable_global),
                       key=lambda x: topo_nodes_dict[x[1]] - topo_nodes_dict[x[0]])           # :939-946
        try:
            for combination in order:
                start, end, support = combination
                all_paths = []
                sg.find_paths(DG, start, end, support, all_paths, no_marks)
                if len(all_paths) < 2:
                    continue
                nodes = {start, end}
                for path in all_paths:
                    nodes.update(path[0])
                certain = nodes.isdisjoint(touched)
                touched |= nodes
                self.stats["certain" if certain else "uncertain"] += 1
                mega = len(all_paths) > 2
                for p1, p2 in itertools.combinations(all_paths, 2):
                    this_combi = (start, end, tuple(sorted(set(p1[1]) | set(p2[1]))))
                    if this_combi in not_viable_multibubble:
                        continue
                    if set(p1[0][1:]).intersection(p2[0][1:]):
                        continue
                    node1 = p1[0][1] if len(p1[0]) > 1 else end
                    node2 = p2[0][1] if len(p2[0]) > 1 else end
                    if node1 == node2:
                        continue
                    infos = {node1: sg.get_consensus_positions(start, end, DG, p1[1]),
                             node2: sg.get_consensus_positions(start, end, DG, p2[1])}
                    # reads of the reference's consensus cache fall through, writes stay in the scratch front map
                    cache = collections.ChainMap({}, multi_consensuses)
                    self._no_new_spoa = not certain
                    try:
                        sg.align_bubble_nodes(all_reads, infos, work_dir, k_size, 0, cache, mega, this_combi, delta_len)
                    except _NeedsSpoa:                            # a spawn for a pair the loop may never reach: skip
                        self.stats["skipped_for_spoa"] += 1
        finally:
            self._no_new_spoa = False
            sg.consensus.parasail_alignment = real_seam
        return recorded
