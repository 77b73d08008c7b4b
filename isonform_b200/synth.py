"""Seeded synthetic ONT-like clusters.

Gene = concatenation of max(6, L/150) random exons of 80-220 bp; isoforms = full gene + random
exon-skip variants; read = isoform copy with iid errors at rate eps split 1/3 substitution,
1/3 deletion, 1/3 insertion, plus a polyA tail of 0-25 A.  Pure Python/numpy, deterministic per
seed; used by bench.py, the tests and scripts/make_golden.py.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_dna(rng, n):
    return _ACGT[rng.integers(0, 4, size=n)].tobytes().decode()


def make_gene(rng, gene_len):
    n_exons = max(6, gene_len // 150)
    exons = []
    total = 0
    while len(exons) < n_exons or total < gene_len:
        ln = int(rng.integers(80, 221))
        exons.append(random_dna(rng, ln))
        total += ln
        if len(exons) >= n_exons and total >= gene_len:
            break
    return exons


def make_isoforms(rng, exons, n_isoforms):
    n = len(exons)
    isoforms = [list(range(n))]
    seen = {tuple(isoforms[0])}
    guard = 0
    while len(isoforms) < n_isoforms and guard < 10000:
        guard += 1
        n_skip = int(rng.integers(1, max(2, n // 4 + 1)))
        skip = set(rng.choice(np.arange(1, n - 1), size=min(n_skip, n - 2), replace=False).tolist())
        iso = [e for e in range(n) if e not in skip]
        if tuple(iso) not in seen:
            seen.add(tuple(iso))
            isoforms.append(iso)
    return ["".join(exons[e] for e in iso) for iso in isoforms]


def mutate(rng, seq, eps):
    """iid errors at rate eps: 1/3 substitution, 1/3 deletion, 1/3 insertion."""
    if eps <= 0:
        return seq
    s = np.frombuffer(seq.encode(), dtype=np.uint8)
    n = len(s)
    u = rng.random(n)
    kind = rng.integers(0, 3, size=n)
    sub_base = rng.integers(1, 4, size=n)
    ins_base = rng.integers(0, 4, size=n)
    out = []
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    for i in range(n):
        c = int(s[i])
        if u[i] < eps:
            kd = int(kind[i])
            if kd == 0:
                out.append(int(_ACGT[(code.get(c, 0) + int(sub_base[i])) & 3]))
            elif kd == 1:
                continue
            else:
                out.append(c)
                out.append(int(_ACGT[int(ins_base[i])]))
        else:
            out.append(c)
    return bytes(out).decode()


def make_cluster(seed, gene_len, n_isoforms, n_reads, eps, polya=True, read_seed=None):
    """Returns (reads: list[(acc, seq)], isoforms: list[str], truth: list[int]).
    `seed` fixes the gene and its isoforms; `read_seed` (default: continue the same stream) re-draws only
    the reads, giving statistically identical clusters of the same gene (used for weak scaling)."""
    rng = np.random.default_rng(seed)
    exons = make_gene(rng, gene_len)
    isoforms = make_isoforms(rng, exons, n_isoforms)
    if read_seed is not None:
        rng = np.random.default_rng([seed, read_seed])
    truth = rng.integers(0, len(isoforms), size=n_reads)
    reads = []
    for r in range(n_reads):
        seq = mutate(rng, isoforms[int(truth[r])], eps)
        if polya:
            seq = seq + "A" * int(rng.integers(0, 26))
        reads.append(("read_%d_iso%d" % (r, int(truth[r])), seq))
    return reads, isoforms, [int(t) for t in truth]


# BASELINE.json configs
CONFIGS = {
    "C1": dict(gene_len=1500, n_isoforms=3, n_reads=200, eps=0.01, seed=1),
    "C2": dict(gene_len=3000, n_isoforms=10, n_reads=1000, eps=0.05, seed=2),
    "C4": dict(gene_len=10000, n_isoforms=30, n_reads=2000, eps=0.05, seed=4),
}


def make_config(name, seed=None, n_reads=None, read_seed=None):
    cfg = dict(CONFIGS[name])
    if seed is not None:
        cfg["seed"] = seed
    if n_reads is not None:
        cfg["n_reads"] = n_reads
    return make_cluster(cfg["seed"], cfg["gene_len"], cfg["n_isoforms"], cfg["n_reads"], cfg["eps"], read_seed=read_seed)


def write_fastq(path, reads):
    with open(path, "w") as fh:
        for acc, seq in reads:
            fh.write("@%s\n%s\n+\n%s\n" % (acc, seq, "I" * len(seq)))


def mutate_pair(rng, length, eps):
    """(x, mutate(x, eps)) for the config-5 pair sweep."""
    x = random_dna(rng, length)
    return x, mutate(rng, x, eps)
