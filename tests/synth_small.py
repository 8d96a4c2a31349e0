"""Small randomized graph/read generators for parity tests (pure Python, seeded).

The unitig sets are deliberately NOT compacted de Bruijn graphs: with k in 4..8 a few hundred
random bases repeat (k-1)-mers all the time, which exercises the reference's quirks: more than
4 unitigs per border (slot-4 overwrite), palindromic borders (k-1 even), hairpin unitigs, tips,
reads that leave the graph, and N characters.
"""
import random

COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def random_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, rate, with_n=0.0):
    out = []
    for c in s:
        x = rng.random()
        if x < with_n:
            out.append("N")
        elif x < with_n + rate:
            out.append(rng.choice([b for b in "ACGT" if b != c]))
        else:
            out.append(c)
    return "".join(out)


def make_graph(rng, k, genome_len=300, mean_unitig=12, n_variants=6, low_complexity=False, genome=None):
    """-> (unitigs in file order, haplotypes). Unitigs overlap their neighbours by k-1 bases.
    genome: the first haplotype, given instead of drawn."""
    alphabet = "ACGT" if not low_complexity else "AACGT"
    g1 = random_seq(rng, genome_len, alphabet) if genome is None else genome
    genome_len = len(g1)
    # second haplotype: a few substitutions
    g2 = list(g1)
    for _ in range(n_variants):
        p = rng.randrange(genome_len)
        g2[p] = rng.choice([b for b in "ACGT" if b != g2[p]])
    g2 = "".join(g2)
    unitigs = []
    for g in (g1, g2):
        pos = 0
        while pos + k <= len(g):
            span = max(1, int(rng.expovariate(1.0 / mean_unitig)))
            end = min(len(g), pos + span + k - 1)
            if len(g) - (pos + span) < k:
                end = len(g)
            u = g[pos:end]
            if len(u) >= k:
                unitigs.append(u)
            pos = end - (k - 1)
            if end == len(g):
                break
    # de-duplicate identical unitigs (both haplotypes share most), keep first occurrence
    seen, uniq = set(), []
    for u in unitigs:
        key = min(u, revcomp(u))
        if key not in seen:
            seen.add(key)
            uniq.append(u)
    uniq = [u if rng.random() < 0.5 else revcomp(u) for u in uniq]
    rng.shuffle(uniq)
    return uniq, (g1, g2)


def make_reads(rng, haplotypes, n, read_len, sub_rate=0.02, n_rate=0.0, off_graph=0.05):
    reads = []
    for _ in range(n):
        if rng.random() < off_graph:
            r = random_seq(rng, read_len)
        else:
            g = rng.choice(haplotypes)
            ln = min(read_len, len(g))
            p = rng.randrange(len(g) - ln + 1)
            r = g[p:p + ln]
            if rng.random() < 0.5:
                r = revcomp(r)
            r = mutate(rng, r, sub_rate, n_rate)
        reads.append(r)
    return reads


def write_unitigs(path, unitigs):
    with open(path, "w") as f:
        for i, u in enumerate(unitigs):
            f.write(">%d\n%s\n" % (i + 1, u))


def write_fasta(path, reads, prefix="r"):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            f.write(">%s%d\n%s\n" % (prefix, i, r))


def fuzz_case(seed):
    """One randomized small case -> dict(k, m, effort, unitigs, reads)."""
    rng = random.Random(seed)
    k = rng.choice([4, 5, 5, 6, 7, 8])
    m = rng.choice([0, 1, 2, 2, 3, 5])
    effort = rng.choice([1, 2, 2, 3, 5])
    unitigs, haps = make_graph(rng, k, genome_len=rng.choice([60, 150, 300]), mean_unitig=rng.choice([3, 8, 20]),
                               n_variants=rng.choice([0, 3, 10]), low_complexity=rng.random() < 0.3)
    read_len = rng.choice([k + 1, 12, 20, 35, 50])
    reads = make_reads(rng, haps, 40, max(read_len, k + 1), sub_rate=rng.choice([0.0, 0.02, 0.06]),
                       n_rate=rng.choice([0.0, 0.0, 0.03]))
    return dict(k=k, m=m, effort=effort, unitigs=unitigs, reads=reads)
