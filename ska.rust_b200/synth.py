"""Deterministic synthetic inputs: star- and tree-model bacterial genomes.

ancestor = i.i.d. uniform {A,C,G,T} of length G.
star: sample i = ancestor with every position independently substituted (prob `snp_rate`) by one of the
three other bases.
tree: the samples are the leaves of a balanced binary tree of depth D = ceil(log2(n_total)); every branch
adds substitutions at rate `snp_rate / D`, so every leaf is `snp_rate` from the root and close relatives
share most of their SNPs (the number of distinct split k-mers grows with the number of BRANCHES, not with
n_samples x SNPs per sample as under the star model).
numpy PCG64 streams seeded from (seed, i) / (seed, depth, node), so any subset of samples can be regenerated
independently (used for multi-GPU sharding).
"""
import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def ancestor(length, seed):
    rng = np.random.Generator(np.random.PCG64([seed, 0xA11CE]))
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def sample_codes(anc, seed, i, snp_rate=0.001):
    rng = np.random.Generator(np.random.PCG64([seed, 1 + i]))
    n = len(anc)
    n_snp = rng.binomial(n, snp_rate)
    pos = rng.integers(0, n, size=n_snp)
    shift = rng.integers(1, 4, size=n_snp, dtype=np.uint8)
    out = anc.copy()
    out[pos] = (anc[pos] + shift) & 3          # duplicates in pos: last write wins, still a substitution
    return out


def star_genomes(n_samples, length, seed=0x5CA20000, snp_rate=0.001, first=0, out=None):
    """-> list of uint8 ASCII arrays (one record each).  `out`: optional preallocated (n_samples, length) uint8."""
    anc = ancestor(length, seed)
    res = []
    for j in range(n_samples):
        codes = sample_codes(anc, seed, first + j, snp_rate)
        if out is not None:
            np.take(_BASES, codes, out=out[j])
            res.append(out[j])
        else:
            res.append(_BASES[codes])
    return res


def tree_sample_codes(anc, seed, i, n_total, snp_rate=0.001):
    depth = max(1, int(np.ceil(np.log2(max(n_total, 2)))))
    out = anc.copy()
    n = len(anc)
    for d in range(1, depth + 1):
        node = i >> (depth - d)                 # the branch above node (d, node) on the path root -> leaf i
        rng = np.random.Generator(np.random.PCG64([seed, 0x7EE, d, node]))
        n_snp = rng.binomial(n, snp_rate / depth)
        pos = rng.integers(0, n, size=n_snp)
        shift = rng.integers(1, 4, size=n_snp, dtype=np.uint8)
        out[pos] = (out[pos] + shift) & 3
    return out


def tree_genomes(n_samples, length, n_total, seed=0x5CA20000, snp_rate=0.001, first=0, out=None):
    """leaves first .. first+n_samples-1 of the n_total-leaf tree; same return convention as star_genomes"""
    anc = ancestor(length, seed)
    res = []
    for j in range(n_samples):
        codes = tree_sample_codes(anc, seed, first + j, n_total, snp_rate)
        if out is not None:
            np.take(_BASES, codes, out=out[j])
            res.append(out[j])
        else:
            res.append(_BASES[codes])
    return res


def genomes(population, n_samples, length, n_total, seed=0x5CA20000, first=0, out=None):
    if population == "star":
        return star_genomes(n_samples, length, seed=seed, first=first, out=out)
    if population == "tree":
        return tree_genomes(n_samples, length, n_total, seed=seed, first=first, out=out)
    raise ValueError(f"unknown population model {population!r}")


def write_fasta(path, name, seq, width=80):
    with open(path, "wb") as fh:
        fh.write(b">" + name.encode() + b"\n")
        b = seq.tobytes()
        for i in range(0, len(b), width):
            fh.write(b[i:i + width] + b"\n")


_COMP = np.zeros(256, dtype=np.uint8)
_COMP[list(b"ACGT")] = list(b"TGCA")


def paired_reads(genome_ascii, n_pairs, read_len=150, frag_len=400, err_rate=0.005, seed=0, lowq_prob=0.5):
    """SURVEY.md 8d config-4 reads: forward read from the + strand, reverse read = reverse complement of the fragment
    end; per-base substitution errors; quality 'I' (Q40) except error bases get '+' (Q10) with prob `lowq_prob`.
    -> ((seq1, qual1, off1), (seq2, qual2, off2)) flat uint8 arrays + uint64 record offsets."""
    rng = np.random.Generator(np.random.PCG64([seed, 0xBEAD5]))
    g = np.asarray(genome_ascii, dtype=np.uint8)
    starts = rng.integers(0, len(g) - frag_len + 1, size=n_pairs)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    fwd = g[idx]
    ridx = starts[:, None] + (frag_len - 1 - np.arange(read_len))[None, :]
    rev = _COMP[g[ridx]]
    out = []
    for reads in (fwd, rev):
        reads = reads.copy()
        err = rng.random(reads.shape) < err_rate
        codes = (reads >> 1) & 3                                   # A0 C1 T2 G3
        sub = (codes + rng.integers(1, 4, size=reads.shape, dtype=np.uint8)) & 3
        reads[err] = np.frombuffer(b"ACTG", dtype=np.uint8)[sub[err]]
        qual = np.full(reads.shape, ord("I"), dtype=np.uint8)
        qual[err & (rng.random(reads.shape) < lowq_prob)] = ord("+")
        off = (np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(read_len))
        out.append((reads.reshape(-1), qual.reshape(-1), off))
    return out[0], out[1]
