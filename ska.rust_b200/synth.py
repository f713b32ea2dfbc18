"""Deterministic synthetic inputs: star-model bacterial genomes.

ancestor = i.i.d. uniform {A,C,G,T} of length G; sample i = ancestor with every position independently
substituted (prob `snp_rate`) by one of the three other bases.  numpy PCG64 streams seeded from
(seed, i), so any subset of samples can be regenerated independently (used for multi-GPU sharding).
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


def write_fasta(path, name, seq, width=80):
    with open(path, "wb") as fh:
        fh.write(b">" + name.encode() + b"\n")
        b = seq.tobytes()
        for i in range(0, len(b), width):
            fh.write(b[i:i + width] + b"\n")
