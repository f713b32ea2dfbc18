"""FASTQ path on the device (quality masks, ntHash, exact emulation of the bloom + counts filter) against the
oracle's sequential KmerFilter -- reference fixtures of tests/fastq_input.rs and synthetic reads dense enough to
saturate bloom words, so that bloom FALSE POSITIVES (which change counts by one) are exercised."""
import numpy as np
import pytest

import ska_rust_b200 as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def same(g, o):
    assert g.dims() == o.dims(), (g.dims(), o.dims())
    gk, gv, gc = g.export()
    ok, ov, oc = o.export()
    assert [int(x) for x in gk] == ok
    assert np.array_equal(gv, ov) and np.array_equal(gc, oc)


def pair(gold_in, prefix):
    return [(gold_in(f"{prefix}_{i}_fwd.fastq.gz"), gold_in(f"{prefix}_{i}_rev.fastq.gz")) for i in (1, 2)]


FIXTURE_CASES = [  # tests/fastq_input.rs
    ("test", dict(k=9, min_count=2, min_qual=2)),                                   # align_fastq :12-55
    ("test", dict(k=9, min_count=1, min_qual=2)),                                   # map_fastq build :196-210
    ("test", dict(k=9, min_count=3, min_qual=2)),                                   # error_fastq :283-297
    ("test_count", dict(k=7, min_count=1)),                                         # count_check :59-80
    ("test_count", dict(k=7, min_count=3)),                                         # :84-109
    ("test_quality", dict(k=9, min_count=5, qual_filter="no-filter")),              # :300-325
    ("test_quality_base", dict(k=9, min_count=5, qual_filter="middle", min_qual=5)),   # :328-355
    ("test_error", dict(k=9, min_count=5, min_qual=2)),                             # :358-378
    ("test_quality", dict(k=9, min_count=5, min_qual=30)),                          # :381-401
    ("test_quality_base", dict(k=9, min_count=5, min_qual=5, qual_filter="strict")),   # :404-431
    ("test_quality_base", dict(k=9, min_count=5)),                                  # :434-454
]


@pytest.mark.parametrize("prefix,kw", FIXTURE_CASES)
def test_fastq_fixtures_match_oracle(gold_in, prefix, kw):
    files = pair(gold_in, prefix)
    names = [f"{prefix}_1", f"{prefix}_2"]
    kw = dict(kw)
    k = kw.pop("k")
    o = O.OracleArray.build_from_files(files, k=k, names=names, **kw)
    g = S.build_and_merge([(n, f1, f2) for n, (f1, f2) in zip(names, files)], k, True, S.QualOpts(**kw))
    same(g, o)


def test_fastq_single_strand_and_fasta_equivalence(gold_in):
    files = pair(gold_in, "test")
    g = S.build_and_merge([("a", *files[0]), ("b", *files[1])], 9, False, S.QualOpts(min_count=2, min_qual=2))
    o = O.OracleArray.build_from_files(files, k=9, rc=False, names=["a", "b"], min_count=2, min_qual=2)
    same(g, o)
    # tests/fastq_input.rs:12-55: perfect reads give the FASTA alignment
    import io
    gq = S.build_and_merge([("a", *files[0]), ("b", *files[1])], 9, True, S.QualOpts(min_count=2, min_qual=2))
    gf = S.build_and_merge([("a", gold_in("test_1.fa"), None), ("b", gold_in("test_2.fa"), None)], 9, True)
    b1, b2 = io.BytesIO(), io.BytesIO()
    S.api.align(gq, b1)
    S.api.align(gf, b2)
    assert O.var_hash(b1.getvalue()) == O.var_hash(b2.getvalue())


def make_reads_sample(genome, n_pairs, err, seed):
    from ska_rust_b200 import synth
    (s1, q1, o1), (s2, q2, o2) = synth.paired_reads(genome, n_pairs, err_rate=err, seed=seed)
    seq = np.concatenate([s1, s2])
    qual = np.concatenate([q1, q2])
    off = np.concatenate([o1, o2[1:] + o1[-1]])
    gs = S.Sample(seq, off, qual, n_rec_file1=n_pairs)
    recs1 = [(s1[int(o1[i]):int(o1[i + 1])].tobytes(), q1[int(o1[i]):int(o1[i + 1])].tobytes()) for i in range(n_pairs)]
    recs2 = [(s2[int(o2[i]):int(o2[i + 1])].tobytes(), q2[int(o2[i]):int(o2[i + 1])].tobytes()) for i in range(n_pairs)]
    return gs, O.SampleData(recs1, recs2, is_reads=True)


@pytest.mark.parametrize("min_count,qual_filter,min_qual", [(5, "strict", 20), (2, "middle", 20), (3, "no-filter", 0), (1, "strict", 20)])
def test_fastq_synthetic_small(min_count, qual_filter, min_qual):
    from ska_rust_b200 import synth
    genomes = synth.star_genomes(2, 20_000, seed=0x5CA20004, snp_rate=0.002)
    gs, os_ = zip(*[make_reads_sample(g, 3000, 0.01, 77 + i) for i, g in enumerate(genomes)])
    names = ["r0", "r1"]
    q = S.QualOpts(min_count=min_count, min_qual=min_qual, qual_filter=qual_filter)
    g = S.build_from_samples(list(gs), names, 31, True, q)
    o = O.OracleArray.build(list(os_), names, k=31, rc=True, min_count=min_count, min_qual=min_qual, qual_filter=qual_filter)
    same(g, o)


@pytest.mark.parametrize("min_count", [5, 2])
def test_fastq_bloom_false_positives_saturated(min_count):
    """~3.5e7 distinct hashes land in the 3.1e6 bloom words (>10 per 64-bit word): most first observations of late
    k-mers are bloom false positives, which shifts their count by one -- the batch emulation must agree with the
    sequential filter observation by observation."""
    from ska_rust_b200 import synth
    genome = synth.star_genomes(1, 300_000, seed=0x5CA20005)[0]
    gs, os_ = make_reads_sample(genome, 200_000, 0.02, 5)           # 6e7 bases, ~1.2e6 errors -> ~3.7e7 novel k-mers
    q = S.QualOpts(min_count=min_count, min_qual=20, qual_filter="no-filter")
    g = S.build_from_samples([gs], ["r"], 31, True, q)
    o = O.OracleArray.build([os_], ["r"], k=31, rc=True, min_count=min_count, min_qual=20, qual_filter="no-filter")
    same(g, o)
    # the effect under test must actually be present: without false positives the result differs
    assert g.dims()["n_rows"] > 300_000


@pytest.mark.parametrize("min_count", [2, 5, 65535])
def test_fastq_heavy_repeats(min_count):
    """homopolymer and tandem-repeat reads: one hash observed tens of thousands of times (the sort's oversized-bucket
    path, the binary-searched run length, and the long accept lists of min_count == 2 / the saturated u16 count)"""
    rng = np.random.default_rng(0x5CA2000A)
    L, n_pairs = 150, 1000
    genome = rng.integers(0, 4, 5000)
    def read(i):
        if i % 3 == 0:
            return np.full(L, ord("A"), dtype=np.uint8)
        if i % 3 == 1:
            return np.frombuffer((b"ACGGT" * 40)[i % 5:i % 5 + L], dtype=np.uint8)
        a = int(rng.integers(0, 5000 - L))
        return np.frombuffer(b"ACGT", dtype=np.uint8)[genome[a:a + L]]
    files = []
    for f in range(2):
        seqs = [read(i + f) for i in range(n_pairs)]
        quals = [rng.integers(25, 41, L).astype(np.uint8) + 33 for _ in range(n_pairs)]
        files.append((seqs, quals))
    seq = np.concatenate(files[0][0] + files[1][0])
    qual = np.concatenate(files[0][1] + files[1][1])
    off = np.arange(0, 2 * n_pairs + 1, dtype=np.int64) * L
    gs = S.Sample(seq, off, qual, n_rec_file1=n_pairs)
    od = O.SampleData([(s.tobytes(), q.tobytes()) for s, q in zip(*files[0])], [(s.tobytes(), q.tobytes()) for s, q in zip(*files[1])],
                      is_reads=True)
    q = S.QualOpts(min_count=min_count, min_qual=20, qual_filter="strict")
    for k in (31, 41):
        g = S.build_from_samples([gs], ["r"], k, True, q)
        o = O.OracleArray.build([od], ["r"], k=k, rc=True, min_count=min_count, min_qual=20, qual_filter="strict")
        same(g, o)
        if min_count < 65535:
            assert g.dims()["n_rows"] > 0


def test_fastq_sort_fallback_for_ambiguous_buckets(monkeypatch):
    """the 96-bit observation sort orders a bucket by (32 hash bits, position); two hashes that agree in those bits send the
    bucket to the full-width path.  With only 12 compared bits that happens in most buckets: the result must not change."""
    from ska_rust_b200 import synth
    genome = synth.star_genomes(1, 100_000, seed=0x5CA20006)[0]
    gs, os_ = make_reads_sample(genome, 120_000, 0.01, 9)            # 3.6e7 windows -> 19 prefix bits, 12 + 32 >= 64 - 19
    q = S.QualOpts(min_count=3, min_qual=20, qual_filter="strict")
    o = O.OracleArray.build([os_], ["r"], k=31, rc=True, min_count=3, min_qual=20, qual_filter="strict")
    for bits in ("12", None):
        if bits:
            monkeypatch.setenv("SKA_SORT_KH_BITS", bits)
        else:
            monkeypatch.delenv("SKA_SORT_KH_BITS")
        same(S.build_from_samples([gs], ["r"], 31, True, q), o)


def test_fastq_samples_shard_by_rows():
    """FASTQ samples through the multi-GPU "rows" path (every 'rank' builds its own reads with the count filter, the
    owners join the partial rows), emulated on one GPU with the exchange replaced by copies: equals the oracle's array"""
    import torch

    import ska_rust_b200.distributed as D
    from ska_rust_b200 import synth
    genomes = synth.star_genomes(4, 12_000, seed=0x5CA20005, snp_rate=0.003)
    gs, os_ = zip(*[make_reads_sample(g, 2500, 0.01, 91 + i) for i, g in enumerate(genomes)])
    names = [f"r{i}" for i in range(4)]
    k, q = 21, S.QualOpts(min_count=3, min_qual=20, qual_filter="strict")
    backend = D.GpuBackend()
    bounds = [(0, 1), (1, 4)]                                  # uneven ranks
    world = len(bounds)
    region_bits = min(backend.region_bits_max, 2 * (k - 1))
    local = [backend.build_local(list(gs[lo:hi]), names[lo:hi], k, True, q) for lo, hi in bounds]
    all_cum = np.stack([backend.key_cumulative(a, region_bits) for a in local])
    tensors = [backend.array_tensors(a) for a in local]
    cols = np.array([hi - lo for lo, hi in bounds], dtype=np.int64)
    col_base = np.array([lo for lo, _ in bounds], dtype=np.int64)
    parts = []
    for g in range(world):
        cuts, send, recv = D.rows_exchange_plan(all_cum, world, g)
        b = all_cum[:, cuts]
        keys = torch.cat([tensors[r][0][int(b[r, g]):int(b[r, g + 1])] for r in range(world)])
        cnt = torch.cat([tensors[r][2][int(b[r, g]):int(b[r, g + 1])] for r in range(world)])
        var = torch.cat([tensors[r][1][int(b[r, g]) * int(cols[r]):int(b[r, g + 1]) * int(cols[r])] for r in range(world)])
        slab = backend.merge_partial(keys, var, cnt, recv, cols, col_base, 4, k, True, names)
        parts.append(slab.export())
    o = O.OracleArray.build(list(os_), names, k=k, rc=True, min_count=3, min_qual=20, qual_filter="strict")
    ok, ov, oc = o.export()
    assert [int(x) for p in parts for x in p[0]] == ok
    assert np.array_equal(np.concatenate([p[1] for p in parts]), ov)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), oc)


@pytest.mark.parametrize("k,rc", [(9, True), (15, False), (33, True), (63, True)])
def test_kmer_count_histogram_matches_dictionary(gold_in, k, rc):
    """the counting half of `ska cov` (coverage.rs:70-135, 156-163): distinct split k-mers by multiplicity, qualities ignored"""
    import collections
    stem = "test_long" if k > 31 else "test"
    s = S.Sample.from_files(gold_in(f"{stem}_1_fwd.fastq.gz"), gold_in(f"{stem}_1_rev.fastq.gz"))
    # repeat the reads a few times so that multiplicities above 1 dominate, plus a read with N
    seq = np.concatenate([s.seq] * 3 + [np.frombuffer(b"ACGTNACGT" * 20, dtype=np.uint8)])
    off = np.concatenate([s.rec_off, s.rec_off[1:] + s.rec_off[-1], s.rec_off[1:] + 2 * s.rec_off[-1], [3 * s.rec_off[-1] + 180]]).astype(np.uint64)
    big = S.Sample(seq, off, np.full(len(seq), ord("!"), dtype=np.uint8))
    counts = collections.Counter()
    for i in range(big.n_rec):
        w = O.windows(bytes(big.seq[int(big.rec_off[i]):int(big.rec_off[i + 1])]), k, rc=rc)
        counts.update(w["key"])
    exp = np.zeros(1000, dtype=np.uint64)
    for c in counts.values():
        if c <= 1000:
            exp[c - 1] += 1
    got = S.api.kmer_count_hist(big, k, rc)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("min_count,qual_filter", [(3, "strict"), (1, "middle")])
def test_build_mixes_fasta_and_fastq_samples(min_count, qual_filter):
    """is_reads is a per-sample property (ska_dict.rs:356-366): an assembly next to read sets in one build keeps every
    valid window (no count filter, no quality test), the read sets are filtered."""
    from ska_rust_b200 import synth
    genomes = synth.star_genomes(3, 20_000, seed=0x5CA2000C, snp_rate=0.002)
    gs0, os0 = make_reads_sample(genomes[0], 3000, 0.01, 11)
    gs2, os2 = make_reads_sample(genomes[2], 3000, 0.01, 12)
    asm = genomes[1]
    off = np.array([0, 7000, 7000 + 31, len(asm)], dtype=np.uint64)          # three contigs, one of exactly k bases
    g_asm = S.Sample(asm, off)
    o_asm = O.SampleData([(asm[int(off[i]):int(off[i + 1])].tobytes(), None) for i in range(3)], is_reads=False)
    names = ["reads0", "assembly", "reads2"]
    q = S.QualOpts(min_count=min_count, min_qual=20, qual_filter=qual_filter)
    g = S.build_from_samples([gs0, g_asm, gs2], names, 31, True, q)
    o = O.OracleArray.build([os0, o_asm, os2], names, k=31, rc=True, min_count=min_count, min_qual=20, qual_filter=qual_filter)
    same(g, o)
    # the assembly's column is what a FASTA-only build gives
    alone = S.build_from_samples([g_asm], ["assembly"], 31, True)
    gk, gv, _ = g.export()
    ak, av, _ = alone.export()
    col = {int(k_): int(v) for k_, v in zip(gk, gv[:, 1]) if v != ord("-")}
    assert col == {int(k_): int(v) for k_, v in zip(ak, av[:, 0])}
