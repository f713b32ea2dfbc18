"""GPU parity tests: the CUDA path (through the C ABI of libskagpu.so) against the CPU oracle and
against the reference's own golden outputs.  Bit-exact everywhere (integer/byte work); the only
floating-point output -- the `--allow-ambiguous` distance -- is compared to 1e-9 and, as printed
text (`{:.2}`), exactly.  Run on the B200 box:  python -m pytest tests -m gpu
"""
import io
import os

import numpy as np
import pytest

import ska_rust_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as O
from oracle import skf_py

pytestmark = pytest.mark.gpu

SENT = np.uint64(0xFFFFFFFFFFFFFFFF)


# ------------------------------------------------------------------ helpers
def gpu_build_files(files, k=31, rc=True, names=None, **kw):
    inputs = []
    for i, f in enumerate(files):
        f1, f2 = (f, None) if isinstance(f, str) else f
        inputs.append(((names[i] if names else O.sample_name(f1)), f1, f2))
    prop = kw.pop("proportion_reads", None)
    return S.build_and_merge(inputs, k, rc, S.QualOpts(**kw), proportion_reads=prop)


def ora_from_gpu_samples(samples, names, k, rc):
    osamples = []
    for s in samples:
        recs = [(bytes(s.seq[int(s.rec_off[i]):int(s.rec_off[i + 1])]), None) for i in range(s.n_rec)]
        osamples.append(O.SampleData(recs))
    return O.OracleArray.build(osamples, names, k=k, rc=rc)


def assert_same_array(g, o):
    gd, od = g.dims(), o.dims()
    assert gd == od, (gd, od)
    gk, gv, gc = g.export()
    ok, ov, oc = o.export()
    assert [int(x) for x in gk] == ok
    assert np.array_equal(gv, ov)
    assert np.array_equal(gc, oc)


def random_records(rng, n_rec, max_len, p_n=0.02, lower=True):
    recs = []
    alphabet = np.frombuffer(b"ACGTacgtNnRY", dtype=np.uint8)
    probs = np.array([0.24, 0.24, 0.24, 0.24, 0.002, 0.002, 0.002, 0.002, p_n / 2, p_n / 2, 0.001, 0.001])
    probs = probs / probs.sum()
    for _ in range(n_rec):
        ln = int(rng.integers(0, max_len))
        recs.append(rng.choice(alphabet, size=ln, p=probs).tobytes())
    return recs


def expected_records(samples_records, k, rc):
    """per-position K1 records from the oracle's rolling iterator"""
    out = []
    for recs in samples_records:
        for r in recs:
            e = np.full(len(r), SENT, dtype=np.uint64)
            w = O.windows(r, k, rc=rc)
            h = (k - 1) // 2
            for i in range(w["n"]):
                start = int(w["mid_pos"][i]) - h
                e[start] = np.uint64((w["key"][i] << 4) | (1 << int(w["mid"][i])))
            out.append(e)
    return np.concatenate(out) if out else np.zeros(0, dtype=np.uint64)


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("k,rc", [(5, True), (7, False), (9, True), (15, True), (15, False), (21, True), (31, True), (31, False)])
def test_k1_windows_match_rolling_iterator(k, rc):
    rng = np.random.default_rng(1234 + k)
    samples_records = [random_records(rng, int(rng.integers(1, 6)), 3000) for _ in range(5)]
    # edge cases: record of exactly k, k+1, k-1 bases, empty record, run of exactly k after an N at the end
    core = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=k + 1))
    samples_records.append([core[:k], core, core[:k - 1], b"", core + b"N" + core[:k], core[:k] + b"N" + core,
                            b"N" * 40, core * 3])
    samples = [S.Sample.from_records(r) for r in samples_records]
    batch = S.SeqBatch(samples)
    rec, _ms = batch.k1_extract(k, rc=rc)
    exp = expected_records(samples_records, k, rc)
    assert rec.shape == exp.shape
    bad = np.nonzero(rec != exp)[0]
    assert bad.size == 0, f"{bad.size} mismatching positions, first at {bad[:5]}"


def test_k1_tile_boundaries_long_record():
    """windows that straddle the 4096-position tiles of the kernel, N runs at tile edges"""
    rng = np.random.default_rng(7)
    seq = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=20000).tobytes())
    for p in (4090, 4095, 4096, 4097, 8191, 8192, 12287 + 31, 16384 - 31):
        seq[p] = ord("N")
    recs = [[bytes(seq)], [bytes(seq[:4096])], [bytes(seq[:4096 + 30])], [bytes(seq[100:4196])]]
    batch = S.SeqBatch([S.Sample.from_records(r) for r in recs])
    for k in (31, 9):
        rec, _ = batch.k1_extract(k, rc=True)
        assert np.array_equal(rec, expected_records(recs, k, True))


def test_k1_strict_quality_mask():
    """split_kmer.rs:66-71,98-101: with QualFilter::Strict a base is valid only if qual-33 > min_qual"""
    rng = np.random.default_rng(11)
    n = 5000
    seq = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n, p=[.245, .245, .245, .245, .02])
    qual = rng.choice(np.array([33 + 2, 33 + 20, 33 + 21, 33 + 40], dtype=np.uint8), size=n, p=[.02, .02, .06, .9])
    off = np.array([0, 1700, 1700, 3300, n], dtype=np.uint64)
    s = S.Sample(seq, off, qual)
    batch = S.SeqBatch([s])
    k = 15
    rec, _ = batch.k1_extract(k, rc=True, min_qual=20, qual_filter="strict")
    exp = []
    for i in range(4):
        a, b = int(off[i]), int(off[i + 1])
        e = np.full(b - a, SENT, dtype=np.uint64)
        w = O.windows(seq[a:b].tobytes(), k, rc=True, qual=qual[a:b].tobytes(), min_qual=20, qual_filter="strict", is_reads=True)
        for j in range(w["n"]):
            e[int(w["mid_pos"][j]) - 7] = np.uint64((w["key"][j] << 4) | (1 << int(w["mid"][j])))
        exp.append(e)
    assert np.array_equal(rec, np.concatenate(exp))


# ------------------------------------------------------------------ build: reference fixtures
SKF_RECIPES = {
    "multidist.skf": (9, ["N_test_1.fa", "N_test_2.fa", "ambig_test_1.fa", "ambig_test_2.fa", "test_1.fa", "test_2.fa"]),
    "merge_k9.skf": (9, ["test_1.fa", "test_2.fa"]),
    "merge.skf": (17, ["test_1.fa", "test_2.fa"]),
}


@pytest.mark.parametrize("fixture", sorted(SKF_RECIPES))
def test_build_reproduces_reference_skf(gold_in, fixture):
    k, files = SKF_RECIPES[fixture]
    ref = skf_py.load_skf(gold_in(fixture))
    var = np.frombuffer(ref["variants"], dtype=np.uint8).reshape(ref["n_rows"], ref["n_samples"])
    ref_map = {key: bytes(var[i]) for i, key in enumerate(ref["split_kmers"])}
    ref_cnt = {key: ref["variant_count"][i] for i, key in enumerate(ref["split_kmers"])}
    g = gpu_build_files([gold_in(f) for f in files], k=k)
    assert g.names() == ref["names"]
    keys, gv, gc = g.export()
    assert list(keys) == sorted(keys)
    assert {int(key): bytes(gv[i]) for i, key in enumerate(keys)} == ref_map
    assert {int(key): int(gc[i]) for i, key in enumerate(keys)} == ref_cnt


def aln(arr, **kw):
    buf = io.BytesIO()
    S.api.align(arr, buf, **kw)
    return buf.getvalue()


def dist(arr, **kw):
    buf = io.BytesIO()
    S.api.distance(arr, buf, **kw)
    return buf.getvalue()


def test_align_goldens(gold_in, gold_ok):
    """tests/fasta_input.rs: align_n, repeats, palindromes -- byte-exact stdout"""
    b = lambda names, **kw: gpu_build_files([gold_in(n) for n in names], **kw)
    assert aln(b(["N_test_1.fa", "N_test_2.fa"])) == gold_ok("align_N.stdout")
    assert aln(b(["dup_test_1.fa", "dup_test_2.fa"], k=9, rc=False)) == gold_ok("dup_ss.stdout")
    assert aln(b(["dup_test_1.fa", "dup_test_2.fa"], k=9)) == gold_ok("dup_rc.stdout")
    assert aln(b(["palindrome_1.fa", "palindrome_2.fa"], k=15), filter_type="no-filter") == gold_ok("palindrome.stdout")
    assert aln(b(["palindrome_1.fa", "palindrome_2.fa"], k=15, rc=False)) == gold_ok("palindrome_norc.stdout")
    assert aln(b(["palindrome_reps_1.fa", "palindrome_reps_2.fa"], k=15), filter_type="no-filter") == gold_ok("palindrome_reps.stdout")


def test_align_column_sets(gold_in):
    """tests/align.rs:86-115,168-183; tests/fasta_input.rs:60-154"""
    b = lambda names, **kw: gpu_build_files([gold_in(n) for n in names], **kw)
    assert O.var_hash(aln(b(["test_1.fa", "test_2.fa"], k=15))) == {("A", "T"), ("C", "T")}
    assert O.var_hash(aln(b(["test_1.fa", "test_2.fa"]))) == {("A", "T"), ("C", "T")}
    assert O.var_hash(aln(b(["test_1.fa", "test_2_rc.fa"], k=15))) == {("A", "T"), ("C", "T")}
    assert O.var_hash(aln(b(["test_1.fa", "test_2_rc.fa"], k=15, rc=False))) == set()


def test_nk_golden(gold_in, gold_ok):
    g = gpu_build_files([gold_in("test_1.fa"), gold_in("test_2.fa")], k=31)
    assert g.display().split("\n", 1)[1] == gold_ok("merge_nk.stdout").decode().split("\n", 1)[1]
    g = gpu_build_files([gold_in("dup_test_1.fa"), gold_in("dup_test_2.fa")], k=9, rc=False)
    g.filter(2, False, "no-const", False, False, True)
    assert g.display(full_info=True).split("\n", 1)[1] == gold_ok("dup_ss_nk.stdout").decode().split("\n", 1)[1]


def test_proportion_reads(gold_in):
    g = gpu_build_files([gold_in("proportion_reads.fa")], k=17, rc=False, proportion_reads=0.5)
    o = O.OracleArray.build_from_files([gold_in("proportion_reads.fa")], k=17, rc=False, proportion_reads=0.5)
    assert_same_array(g, o)
    assert g.dims()["n_rows"] == 2


def test_parallel_align_45_samples(gold_in):
    """tests/align.rs:349-397 -- 45 gz copies; result independent of scheduling"""
    d = gold_in("par_test")
    files = sorted(os.path.join(d, f) for f in os.listdir(d))
    g = gpu_build_files(files, k=15)
    o = O.OracleArray.build_from_files(files, k=15, threads=4)
    assert_same_array(g, o)
    assert aln(g) == o.align()


def test_errors(gold_in):
    f = [gold_in("test_1.fa"), gold_in("test_2.fa")]
    for bad_k in (65, 30, 3):
        with pytest.raises(S.SkaError, match="Invalid k-mer length"):
            gpu_build_files(f, k=bad_k)
    with pytest.raises(S.SkaError, match="no valid sequence"):          # ska_dict.rs:374-376
        S.build_from_samples([S.Sample.from_records([b"ACGTNACGTNACGT"]), S.Sample.from_records([b"ACGTACGTACGTAAA"])],
                             ["a", "b"], 9, True)
    with pytest.raises(S.SkaError):
        gpu_build_files(["/nonexistent/x.fa"], k=9)


# ------------------------------------------------------------------ build: synthetic, vs oracle
@pytest.mark.parametrize("n,length,k,rc", [(4, 100_000, 31, True), (6, 30_000, 15, False), (3, 50_000, 21, True),
                                           (9, 20_011, 9, True), (130, 3_001, 31, True)])
def test_build_synthetic_matches_oracle(n, length, k, rc):
    """BASELINE config 1 (4 x 100 kbp, k=31) and config-5-shaped (k=15 single strand) inputs"""
    from ska_rust_b200 import synth
    genomes = synth.star_genomes(n, length, seed=0x5CA20001 + k, snp_rate=0.002)
    samples = [S.Sample.from_records([g.tobytes()]) for g in genomes]
    names = [f"s{i}" for i in range(n)]
    g = S.build_from_samples(samples, names, k, rc)
    o = ora_from_gpu_samples(samples, names, k, rc)
    assert_same_array(g, o)


@pytest.mark.parametrize("switch", [{"SKA_SCATTER_V1": "1"}, {"SKA_HIST2_BYTES": "0"}, {"SKA_PREFIX_BITS": "13"}, {"SKA_EMIT_CAP_FACTOR": "2"}])
@pytest.mark.parametrize("n,length,k,rc", [(5, 60_000, 31, True), (130, 3_001, 21, False), (300, 1_500, 31, True)])
def test_build_partition_forms(n, length, k, rc, switch, monkeypatch):
    """the partition's alternative forms give the same array: the first scatter kernel (staged digits), the second
    histogram from the records instead of the digit bytes, a digit split other than 8 + 8 (run-time digit widths: the
    warp-0 scan and the record histogram), a denser emit table.  Many short samples put several sample boundaries into one
    scatter tile (the per-sample rounds of the straddling launch); 300 samples take the sample search of thread 0."""
    from ska_rust_b200 import synth
    for k_, v in switch.items():
        monkeypatch.setenv(k_, v)
    genomes = synth.star_genomes(n, length, seed=0x5CA20700 + n, snp_rate=0.003)
    samples = [S.Sample.from_records([g.tobytes()]) for g in genomes]
    names = [f"s{i}" for i in range(n)]
    g = S.build_from_samples(samples, names, k, rc)
    assert_same_array(g, ora_from_gpu_samples(samples, names, k, rc))


def test_merge_kernels_cross_their_record_chunks():
    """bucket_count / bucket_emit walk a bucket's records in chunks of 2^31 with 32-bit positions inside a chunk; the
    library variant built by __graft_entry__.build() with chunks of 4096 records crosses that boundary in every
    bucket (another process: a library is bound once per process)."""
    import subprocess
    import sys
    lib = os.path.join(ROOT, "ska.rust_b200", "variants", "libskagpu_chunk12.so")
    if not os.path.exists(lib):
        pytest.skip("variants/libskagpu_chunk12.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    code = (
        "import numpy as np, ska_rust_b200 as S\n"
        "from oracle import oracle as O\n"
        "from ska_rust_b200 import synth\n"
        "for n, length, k in ((40, 30000, 31), (8, 150000, 15)):\n"
        "    gs = synth.star_genomes(n, length, seed=0x5CA20900 + n, snp_rate=0.002)\n"
        "    names = [f's{i}' for i in range(n)]\n"
        "    g = S.build_from_samples([S.Sample.from_records([x.tobytes()]) for x in gs], names, k, True)\n"
        "    o = O.OracleArray.build([O.SampleData([(x.tobytes(), None)]) for x in gs], names, k=k, rc=True)\n"
        "    gk, gv, gc = g.export(); ok, ov, oc = o.export()\n"
        "    assert [int(x) for x in gk] == ok and np.array_equal(gv, ov) and np.array_equal(gc, oc)\n"
        "print('chunks ok')\n")
    env = dict(os.environ, SKAGPU_LIB=lib)                            # buckets hold ~10 k records: two to three chunks each
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "chunks ok" in p.stdout, p.stderr[-2000:]


MERGE_PATHS = {"fused": ("1", "0"), "two-pass": ("0", "0"), "two-pass-global": ("0", "1")}   # SKA_MERGE_FUSED, SKA_EMIT_GLOBAL


@pytest.mark.parametrize("n,k,rc", [(8, 31, True), (24, 15, True), (16, 9, True), (40, 11, False), (128, 21, True)])
@pytest.mark.parametrize("path", sorted(MERGE_PATHS))
def test_build_emit_paths(n, k, rc, path, monkeypatch):
    """the three merge paths of merge.cu on the same inputs: the fused single pass (table + base sets in shared memory,
    row offsets by look-back), count + shared-memory emit, count + emit with the base sets OR-ed into the output in
    global memory (wide rows).  Sample counts are multiples of 8 (8-byte row stores; set words per row a power of
    two or not)."""
    from ska_rust_b200 import synth
    monkeypatch.setenv("SKA_MERGE_FUSED", MERGE_PATHS[path][0])
    monkeypatch.setenv("SKA_EMIT_GLOBAL", MERGE_PATHS[path][1])
    genomes = synth.star_genomes(n, 6_000, seed=0x5CA20100 + n, snp_rate=0.01)
    samples = []
    for i, g in enumerate(genomes):
        recs = [g.tobytes(), b"ACGT" * 20, b"AATT" * 10 + b"GGCC" * 10]       # palindromic keys and IUPAC merges in every sample
        if i % 3 == 0:
            recs.append(g[100:400].tobytes())                                  # repeats inside a sample
        samples.append(S.Sample.from_records(recs))
    names = [f"s{i}" for i in range(n)]
    g = S.build_from_samples(samples, names, k, rc)
    assert_same_array(g, ora_from_gpu_samples(samples, names, k, rc))


@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("prefix_bits", [None, 0, 6])
def test_build_skewed_keys_long_probe_runs(prefix_bits, fused, monkeypatch):
    """low-complexity sequence: many distinct keys share long prefixes, so the order-preserving table of the count pass
    has long runs of occupied slots (and, with few prefix bits, items that overflow and are split); order must survive.
    With the fused merge forced on, buckets that overflow its table make it give up and the two-pass path takes over."""
    monkeypatch.setenv("SKA_MERGE_FUSED", fused)
    if prefix_bits is not None:
        monkeypatch.setenv("SKA_PREFIX_BITS", str(prefix_bits))
    rng = np.random.default_rng(77)
    samples_records = []
    for s in range(5):
        recs = []
        for _ in range(40):
            head = b"A" * int(rng.integers(12, 20))                            # keys starting AAAA...: one corner of the key space
            tail = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=600).tobytes()
            recs.append(head + tail + b"T" * 25 + tail[:50])
        recs.append(rng.choice(np.frombuffer(b"AC", dtype=np.uint8), size=20_000).tobytes())   # two-letter alphabet
        samples_records.append(recs)
    samples = [S.Sample.from_records(r) for r in samples_records]
    names = [f"s{i}" for i in range(len(samples))]
    for k, rc in ((31, True), (15, False), (9, True)):
        g = S.build_from_samples(samples, names, k, rc)
        assert_same_array(g, ora_from_gpu_samples(samples, names, k, rc))


@pytest.mark.parametrize("fused", ["0", "1"])
def test_build_unrelated_samples_forces_bucket_splitting(fused, monkeypatch):
    """samples that share nothing: every record is a distinct row, so buckets exceed the shared-memory table and are
    split by further key bits (merge.cu, rare path); the fused merge must decline such input"""
    monkeypatch.setenv("SKA_MERGE_FUSED", fused)
    rng = np.random.default_rng(4242)
    samples_records = [[rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=400_000).tobytes()] for _ in range(8)]
    samples = [S.Sample.from_records(r) for r in samples_records]
    names = [f"u{i}" for i in range(8)]
    g = S.build_from_samples(samples, names, 31, True)
    o = ora_from_gpu_samples(samples, names, 31, True)
    assert_same_array(g, o)
    assert g.dims()["n_rows"] > 3_100_000


def test_build_multi_record_with_ns_and_repeats():
    rng = np.random.default_rng(99)
    samples_records = []
    for _ in range(7):
        recs = random_records(rng, int(rng.integers(1, 30)), 2500, p_n=0.01)
        recs.append(recs[0] * 2)                       # repeats -> IUPAC merges within a sample
        recs.append(b"ACGT" * 50)                      # low complexity: same key with several middles, palindromes
        samples_records.append(recs)
    samples = [S.Sample.from_records(r) for r in samples_records]
    names = [f"s{i}" for i in range(len(samples))]
    for k, rc in ((9, True), (7, True), (11, False), (31, True)):
        g = S.build_from_samples(samples, names, k, rc)
        assert_same_array(g, ora_from_gpu_samples(samples, names, k, rc))


@pytest.mark.parametrize("k,rc", [(31, True), (15, False), (9, True)])
def test_build_k1_histogram_fusion(k, rc, monkeypatch):
    """SKA_K1_HIST=1: K1 also counts the first partition digit (the partition then skips its own histogram pass);
    same array either way"""
    from ska_rust_b200 import synth
    n = 6
    genomes = synth.star_genomes(n, 150_000, seed=0x5CA20044 + k, snp_rate=0.003)
    samples = []
    for i, g in enumerate(genomes):
        g = g.copy()
        g[700 * (i + 1):700 * (i + 1) + 40] = ord("N")
        samples.append(S.Sample.from_records([g[:70_001].tobytes(), g[70_001:].tobytes(), b"ACGT" * 9]))
    names = [f"s{i}" for i in range(n)]
    monkeypatch.setenv("SKA_K1_HIST", "1")
    a = S.build_from_samples(samples, names, k, rc)
    monkeypatch.setenv("SKA_K1_HIST", "0")
    b = S.build_from_samples(samples, names, k, rc)
    ka, va, ca = a.export()
    kb, vb, cb = b.export()
    assert np.array_equal(ka, kb) and np.array_equal(va, vb) and np.array_equal(ca, cb)
    assert_same_array(a, ora_from_gpu_samples(samples, names, k, rc))


@pytest.mark.parametrize("small_bytes", ["1500", "4000", "20000"])
def test_build_emit_size_classes(small_bytes, monkeypatch):
    """bucket_emit_smem_kernel runs in two launches when shared memory sized for the largest bucket would cost
    occupancy: buckets of up to u_small rows first, the rest after.  SKA_EMIT_SMALL_BYTES moves the class boundary down
    so that both classes are populated at test sizes; the array must equal the one-launch array and the oracle's."""
    from ska_rust_b200 import synth
    n, k = 16, 31
    genomes = synth.star_genomes(n, 120_000, seed=0x5CA20066, snp_rate=0.004)
    samples = [S.Sample.from_records([g.tobytes(), b"ACGT" * 16]) for g in genomes]
    names = [f"s{i}" for i in range(n)]
    monkeypatch.setenv("SKA_PREFIX_BITS", "10")                    # ~1900 records, ~350 rows per bucket (up to ~700 at low prefixes)
    monkeypatch.setenv("SKA_EMIT_SMALL_BYTES", small_bytes)
    a = S.build_from_samples(samples, names, k, True)
    monkeypatch.setenv("SKA_EMIT_ONE_CLASS", "1")
    b = S.build_from_samples(samples, names, k, True)
    ka, va, ca = a.export()
    kb, vb, cb = b.export()
    assert np.array_equal(ka, kb) and np.array_equal(va, vb) and np.array_equal(ca, cb)
    assert_same_array(a, ora_from_gpu_samples(samples, names, k, True))


def test_build_overlapped_upload_matches_plain(monkeypatch):
    """skagpu_build on FASTA input runs K1 chunk by chunk behind the host->device copies (8 MB chunks; tiles that
    straddle a chunk boundary wait for the next chunk); SKA_H2D_OVERLAP=0 uploads first and extracts afterwards.
    44 samples of uneven size (26 MB in all) so that several chunks and many sample boundaries are involved."""
    from ska_rust_b200 import synth
    rng = np.random.default_rng(12)
    lengths = [int(x) for x in rng.integers(1_000, 1_200_000, size=44)]
    anc = synth.star_genomes(1, 1_200_000, seed=0x5CA20055)[0]
    samples = []
    for i, ln in enumerate(lengths):
        g = anc[:ln].copy()
        pos = rng.integers(0, ln, size=max(1, ln // 800))
        g[pos] = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=len(pos))
        cut = ln // 3
        samples.append(S.Sample.from_records([g[:cut].tobytes(), g[cut:].tobytes()]))
    names = [f"s{i}" for i in range(len(samples))]
    monkeypatch.setenv("SKA_H2D_OVERLAP", "1")
    a = S.build_from_samples(samples, names, 31, True)
    assert any(st == "h2d+k1" for st, _, _ in S.api.last_timings())
    monkeypatch.setenv("SKA_H2D_OVERLAP", "0")
    b = S.build_from_samples(samples, names, 31, True)
    assert any(st == "h2d" for st, _, _ in S.api.last_timings())
    ka, va, ca = a.export()
    kb, vb, cb = b.export()
    assert np.array_equal(ka, kb) and np.array_equal(va, vb) and np.array_equal(ca, cb)
    sub = S.build_from_samples(samples[:5], names[:5], 31, True)
    assert_same_array(sub, ora_from_gpu_samples(samples[:5], names[:5], 31, True))


def test_build_fused_many_buckets_matches_two_pass(monkeypatch):
    """24 x 700 kbp: ~1000 buckets, so the fused merge's look-back chain spans many waves of CTAs; its array must equal
    the two-pass array bit for bit (and the oracle's on a 3-sample subset)"""
    from ska_rust_b200 import synth
    n, length, k = 24, 700_000, 31
    genomes = synth.star_genomes(n, length, seed=0x5CA20033, snp_rate=0.002)
    samples = [S.Sample.from_records([g.tobytes()]) for g in genomes]
    names = [f"s{i}" for i in range(n)]
    monkeypatch.setenv("SKA_PREFIX_BITS", "12")
    monkeypatch.setenv("SKA_MERGE_FUSED", "1")
    a = S.build_from_samples(samples, names, k, True)
    stages = [s_ for s_, _, _ in S.api.last_timings()]
    monkeypatch.setenv("SKA_MERGE_FUSED", "0")
    b = S.build_from_samples(samples, names, k, True)
    ka, va, ca = a.export()
    kb, vb, cb = b.export()
    assert np.array_equal(ka, kb) and np.array_equal(va, vb) and np.array_equal(ca, cb)
    assert "merge" in stages and "emit" not in stages, f"the fused merge should not have fallen back: {stages}"
    monkeypatch.setenv("SKA_MERGE_FUSED", "1")
    sub = S.build_from_samples(samples[:8], names[:8], k, True)
    assert_same_array(sub, ora_from_gpu_samples(samples[:8], names[:8], k, True))


def test_build_properties_medium():
    """size-independent properties on 16 x 1 Mbp: sorted unique keys, counts == non-gap cells,
    idempotence, column permutation equivariance"""
    from ska_rust_b200 import synth
    n, length, k = 16, 1_000_000, 31
    genomes = synth.star_genomes(n, length, seed=0x5CA20002)
    samples = [S.Sample.from_records([g.tobytes()]) for g in genomes]
    names = [f"s{i}" for i in range(n)]
    g = S.build_from_samples(samples, names, k, True)
    keys, var, cnt = g.export()
    assert np.all(keys[1:] > keys[:-1])
    assert np.array_equal((var != ord("-")).sum(axis=1), cnt)
    assert int(keys.max()) < (1 << 60)
    # every sample sees ~length distinct split k-mers
    per_sample = (var != ord("-")).sum(axis=0)
    assert np.all(per_sample > 0.99 * (length - k + 1)) and np.all(per_sample <= length - k + 1)
    g2 = S.build_from_samples(samples, names, k, True)
    k2, v2, c2 = g2.export()
    assert np.array_equal(keys, k2) and np.array_equal(var, v2) and np.array_equal(cnt, c2)
    perm = np.random.default_rng(5).permutation(n)
    g3 = S.build_from_samples([samples[i] for i in perm], [names[i] for i in perm], k, True)
    k3, v3, _ = g3.export()
    assert np.array_equal(keys, k3) and np.array_equal(var[:, perm], v3)
    # one sample against the oracle
    o = ora_from_gpu_samples(samples[:2], names[:2], k, True)
    g4 = S.build_from_samples(samples[:2], names[:2], k, True)
    assert_same_array(g4, o)


# ------------------------------------------------------------------ filter
def load_pair(gold_in, name):
    d = skf_py.load_skf(gold_in(name))
    var = np.frombuffer(d["variants"], dtype=np.uint8).reshape(d["n_rows"], d["n_samples"])
    key_bits = 64 if all(v < (1 << 64) for v in d["split_kmers"]) else 128
    g = S.MergeSkaArray.from_arrays(d["k"], d["rc"], key_bits, d["split_kmers"], var, d["variant_count"], d["names"])
    o = O.OracleArray.from_arrays(d["k"], d["rc"], key_bits, d["split_kmers"], var, d["variant_count"], d["names"])
    return g, o


FILTER_CASES = [
    (dict(filter_type="no-filter", ignore_const_gaps=True), "len", 38),
    (dict(filter_type="no-ambig", filter_ambig_as_missing=True), "len", 37),
    (dict(filter_type="no-const"), "set", {("T", "A"), ("C", "T"), ("S", "G")}),
    (dict(filter_type="no-ambig-or-const"), "set", {("T", "A"), ("C", "T")}),
    (dict(filter_type="no-const", mask_ambig=True), "set", {("T", "A"), ("C", "T"), ("N", "G")}),
    (dict(filter_type="no-const", min_freq=0), "len", 33),
    (dict(filter_type="no-const", min_freq=0, ignore_const_gaps=True), "len", 3),
    (dict(filter_type="no-ambig-or-const", min_freq=0), "len", 32),
    (dict(filter_type="no-ambig-or-const", min_freq=0, ignore_const_gaps=True), "len", 2),
]


@pytest.mark.parametrize("kw,kind,want", FILTER_CASES)
def test_filter_matrix_merge_k9(gold_in, kw, kind, want):
    """tests/align.rs:186-347"""
    g, _ = load_pair(gold_in, "merge_k9.skf")
    out = aln(g, **kw)
    if kind == "len":
        assert O.aln_length(out) == [want, want]
    else:
        assert O.var_hash(out) == want


def random_array(rng, n_rows, n_samples):
    alphabet = np.frombuffer(b"ACGT-ACGT-ACGT-RYSWKMBDHVNacgtuU", dtype=np.uint8)
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n_rows, 1))
    var = np.repeat(base, n_samples, axis=1)
    noise = rng.random((n_rows, n_samples)) < 0.08
    var[noise] = rng.choice(alphabet, size=int(noise.sum()))
    gaps = rng.random((n_rows, n_samples)) < 0.05
    var[gaps] = ord("-")
    var[rng.random(n_rows) < 0.02] = ord("-")                       # all-gap rows
    cnt = (var != ord("-")).sum(axis=1).astype(np.uint64)
    keys = np.sort(rng.choice(1 << 40, size=n_rows, replace=False)).astype(np.uint64)
    return keys, var, cnt


@pytest.mark.parametrize("n_rows,n_samples", [(1000, 7), (5000, 128), (3001, 33), (300, 200), (700, 256), (257, 1024)])
def test_filter_random_arrays_match_oracle(n_rows, n_samples):
    rng = np.random.default_rng(n_rows + n_samples)
    keys, var, cnt = random_array(rng, n_rows, n_samples)
    names = [f"s{i}" for i in range(n_samples)]
    for ft in ("no-filter", "no-const", "no-ambig", "no-ambig-or-const"):
        for ambig_missing in (False, True):
            for gaps in (False, True):
                for mask in (False, True):
                    for upd in (False, True):
                        for min_count in (0, n_samples // 2, n_samples):
                            g = S.MergeSkaArray.from_arrays(31, True, 64, keys, var, cnt, names)
                            o = O.OracleArray.from_arrays(31, True, 64, [int(x) for x in keys], var, cnt, names)
                            rg = g.filter(min_count, ambig_missing, ft, mask, gaps, upd)
                            ro = o.filter(min_count, ambig_missing, ft, mask, gaps, upd)
                            assert rg == ro, (ft, ambig_missing, gaps, mask, upd, min_count)
                            assert_same_array(g, o)


@pytest.mark.parametrize("cols", [None, "128"])
def test_transpose_matches_write_fasta(cols, monkeypatch):
    """both tile widths of transpose_words_kernel (64 columns by default; whole 128-byte row lines on request)"""
    if cols:
        monkeypatch.setenv("SKA_TRANSPOSE_COLS", cols)
    rng = np.random.default_rng(3)
    for n_rows, n_samples in ((1, 1), (63, 5), (64, 64), (1000, 129), (4097, 128), (1, 4), (3, 8), (130, 68), (1027, 200),
                              (12_345, 128), (517, 256), (2, 1024)):
        keys, var, cnt = random_array(rng, n_rows, n_samples)
        g = S.MergeSkaArray.from_arrays(31, True, 64, keys, var, cnt, [f"s{i}" for i in range(n_samples)])
        assert np.array_equal(g.transposed(), var.T)


# ------------------------------------------------------------------ distance
DIST_CASES = [
    ("merge.skf", {}, "merge.dist.stdout"),
    ("merge_k41.skf", {}, "merge_k41.dist.stdout"),
    ("merge_k9.skf", {"filt_ambig": False}, "merge_k9.dist.stdout"),
    ("merge_k9.skf", {}, "merge_k9_no_ambig.dist.stdout"),
    ("merge_k9.skf", {"min_freq": 1.0}, "merge_k9_min_freq.dist.stdout"),
    ("multidist.skf", {}, "multidist.stdout"),
    ("multidist.skf", {"min_freq": 0.9}, "multidist.minfreq.stdout"),
    ("multidist.skf", {"filt_ambig": False}, "multidist.ambig.stdout"),
]


@pytest.mark.parametrize("skf,kw,golden", DIST_CASES)
def test_distance_goldens(gold_in, gold_ok, skf, kw, golden):
    """tests/distance.rs -- byte-exact stdout, including the u128 fixture merge_k41.skf"""
    g, _ = load_pair(gold_in, skf)
    assert dist(g, **kw) == gold_ok(golden)


def test_distance_after_gpu_build_multidist(gold_in, gold_ok):
    k, files = SKF_RECIPES["multidist.skf"]
    g = gpu_build_files([gold_in(f) for f in files], k=k)
    assert dist(g) == gold_ok("multidist.stdout")


@pytest.mark.parametrize("n_rows,n_samples", [(5000, 5), (20000, 40), (1000, 129), (33, 17), (0, 4)])
@pytest.mark.parametrize("filt_ambig", [True, False])
def test_distance_random_arrays_match_oracle(n_rows, n_samples, filt_ambig):
    rng = np.random.default_rng(17 + n_rows)
    if n_rows:
        keys, var, cnt = random_array(rng, n_rows, n_samples)
        if filt_ambig and n_rows == 20000:      # plain upper-case alphabet -> exercises the bit-plane kernel
            var = np.where(np.isin(var, np.frombuffer(b"acgtuU", dtype=np.uint8)), np.uint8(ord("A")), var)
    else:
        keys, var, cnt = np.zeros(0, np.uint64), np.zeros((0, n_samples), np.uint8), np.zeros(0, np.uint64)
    names = [f"s{i}" for i in range(n_samples)]
    g = S.MergeSkaArray.from_arrays(31, True, 64, keys, var, cnt, names)
    o = O.OracleArray.from_arrays(31, True, 64, [int(x) for x in keys], var, cnt, names)
    dg = g.distance(12.0, filt_ambig)
    do = o.distance_raw(12.0, filt_ambig, threads=4)
    assert len(dg) == len(do)
    for a, b in zip(dg, do):
        assert a[2] == b[2] and a[3] == b[3]
        if filt_ambig:
            assert a[0] == b[0] and a[1] == b[1]
        else:
            assert abs(a[0] - b[0]) < 1e-9 and a[1] == b[1]
            assert f"{a[0]:.2f}" == f"{b[0]:.2f}"


def test_distance_bitplane_path_large():
    """upper-case-only array big enough to split rows over many CTAs"""
    rng = np.random.default_rng(21)
    n_rows, n_samples = 200_000, 24
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n_rows, 1))
    var = np.repeat(base, n_samples, axis=1)
    noise = rng.random((n_rows, n_samples)) < 0.05
    var[noise] = rng.choice(np.frombuffer(b"ACGT-RYN", dtype=np.uint8), size=int(noise.sum()))
    cnt = (var != ord("-")).sum(axis=1).astype(np.uint64)
    keys = np.arange(n_rows, dtype=np.uint64)
    names = [f"s{i}" for i in range(n_samples)]
    g = S.MergeSkaArray.from_arrays(31, True, 64, keys, var, cnt, names)
    o = O.OracleArray.from_arrays(31, True, 64, list(range(n_rows)), var, cnt, names)
    assert g.distance(3.0, True) == o.distance_raw(3.0, True, threads=8)


# ------------------------------------------------------------------ host-packed input (skagpu_build_packed)
@pytest.mark.parametrize("k,rc", [(9, True), (15, False), (31, True)])
def test_packed_build_equals_ascii_build(k, rc):
    """2-bit codes + invalid mask from the host (0.375 B/base across PCIe) give the array of the ASCII path and of the oracle:
    N runs, lower case, IUPAC letters mis-encoded as bases, records of exactly k / k+1 bases, runs of exactly k at a record end
    followed by the 32-position gap between samples"""
    rng = np.random.default_rng(99 + k)
    core = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=k + 1))
    samples_records = [random_records(rng, int(rng.integers(1, 5)), 9000) + [core * 2] for _ in range(6)]
    samples_records.append([core[:k], core, core + b"N" + core[:k], core[:k] + b"N" + core, b"", core * 3])
    samples_records.append([bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=20000)) + b"N" + core[:k]])
    samples = [S.Sample.from_records(r) for r in samples_records]
    names = [f"s{i}" for i in range(len(samples))]
    packed = [S.PackedSample.from_sample(s) for s in samples]
    g_ascii = S.build_from_samples(samples, names, k, rc)
    g_packed = S.api.build_from_packed(packed, names, k, rc)
    g_batch = S.api.build_from_batch(S.SeqBatch(packed), names, k, rc)
    o = ora_from_gpu_samples(samples, names, k, rc)
    assert_same_array(g_ascii, o)
    assert_same_array(g_packed, o)
    assert_same_array(g_batch, o)


def test_packed_k1_records_single_sample():
    rng = np.random.default_rng(3)
    recs = random_records(rng, 7, 6000)
    s = S.Sample.from_records(recs)
    batch = S.SeqBatch([S.PackedSample.from_sample(s)])
    for k in (31, 9):
        rec, _ = batch.k1_extract(k, rc=True)
        exp = expected_records([recs], k, True)
        assert np.array_equal(rec[:len(exp)], exp)
        assert np.all(rec[len(exp):] == SENT)


def test_array_checksum_matches_numpy_and_flags_damage():
    """skagpu_array_checksum: device cell checksum == numpy restatement; count / order damage is detected"""
    rng = np.random.default_rng(17)
    recs = [random_records(rng, 3, 4000) for _ in range(5)]
    samples = [S.Sample.from_records(r) for r in recs]
    names = [f"s{i}" for i in range(5)]
    g = S.build_from_samples(samples, names, 15, True)
    keys, var, cnt = g.export()

    def mix(x):
        x = np.asarray(x, dtype=np.uint64)
        x = x ^ (x >> np.uint64(33)); x = x * np.uint64(0xff51afd7ed558ccd)
        x = x ^ (x >> np.uint64(33)); x = x * np.uint64(0xc4ceb9fe1a85ec53)
        return x ^ (x >> np.uint64(33))

    for lo, hi in ((0, 5), (1, 3), (4, 5)):
        tot = np.uint64(0)
        with np.errstate(over="ignore"):
            for c in range(lo, hi):
                m = var[:, c] != ord("-")
                cell = mix(keys[m] ^ mix((np.uint64(c - lo) << np.uint64(8)) | var[m, c].astype(np.uint64)))
                tot = tot + cell.sum(dtype=np.uint64)
        rows = int((var[:, lo:hi] != ord("-")).any(axis=1).sum())
        assert g.checksum(lo, hi) == (int(tot), rows, 0, 0)
    cnt2 = cnt.copy(); cnt2[3] += 1
    bad = S.MergeSkaArray.from_arrays(15, True, 64, keys, var, cnt2, names)
    assert bad.checksum()[2] == 1
    keys2 = keys.copy(); keys2[[5, 6]] = keys2[[6, 5]]
    bad = S.MergeSkaArray.from_arrays(15, True, 64, keys2, var, cnt, names)
    assert bad.checksum()[3] >= 1


# ------------------------------------------------------------------ joins on the key: merge / delete / weed / map
def _rand_samples(seed, n, k, rc, shared=0.7):
    """samples that share most split k-mers (a mutated common ancestor) so that joins have hits and misses"""
    rng = np.random.default_rng(seed)
    anc = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=6000)
    recs = []
    for _ in range(n):
        g = anc.copy()
        pos = rng.integers(0, len(g), size=int(len(g) * (1 - shared) / 20))
        g[pos] = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=len(pos))
        cut = int(rng.integers(1000, 5000))
        recs.append([g[:cut].tobytes(), g[cut:].tobytes()])
    samples = [S.Sample.from_records(r) for r in recs]
    names = [f"s{i}" for i in range(n)]
    return recs, samples, names


@pytest.mark.parametrize("k,rc", [(9, True), (21, False), (31, True)])
def test_merge_arrays_equals_joint_build(k, rc):
    """MergeSkaDict::extend: build(A) (+) build(B) == build(A + B), also when one side arrives in hash order (unsorted rows)"""
    recs, samples, names = _rand_samples(50 + k, 7, k, rc)
    a = S.build_from_samples(samples[:3], names[:3], k, rc)
    b = S.build_from_samples(samples[3:], names[3:], k, rc)
    joint = S.build_from_samples(samples, names, k, rc)
    m = S.api.merge_arrays(a, b)
    jk, jv, jc = joint.export()
    mk, mv, mc = m.export()
    assert np.array_equal(jk, mk) and np.array_equal(jv, mv) and np.array_equal(jc, mc)
    assert m.names() == names
    # the same with B's rows shuffled (as a reference-written .skf would deliver them)
    bk, bv, bc = b.export()
    perm = np.random.default_rng(1).permutation(len(bk))
    b2 = S.MergeSkaArray.from_arrays(k, rc, 64, bk[perm], bv[perm], bc[perm], names[3:])
    m2 = S.api.merge_arrays(a, b2)
    mk, mv, mc = m2.export()
    assert np.array_equal(jk, mk) and np.array_equal(jv, mv) and np.array_equal(jc, mc)
    # mismatching k / strand panic in the reference
    other = S.build_from_samples(samples[:2], names[:2], k + 2 if k < 31 else 29, rc)
    with pytest.raises(S.SkaError, match="K-mer lengths do not match"):
        S.api.merge_arrays(a, other)


def test_delete_samples_equals_build_without_them():
    k, rc = 15, True
    recs, samples, names = _rand_samples(7, 6, k, rc)
    full = S.build_from_samples(samples, names, k, rc)
    full.delete_samples(["s1", "s4"])
    keep = [0, 2, 3, 5]
    ref = S.build_from_samples([samples[i] for i in keep], [names[i] for i in keep], k, rc)
    assert full.names() == ref.names()
    fk, fv, fc = full.export()
    rk, rv, rc_ = ref.export()
    assert np.array_equal(fk, rk) and np.array_equal(fv, rv) and np.array_equal(fc, rc_)
    with pytest.raises(S.SkaError):
        full.delete_samples(["nope"])
    with pytest.raises(S.SkaError):
        full.delete_samples(full.names())


@pytest.mark.parametrize("reverse", [False, True])
def test_weed_matches_set_difference(reverse):
    k, rc = 17, True
    recs, samples, names = _rand_samples(11, 4, k, rc)
    arr = S.build_from_samples(samples, names, k, rc)
    keys, var, cnt = arr.export()
    weed_recs = [recs[0][0][100:900], recs[2][1][:700] + b"N" + recs[1][0][50:400], b"ACGT"]
    wkeys = set()
    n_w = 0
    for r in weed_recs:
        w = O.windows(r, k, rc=rc)
        wkeys.update(w["key"])
        n_w += w["n"]
    n = arr.weed(S.Sample.from_records(weed_recs), reverse=reverse)
    assert n == n_w
    found = np.array([int(x) in wkeys for x in keys])
    sel = found if reverse else ~found
    k2, v2, c2 = arr.export()
    assert np.array_equal(k2, keys[sel]) and np.array_equal(v2, var[sel]) and np.array_equal(c2, cnt[sel])


@pytest.mark.parametrize("k,rc", [(9, True), (15, False), (31, True)])
def test_map_lookup_matches_dictionary_walk(k, rc):
    """RefSka::map: every reference split k-mer in reference order, rows complemented (RC_IUPAC) on reverse-strand hits"""
    recs, samples, names = _rand_samples(23 + k, 5, k, rc)
    arr = S.build_from_samples(samples, names, k, rc)
    rows = arr.row_map()
    ref_recs = [recs[1][0][:1500] + b"NN" + recs[3][1][:800], recs[0][1][200:1200]]
    comp = {ord(a): ord(b) for a, b in zip("ABCDGHKMNRSTVWY", "TVGHCDMKNYSABWR")}
    rc_tab = np.full(256, ord("-"), dtype=np.uint8)
    for a, b in comp.items():
        rc_tab[a] = b
        rc_tab[a | 0x20] = b
    exp_pos, exp_rows, n_ref, base = [], [], 0, 0
    for r in ref_recs:
        w = O.windows(r, k, rc=rc)
        n_ref += w["n"]
        for i in range(w["n"]):
            key = w["key"][i]
            if key in rows:
                row = np.frombuffer(rows[key], dtype=np.uint8)
                exp_rows.append(rc_tab[row] if w["is_rc"][i] else row)
                exp_pos.append(base + int(w["mid_pos"][i]))
        base += len(r)
    pos, var, n = arr.map(S.Sample.from_records(ref_recs))
    assert n == n_ref
    assert [int(p) for p in pos] == exp_pos
    assert np.array_equal(var, np.array(exp_rows, dtype=np.uint8).reshape(len(exp_pos), len(names)))


# ------------------------------------------------------------------ k > 31: 128-bit split k-mers (src/lib.rs:607-622)
@pytest.mark.parametrize("k,rc", [(33, True), (41, True), (41, False), (63, True)])
def test_wide_k1_windows_match_rolling_iterator(k, rc):
    rng = np.random.default_rng(4321 + k)
    samples_records = [random_records(rng, int(rng.integers(1, 5)), 3000) for _ in range(4)]
    core = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=k + 1))
    samples_records.append([core[:k], core, core[:k - 1], b"", core + b"N" + core[:k], core[:k] + b"N" + core, b"N" * 70, core * 3])
    long = bytearray(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=9000).tobytes())
    for p in (4090, 4096, 4097, 8191):
        long[p] = ord("N")
    samples_records.append([bytes(long)])
    batch = S.SeqBatch([S.Sample.from_records(r) for r in samples_records])
    rec, _ms = batch.k1_extract(k, rc=rc)
    rec = rec.reshape(-1, 2)
    got = [(int(lo) | (int(hi) << 64)) for lo, hi in rec]
    exp = []
    h = (k - 1) // 2
    for recs in samples_records:
        for r in recs:
            e = [(1 << 128) - 1] * len(r)
            w = O.windows(r, k, rc=rc)
            for i in range(w["n"]):
                e[int(w["mid_pos"][i]) - h] = (w["key"][i] << 4) | (1 << int(w["mid"][i]))
            exp.extend(e)
    assert len(got) == len(exp)
    bad = [i for i, (a, b) in enumerate(zip(got, exp)) if a != b]
    assert not bad, f"{len(bad)} mismatching positions, first at {bad[:5]}"


def assert_same_array_wide(g, o):
    gd, od = g.dims(), o.dims()
    assert gd == od, (gd, od)
    gk, gv, gc = g.export()
    ok, ov, oc = o.export()
    assert [int(x) for x in gk] == ok
    assert np.array_equal(gv, ov) and np.array_equal(gc, oc)


@pytest.mark.parametrize("k,rc", [(33, True), (41, False), (51, True), (63, True)])
def test_wide_build_matches_oracle(k, rc):
    recs, samples, names = _rand_samples(900 + k, 6, k, rc)
    # repeats and IUPAC merges: a duplicated stretch with one changed middle base, a palindromic stretch
    extra = recs[0][0][:300]
    mutated = bytearray(extra); mutated[150] = ord("A") if extra[150] != ord("A") else ord("C")
    recs[2].append(extra + bytes(mutated))
    samples = [S.Sample.from_records(r) for r in recs]
    g = S.build_from_samples(samples, names, k, rc)
    o = ora_from_gpu_samples(samples, names, k, rc)
    assert g.dims()["key_bits"] == 128
    assert_same_array_wide(g, o)
    # the batch path (device-resident input) and forced bucket splitting take the same answer
    g2 = S.api.build_from_batch(S.SeqBatch(samples), names, k, rc)
    assert_same_array_wide(g2, o)
    os.environ["SKA_WIDE_LIMIT"] = "8"
    os.environ["SKA_PREFIX_BITS_WIDE"] = "3"
    try:
        g3 = S.build_from_samples(samples, names, k, rc)
    finally:
        del os.environ["SKA_WIDE_LIMIT"], os.environ["SKA_PREFIX_BITS_WIDE"]
    assert_same_array_wide(g3, o)


def test_wide_build_reproduces_reference_k41_skf(gold_in):
    """merge_k41.skf = `ska build -k 41 test_1.fa test_2.fa` written by the reference (u128 keys)"""
    ref = skf_py.load_skf(gold_in("merge_k41.skf"))
    var = np.frombuffer(ref["variants"], dtype=np.uint8).reshape(ref["n_rows"], ref["n_samples"])
    want = {k_: (bytes(var[i]), ref["variant_count"][i]) for i, k_ in enumerate(ref["split_kmers"])}
    g = gpu_build_files([gold_in("test_1.fa"), gold_in("test_2.fa")], k=41)
    gk, gv, gc = g.export()
    got = {int(k_): (bytes(gv[i]), int(gc[i])) for i, k_ in enumerate(gk)}
    assert got == want
    # distances and filters on the array we built == on the file the reference built
    buf = io.BytesIO()
    S.api.distance(g, buf)
    with open(os.path.join(os.path.dirname(gold_in("x")), "..", "correct", "merge_k41.dist.stdout"), "rb") as fh:
        assert buf.getvalue() == fh.read()


def test_wide_fastq_count_filter_k63(gold_in):
    """tests/fastq_input.rs:113-192 count_check_long: k = 63 reads, min-count 2 -> {GM}, min-count 3 -> {GA} (column sets)"""
    files = [(gold_in(f"test_long_{i}_fwd.fastq.gz"), gold_in(f"test_long_{i}_rev.fastq.gz")) for i in (1, 2)]
    for mc in (1, 2, 3):
        g = gpu_build_files(files, k=63, names=["a", "b"], min_count=mc)
        o = O.OracleArray.build_from_files(files, k=63, names=["a", "b"], min_count=mc)
        assert_same_array_wide(g, o)


# ---------------------------------------------------------------------------------- the library's own communicator (csrc/comm.cu)
def _threads_build_slabs(samples, names, k, world, bounds=None):
    """one host thread per device: build the rank's samples, exchange rows inside the library -> exported slabs in rank order"""
    import threading
    comm_id = S.api.comm_unique_id()
    n = len(samples)
    bounds = bounds or [(n * r // world, n * (r + 1) // world) for r in range(world)]
    out, errs = [None] * world, [None] * world

    def rank_main(r):
        try:
            S.set_device(r)
            S.api.comm_init(comm_id, world, r)
            try:
                lo, hi = bounds[r]
                local = S.build_from_samples(samples[lo:hi], names[lo:hi], k, True)
                slab, stats = S.api.exchange_rows(local, lo, names)
                tot = S.api.allreduce_u64([slab.dims()["n_rows"]])
                exported = slab.export()
                constant = S.api.apply_filters(slab, 0.0, False, "no-const", False, False)
                out[r] = (exported, stats, int(tot[0]), slab.distance_sharded(constant, True))
            finally:
                S.api.comm_destroy()
        except Exception as e:                                # noqa: BLE001 -- reported by the main thread
            errs[r] = e

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return out


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_library_communicator_slabs_equal_single_gpu(world):
    """skagpu_comm_init + skagpu_exchange_rows with one host thread per GPU: the slabs in rank order are the single-GPU array"""
    if _n_devices() < world:
        pytest.skip(f"needs {world} GPUs")
    from ska_rust_b200 import synth
    genomes = synth.tree_genomes(12, 60_000, 12, seed=0x5CA20011)
    samples = [S.Sample(g, np.array([0, len(g)], dtype=np.uint64)) for g in genomes]
    names = [f"s{i}" for i in range(12)]
    S.set_device(0)
    ref = S.build_from_samples(samples, names, 31, True)
    rk, rv, rc_ = ref.export()
    parts = _threads_build_slabs(samples, names, 31, world)
    keys = np.concatenate([np.asarray(p[0][0], dtype=np.uint64) for p in parts])
    assert np.array_equal(keys, np.asarray(rk, dtype=np.uint64))
    assert np.array_equal(np.concatenate([p[0][1] for p in parts]), rv)
    assert np.array_equal(np.concatenate([p[0][2] for p in parts]), rc_)
    assert all(p[2] == len(rk) for p in parts)              # the all-reduce of the slab sizes
    # ska distance on the slabs (counters summed inside the library) == on the single-GPU array
    constant = S.api.apply_filters(ref, 0.0, False, "no-const", False, False)
    want = ref.distance(float(constant), True)
    for p in parts:
        got = [(float(a), float(b), int(c), int(d)) for a, b, c, d in p[3].tolist()]
        assert got == want
    if world > 1:
        assert sum(p[1]["rows_sent"] for p in parts) == sum(p[1]["rows_received"] for p in parts) > 0
