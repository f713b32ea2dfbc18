"""Pins the CPU oracle against the reference's OWN fixtures and golden outputs.

Each test cites the reference test it restates (tests/*.rs of bacpop/ska.rust v0.5.2).  These run
without a GPU.  The oracle is only trusted by the GPU parity tests because this file is green.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import skf_py


def build(gold_in, names, **kw):
    return O.OracleArray.build_from_files([gold_in(n) for n in names], **kw)


def rfile(gold_in, prefix, fastq):
    """tests/common/mod.rs:90-150 create_rfile"""
    if fastq:
        return [(gold_in(f"{prefix}_{i}_fwd.fastq.gz"), gold_in(f"{prefix}_{i}_rev.fastq.gz")) for i in (1, 2)], \
               [f"{prefix}_1", f"{prefix}_2"]
    return [gold_in(f"{prefix}_{i}.fa") for i in (1, 2)], [f"{prefix}_1", f"{prefix}_2"]


# ------------------------------------------------------------------ .skf decode + build fixtures
SKF_RECIPES = {
    # fixture -> (k, files)   recipes: tests/distance.rs:85-102, tests/skf_ops.rs / fixture provenance SURVEY 8c
    "multidist.skf": (9, ["N_test_1.fa", "N_test_2.fa", "ambig_test_1.fa", "ambig_test_2.fa", "test_1.fa", "test_2.fa"]),
    "merge_k9.skf": (9, ["test_1.fa", "test_2.fa"]),
    "merge.skf": (17, ["test_1.fa", "test_2.fa"]),
    "merge_k41.skf": (41, ["test_1.fa", "test_2.fa"]),
}


@pytest.mark.parametrize("fixture", sorted(SKF_RECIPES))
def test_build_reproduces_skf_fixture(gold_in, fixture):
    """build_and_merge + MergeSkaArray::new reproduce the committed .skf files as {key -> row}."""
    k, files = SKF_RECIPES[fixture]
    ref = skf_py.load_skf(gold_in(fixture))
    assert ref["k"] == k and ref["rc"] is True
    ref_var = np.frombuffer(ref["variants"], dtype=np.uint8).reshape(ref["n_rows"], ref["n_samples"])
    ref_map = {key: bytes(ref_var[i]) for i, key in enumerate(ref["split_kmers"])}
    ref_cnt = {key: ref["variant_count"][i] for i, key in enumerate(ref["split_kmers"])}
    arr = build(gold_in, files, k=k)
    assert arr.names == ref["names"]
    keys, var, cnt = arr.export()
    assert keys == sorted(keys), "canonical row order is ascending key"
    assert {key: bytes(var[i]) for i, key in enumerate(keys)} == ref_map
    assert {key: int(cnt[i]) for i, key in enumerate(keys)} == ref_cnt
    d = arr.dims()
    assert d["key_bits"] == ref["k_bits"] == (64 if k <= 31 else 128)


def test_skf_roundtrip_python_codec(gold_in):
    for fixture in SKF_RECIPES:
        d = skf_py.load_skf(gold_in(fixture))
        blob = skf_py.dump_skf_bytes(d["k"], d["rc"], d["names"], d["split_kmers"], d["variants"], d["n_rows"],
                                     d["n_samples"], d["variant_count"], d["ska_version"], d["k_bits"])
        assert skf_py.load_skf_bytes(blob) == d


# ------------------------------------------------------------------ distance goldens (tests/distance.rs)
DIST_CASES = [
    ("merge.skf", {}, "merge.dist.stdout"),                                         # :10-23
    ("merge_k41.skf", {}, "merge_k41.dist.stdout"),                                 # :34-42
    ("merge_k9.skf", {"filt_ambig": False}, "merge_k9.dist.stdout"),                # :50-59 --allow-ambiguous
    ("merge_k9.skf", {}, "merge_k9_no_ambig.dist.stdout"),                          # :62-69
    ("merge_k9.skf", {"min_freq": 1.0}, "merge_k9_min_freq.dist.stdout"),           # :72-80
    ("multidist.skf", {}, "multidist.stdout"),                                      # :105-113
    ("multidist.skf", {"min_freq": 0.9}, "multidist.minfreq.stdout"),               # :116-123
    ("multidist.skf", {"filt_ambig": False}, "multidist.ambig.stdout"),             # :126-133
]


@pytest.mark.parametrize("skf,kw,golden", DIST_CASES)
def test_distance_golden(gold_in, gold_ok, skf, kw, golden):
    arr = O.OracleArray.load(gold_in(skf))
    assert arr.distance(threads=2, **kw) == gold_ok(golden)


def test_distance_from_fresh_build_matches_fixture(gold_in, gold_ok):
    """tests/distance.rs:85-113 multisample_dists builds multidist.skf itself, then `distance`."""
    k, files = SKF_RECIPES["multidist.skf"]
    assert build(gold_in, files, k=k).distance() == gold_ok("multidist.stdout")


# ------------------------------------------------------------------ align goldens (tests/fasta_input.rs)
def test_align_n(gold_in, gold_ok):                       # fasta_input.rs:12-32
    assert build(gold_in, ["N_test_1.fa", "N_test_2.fa"]).align() == gold_ok("align_N.stdout")


def test_repeats(gold_in, gold_ok):                       # fasta_input.rs:157-220
    ss = build(gold_in, ["dup_test_1.fa", "dup_test_2.fa"], k=9, rc=False)
    assert ss.align() == gold_ok("dup_ss.stdout")
    # `ska weed dup_ss.skf --filter no-const --min-freq 1` (no weed file): filter with floor(N*min_freq),
    # update_kmers=true (generic_modes.rs:243-262), then `nk --full-info`
    ss = build(gold_in, ["dup_test_1.fa", "dup_test_2.fa"], k=9, rc=False)
    ss.filter(2, False, "no-const", False, False, update_kmers=True)
    assert ss.nk(full_info=True).split("\n", 1)[1] == gold_ok("dup_ss_nk.stdout").decode().split("\n", 1)[1]
    assert build(gold_in, ["dup_test_1.fa", "dup_test_2.fa"], k=9).align() == gold_ok("dup_rc.stdout")


def test_palindromes(gold_in, gold_ok):                   # fasta_input.rs:222-291
    f = ["palindrome_1.fa", "palindrome_2.fa"]
    assert build(gold_in, f, k=15).align(filter_type="no-filter") == gold_ok("palindrome.stdout")
    assert build(gold_in, f, k=15, rc=False).align() == gold_ok("palindrome_norc.stdout")
    f = ["palindrome_reps_1.fa", "palindrome_reps_2.fa"]
    assert build(gold_in, f, k=15).align(filter_type="no-filter") == gold_ok("palindrome_reps.stdout")


def test_rev_comp(gold_in):                               # fasta_input.rs:60-154
    fwd = build(gold_in, ["test_1.fa", "test_2.fa"], k=15).align()
    rcv = build(gold_in, ["test_1.fa", "test_2_rc.fa"], k=15).align()
    assert O.var_hash(fwd) == O.var_hash(rcv)
    ss = build(gold_in, ["test_1.fa", "test_2_rc.fa"], k=15, rc=False).align()
    assert O.var_hash(ss) == set()
    k33 = build(gold_in, ["test_1.fa", "test_2.fa"], k=33, rc=False).align()
    assert O.var_hash(k33) == {("T", "A"), ("G", "A")}


# ------------------------------------------------------------------ tests/align.rs
def test_build_and_align(gold_in):                        # align.rs:86-115, :168-183
    assert O.var_hash(build(gold_in, ["test_1.fa", "test_2.fa"], k=15).align()) == {("A", "T"), ("C", "T")}
    assert O.var_hash(build(gold_in, ["test_1.fa", "test_2.fa"]).align()) == {("A", "T"), ("C", "T")}


def test_long_kmers(gold_in, gold_ok):                    # align.rs:117-166
    arr = build(gold_in, ["test_1.fa", "test_2.fa"], k=33)
    assert arr.nk().split("\n", 1)[1] == gold_ok("k33.stdout").decode().split("\n", 1)[1]
    assert O.var_hash(arr.align()) == {("C", "T"), ("T", "A")}
    with pytest.raises(O.OracleError):
        build(gold_in, ["test_1.fa", "test_2.fa"], k=65)
    with pytest.raises(O.OracleError):
        build(gold_in, ["test_1.fa", "test_2.fa"], k=30)


def test_merge_nk(gold_in, gold_ok):                      # skf_ops.rs merge_nk.stdout (k=31: 62 k-mers, 32/32)
    arr = build(gold_in, ["test_1.fa", "test_2.fa"], k=31)
    assert arr.nk().split("\n", 1)[1] == gold_ok("merge_nk.stdout").decode().split("\n", 1)[1]


def test_proportion_reads(gold_in, gold_ok):              # align.rs:33-60
    arr = build(gold_in, ["proportion_reads.fa"], k=17, rc=False, proportion_reads=0.5)
    got = arr.nk(full_info=True).splitlines()
    want = gold_ok("proportion_reads.stdout").decode().splitlines()
    assert got[1:9] == want[1:9]          # header block (line 0 is the wildcarded version)
    assert len(got) == len(want)          # two k-mer rows (wildcarded in the golden) + blank lines
    # only record 0 (seq1, 18 bases) is kept by step=2, giving exactly two k=17 windows
    assert arr.dims()["n_rows"] == 2


FILTER_CASES = [  # align.rs:186-347 on merge_k9.skf
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
def test_filters(gold_in, kw, kind, want):
    aln = O.OracleArray.load(gold_in("merge_k9.skf")).align(**kw)
    if kind == "len":
        assert O.aln_length(aln) == [want, want]
    else:
        assert O.var_hash(aln) == want


def test_parallel_align(gold_in):                         # align.rs:349-397
    d = gold_in("par_test")
    files = sorted(os.path.join(d, f) for f in os.listdir(d))
    assert len(files) == 45
    serial = O.OracleArray.build_from_files(files, k=15, threads=1).align()
    par = O.OracleArray.build_from_files(files, k=15, threads=4).align()
    assert O.var_hash(serial) == O.var_hash(par)
    assert serial == par  # stronger: canonical row order makes the bytes identical


# ------------------------------------------------------------------ tests/fastq_input.rs
def fq(gold_in, prefix, **kw):
    files, names = rfile(gold_in, prefix, True)
    return O.OracleArray.build_from_files(files, names=names, **kw)


def test_align_fastq(gold_in):                            # fastq_input.rs:12-55
    reads = fq(gold_in, "test", k=9, min_count=2, min_qual=2).align()
    fasta = build(gold_in, ["test_1.fa", "test_2.fa"], k=9).align()
    assert O.var_hash(reads) == O.var_hash(fasta)


def test_count_check(gold_in):                            # fastq_input.rs:59-109
    assert O.var_hash(fq(gold_in, "test_count", k=7, min_count=1).align()) == {("C", "W")}
    assert O.var_hash(fq(gold_in, "test_count", k=7, min_count=3).align()) == {("C", "T")}


def test_count_check_long(gold_in):                       # fastq_input.rs:113-192
    assert O.var_hash(fq(gold_in, "test_long", k=63, min_count=1).align()) == {("G", "M")}
    c3 = fq(gold_in, "test_long", k=63, min_count=3).align()
    assert O.var_hash(c3) == {("G", "A")}
    ss = fq(gold_in, "test_long", k=63, min_count=2, rc=False).align()
    assert O.var_hash(c3) == O.var_hash(ss)


def test_error_fastq(gold_in):                            # fastq_input.rs:278-470
    all_hash = O.var_hash(fq(gold_in, "test", k=9, min_count=3, min_qual=2).align())
    assert O.var_hash(fq(gold_in, "test_quality", k=9, min_count=5, qual_filter="no-filter").align()) == all_hash
    assert O.var_hash(fq(gold_in, "test_quality_base", k=9, min_count=5, qual_filter="middle", min_qual=5).align()) == all_hash
    all_hash.discard(("C", "T"))
    assert O.var_hash(fq(gold_in, "test_error", k=9, min_count=5, min_qual=2).align()) == all_hash
    assert O.var_hash(fq(gold_in, "test_quality", k=9, min_count=5, min_qual=30).align()) == all_hash
    assert O.var_hash(fq(gold_in, "test_quality_base", k=9, min_count=5, min_qual=5, qual_filter="strict").align()) == all_hash
    assert O.var_hash(fq(gold_in, "test_quality_base", k=9, min_count=5).align()) == all_hash


# ------------------------------------------------------------------ in-crate unit tests
def test_palindrome_state_machine():
    """ska_dict.rs:523-566: W/S/N transitions of add_palindrome_to_dict, through a real palindromic
    key: k=5 windows ACxGT (upper AC, lower GT: rc(ACGT)=ACGT) with middle A, T -> W; C, G -> S; mix -> N."""
    def col(seqs):
        recs = [(q, None) for q in seqs]
        arr = O.OracleArray.build([O.SampleData(recs)], ["s"], k=5)
        return arr.row_map()
    key = O.windows(b"ACAGTA", 5)["key"][0]
    assert O.windows(b"ACAGTA", 5)["is_pal"][0] == 1
    assert col([b"ACAGTA"])[key] == b"W"
    assert col([b"ACTGTA"])[key] == b"W"
    assert col([b"ACCGTA"])[key] == b"S"
    assert col([b"ACGGTA"])[key] == b"S"
    assert col([b"ACAGTA", b"ACTGTA"])[key] == b"W"
    assert col([b"ACAGTA", b"ACCGTA"])[key] == b"N"
    assert col([b"ACGGTA", b"ACTGTA"])[key] == b"N"
    assert col([b"ACGGTA", b"ACTGTA", b"ACAGTA"])[key] == b"N"


def test_iupac_table_documented_rows():
    """bit_encoding.rs:354-369 comment block: every documented `new + existing -> code` row."""
    rows = {"A": "A M R W R H V W D M N D H V N".split(), "C": "M C S Y V Y S H B M B N H V N".split(),
            "G": "R S G K R B S D K V B D N V N".split(), "T": "W Y K T D Y B W K H B D H N N".split()}
    existing = "A C G T R Y S W K M B D H V N".split()
    enc = {"A": 0, "C": 1, "T": 2, "G": 3}
    for nb, outs in rows.items():
        for e, want in zip(existing, outs):
            assert chr(O.lib().ora_iupac(enc[nb], ord(e))) == want, (nb, e)
            assert chr(O.lib().ora_iupac(enc[nb], ord(e.lower()))) == want
    assert O.lib().ora_iupac(0, ord("-")) == 0 and O.lib().ora_iupac(0, ord("U")) == 0


def test_base_to_prob_overlaps(gold_in):
    """bit_encoding.rs:520-556 (overlap values) through the --allow-ambiguous distance of 1-row arrays."""
    def dist(a, b):
        arr = O.OracleArray.from_arrays(9, True, 64, [1], np.array([[ord(a), ord(b)]], dtype=np.uint8), [2], ["x", "y"])
        return arr.distance_raw(0.0, filt_ambig=False)[0][0]
    assert dist("A", "A") == 0.0 and dist("A", "C") == 1.0
    assert dist("A", "R") == 0.5 and dist("R", "R") == 0.5
    assert abs(dist("B", "A") - 1.0) < 1e-15 and abs(dist("B", "C") - (1 - 1 / 3)) < 1e-15
    assert dist("N", "A") == 1.0 and dist("N", "N") == 1.0


def test_window_rule_run_of_exactly_k_at_record_end():
    """split_kmer.rs:89,121 -- `idx + k >= seq_len` drops a trailing run of exactly k bases."""
    assert O.windows(b"ACGTACGTA", 9)["n"] == 0                 # L == k
    assert O.windows(b"ACGTACGTAC", 9)["n"] == 2                # L == k+1
    assert O.windows(b"ACGTACGTACNACGTACGTA", 9)["n"] == 2      # run of k after N at the end is dropped
    assert O.windows(b"ACGTACGTANACGTACGTAC", 9)["n"] == 3      # run of exactly k *before* an N survives (1) + 2
    assert O.windows(b"NACGTACGTA", 9)["n"] == 0                # leading N then exactly k
    assert O.windows(b"acgtacgtacn", 9)["n"] == 2               # lower case accepted, trailing n
