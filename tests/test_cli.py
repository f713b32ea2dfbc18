"""The C++ host (`ska` CLI + .skf codec) driven like the reference's snapbox tests (tests/*.rs).
CPU part: the codec alone.  GPU part: build / align / distance / nk end to end against the goldens."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import skf_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKA = os.path.join(ROOT, "ska.rust_b200", "ska")


def run(args, cwd=None, check=True):
    p = subprocess.run([SKA] + args, cwd=cwd, capture_output=True)
    if check:
        assert p.returncode == 0, p.stderr.decode()
    return p


# ------------------------------------------------------------------ codec (no GPU)
@pytest.mark.parametrize("fixture", ["merge.skf", "merge_k9.skf", "merge_k41.skf", "multidist.skf"])
def test_cpp_codec_recode_matches_python_decoder(gold_in, tmp_path, fixture):
    out = tmp_path / "re.skf"
    run(["recode", gold_in(fixture), str(out)])
    a, b = skf_py.load_skf(gold_in(fixture)), skf_py.load_skf(str(out))
    assert a == b
    # the re-encoded CBOR payload is byte-identical to what ciborium wrote (same shortest-form heads)
    with open(gold_in(fixture), "rb") as f1, open(out, "rb") as f2:
        assert skf_py.snappy_frame_decode(f1.read()) == skf_py.snappy_frame_decode(f2.read())


def test_cpp_codec_rejects_damaged_files(tmp_path):
    """a .skf whose arrays disagree (fewer split k-mers than rows would make the device copy read past the key array) or
    whose array header announces more elements than the payload holds is refused with a message, not loaded"""
    good = skf_py.dump_skf_bytes(9, True, ["a", "b"], [5, 9, 12], b"ACGTAC", 3, 2, [2, 2, 2])
    ok = tmp_path / "ok.skf"
    ok.write_bytes(good)
    run(["recode", str(ok), str(tmp_path / "ok2.skf")])
    assert skf_py.load_skf(str(tmp_path / "ok2.skf")) == skf_py.load_skf(str(ok))
    short = tmp_path / "short.skf"
    short.write_bytes(skf_py.dump_skf_bytes(9, True, ["a", "b"], [5, 9], b"ACGTAC", 3, 2, [2, 2, 2]))
    p = run(["recode", str(short), str(tmp_path / "x.skf")], check=False)
    assert p.returncode != 0 and b"split k-mers" in p.stderr
    # an array head that promises 2^40 variant bytes in a payload of a few dozen
    payload = bytearray(skf_py.snappy_frame_decode(good))
    i = payload.index(b"ddata") + 5
    assert payload[i] == 0x86                                      # array(6)
    payload[i:i + 1] = b"\x9b" + (1 << 40).to_bytes(8, "big")
    huge = tmp_path / "huge.skf"
    huge.write_bytes(skf_py.snappy_frame_encode_raw(bytes(payload)))
    p = run(["recode", str(huge), str(tmp_path / "y.skf")], check=False)
    assert p.returncode != 0 and b"longer than the payload" in p.stderr


def test_cpp_codec_large_random_roundtrip(tmp_path):
    """several 64 KiB chunks, compressible and incompressible regions, u128 keys"""
    rng = np.random.default_rng(1)
    n_rows, n_samples = 40_000, 9
    var = rng.choice(np.frombuffer(b"ACGT-NRY", dtype=np.uint8), size=(n_rows, n_samples)).astype(np.uint8)
    var[:20000] = ord("A")
    keys = sorted(int(x) for x in rng.integers(0, 1 << 62, size=n_rows))
    keys[-1] = (1 << 100) + 12345
    cnt = [int(c) for c in (var != ord("-")).sum(axis=1)]
    names = [f"sample {i}" for i in range(n_samples)]
    src = tmp_path / "a.skf"
    src.write_bytes(skf_py.dump_skf_bytes(41, False, names, keys, var.tobytes(), n_rows, n_samples, cnt, "0.5.2", 128))
    out = tmp_path / "b.skf"
    run(["recode", str(src), str(out), "--threads", "4"])
    assert skf_py.load_skf(str(src)) == skf_py.load_skf(str(out))
    assert os.path.getsize(out) < os.path.getsize(src) // 2       # the C++ writer really compresses


def test_cpp_codec_small_variant_bytes_take_the_general_writer(tmp_path):
    """variant bytes below 24 are one-byte CBOR integers: the fixed `18 xx` stride does not apply (never written by ska,
    but a valid file); reader and writer must notice and use the general integer path"""
    rng = np.random.default_rng(2)
    n_rows, n_samples = 70_000, 3                                  # > 64 KiB of payload: several chunks
    var = rng.choice(np.array([0, 7, 23, 24, 45, 65, 200], dtype=np.uint8), size=(n_rows, n_samples)).astype(np.uint8)
    keys = sorted(int(x) for x in rng.integers(0, 1 << 40, size=n_rows))
    cnt = [int(c) for c in (var != 45).sum(axis=1)]
    src = tmp_path / "a.skf"
    src.write_bytes(skf_py.dump_skf_bytes(21, True, ["a", "b", "c"], keys, var.tobytes(), n_rows, n_samples, cnt, "0.5.2", 64))
    out = tmp_path / "b.skf"
    run(["recode", str(src), str(out), "--threads", "3"])
    assert skf_py.load_skf(str(src)) == skf_py.load_skf(str(out))
    with open(src, "rb") as f1, open(out, "rb") as f2:
        assert skf_py.snappy_frame_decode(f1.read()) == skf_py.snappy_frame_decode(f2.read())


def test_cpp_codec_scales_linearly_in_chunks():
    """64 MB of middle bases = ~2000 Snappy chunks: the decoder sizes its payload once (an earlier version re-copied
    the payload for every chunk and needed minutes for this)"""
    p = run(["codec-selftest", "500000", "128", "--threads", "4"])
    fields = p.stdout.decode().split()
    enc_s, dec_s = float(fields[fields.index("encode_s") + 1]), float(fields[fields.index("decode_s") + 1])
    assert dec_s < 20 and enc_s < 20, p.stdout


def test_cli_errors_without_touching_the_device(tmp_path):
    assert run(["build", "-k", "30", "-o", "x", "a.fa"], check=False).returncode == 2      # clap validator (cli.rs:38-47)
    assert run(["build", "-k", "65", "-o", "x", "a.fa"], check=False).returncode == 2
    assert run(["frobnicate"], check=False).returncode == 2
    bad = tmp_path / "bad.skf"
    bad.write_bytes(b"not a snappy stream")
    assert run(["nk", str(bad)], check=False).returncode == 101                            # load error -> panic (lib.rs:660)


# ------------------------------------------------------------------ end to end (GPU)
def g(gold_in, *names):
    return [gold_in(n) for n in names]


@pytest.mark.gpu
def test_cli_build_nk_align_distance(gold_in, gold_ok, tmp_path):
    wd = str(tmp_path)
    # tests/fasta_input.rs:12-32 align_n (also: no ".skf.skf")
    run(["build"] + g(gold_in, "N_test_1.fa", "N_test_2.fa") + ["-o", "N_test.skf"], cwd=wd)
    assert os.path.exists(tmp_path / "N_test.skf") and not os.path.exists(tmp_path / "N_test.skf.skf")
    assert run(["align", "N_test.skf"], cwd=wd).stdout == gold_ok("align_N.stdout")
    # tests/skf_ops.rs / merge_nk.stdout
    run(["build"] + g(gold_in, "test_1.fa", "test_2.fa") + ["-o", "merge"], cwd=wd)
    nk = run(["nk", "merge.skf"], cwd=wd).stdout.decode()
    assert nk.split("\n", 1)[1] == gold_ok("merge_nk.stdout").decode().split("\n", 1)[1]
    # the file we wrote is what the reference's loader expects: decode it independently, compare with the oracle build
    d = skf_py.load_skf(str(tmp_path / "merge.skf"))
    o = O.OracleArray.build_from_files(g(gold_in, "test_1.fa", "test_2.fa"), k=31)
    keys, var, cnt = o.export()
    assert d["split_kmers"] == keys and d["variants"] == var.tobytes() and d["variant_count"] == [int(c) for c in cnt]
    assert d["names"] == ["test_1", "test_2"] and d["k"] == 31 and d["rc"] is True and d["k_bits"] == 64
    # tests/fasta_input.rs:157-220 repeats; :222-291 palindromes
    run(["build", "-k", "9"] + g(gold_in, "dup_test_1.fa", "dup_test_2.fa") + ["-o", "dup_ss", "--single-strand"], cwd=wd)
    assert run(["align", "dup_ss.skf"], cwd=wd).stdout == gold_ok("dup_ss.stdout")
    run(["build", "-k", "15"] + g(gold_in, "palindrome_1.fa", "palindrome_2.fa") + ["-o", "otto", "-v"], cwd=wd)
    assert run(["align", "otto.skf", "--filter", "no-filter"], cwd=wd).stdout == gold_ok("palindrome.stdout")
    # one-shot align from FASTA (io_utils.rs:75-92) and -o
    run(["align"] + g(gold_in, "test_1.fa", "test_2.fa") + ["-o", "basic.aln", "--filter", "no-filter", "--min-freq", "0", "--threads", "2"], cwd=wd)
    assert O.aln_length((tmp_path / "basic.aln").read_bytes()) == [62, 62]


@pytest.mark.gpu
@pytest.mark.parametrize("skf,extra,golden", [
    ("merge.skf", ["-v", "--threads", "2"], "merge.dist.stdout"),
    ("merge_k41.skf", [], "merge_k41.dist.stdout"),
    ("merge_k9.skf", ["--allow-ambiguous"], "merge_k9.dist.stdout"),
    ("merge_k9.skf", [], "merge_k9_no_ambig.dist.stdout"),
    ("merge_k9.skf", ["--min-freq", "1"], "merge_k9_min_freq.dist.stdout"),
    ("multidist.skf", ["--min-freq", "0.9"], "multidist.minfreq.stdout"),
    ("multidist.skf", ["--allow-ambiguous"], "multidist.ambig.stdout"),
])
def test_cli_distance_goldens(gold_in, gold_ok, tmp_path, skf, extra, golden):
    """tests/distance.rs: the reference's own .skf fixtures through our loader, byte-exact stdout"""
    assert run(["distance", gold_in(skf)] + extra).stdout == gold_ok(golden)
    run(["distance", gold_in(skf), "-o", "d.txt"] + extra, cwd=str(tmp_path))
    assert (tmp_path / "d.txt").read_bytes() == gold_ok(golden)


@pytest.mark.gpu
def test_cli_rfile_and_failures(gold_in, tmp_path):
    """tests/align.rs:13-31 build_cli (-f rfile), :155-166 (k=65 fails)"""
    rfile = tmp_path / "rfile"
    rfile.write_text(f"test_1\t{gold_in('test_1.fa')}\ntest_2\t{gold_in('test_2.fa')}\n")
    run(["build", "-f", str(rfile), "-o", "basic_build_opts", "-v", "--threads", "2", "-k", "31"], cwd=str(tmp_path))
    assert os.path.exists(tmp_path / "basic_build_opts.skf")
    # k > 31 builds with 128-bit split k-mers (tests/align.rs:116-144: column set {CT, TA} at k = 33)
    run(["build", "-o", "x", "-k", "33", gold_in("test_1.fa"), gold_in("test_2.fa")], cwd=str(tmp_path))
    assert O.var_hash(run(["align", "x.skf"], cwd=str(tmp_path)).stdout) == {("C", "T"), ("T", "A")}
    assert run(["build", "-o", "x", gold_in("nonexistent.fa")], cwd=str(tmp_path), check=False).returncode == 101


@pytest.mark.gpu
def test_cli_fastq_rfile(gold_in, tmp_path):
    """tests/fastq_input.rs:12-55 align_fastq through the CLI: 3-column rfile = paired FASTQ, --min-count/--min-qual"""
    rfile = tmp_path / "rfile"
    rfile.write_text("".join(f"test_{i}\t{gold_in(f'test_{i}_fwd.fastq.gz')}\t{gold_in(f'test_{i}_rev.fastq.gz')}\n" for i in (1, 2)))
    run(["build", "-f", str(rfile), "-o", "reads", "--min-count", "2", "-v", "-k", "9", "--min-qual", "2"], cwd=str(tmp_path))
    reads = run(["align", "reads.skf"], cwd=str(tmp_path)).stdout
    run(["build", "-k", "9", gold_in("test_1.fa"), gold_in("test_2.fa"), "-o", "fasta_k9"], cwd=str(tmp_path))
    fasta = run(["align", "fasta_k9.skf"], cwd=str(tmp_path)).stdout
    assert O.var_hash(reads) == O.var_hash(fasta)
    # count_check (fastq_input.rs:59-109)
    rfile.write_text("".join(f"test_count_{i}\t{gold_in(f'test_count_{i}_fwd.fastq.gz')}\t{gold_in(f'test_count_{i}_rev.fastq.gz')}\n" for i in (1, 2)))
    run(["build", "-f", str(rfile), "-o", "c1", "--min-count", "1", "-k", "7"], cwd=str(tmp_path))
    assert O.var_hash(run(["align", "c1.skf"], cwd=str(tmp_path)).stdout) == {("C", "W")}
    run(["build", "-f", str(rfile), "-o", "c3", "--min-count", "3", "-k", "7"], cwd=str(tmp_path))
    assert O.var_hash(run(["align", "c3.skf"], cwd=str(tmp_path)).stdout) == {("C", "T")}


# ------------------------------------------------------------------ ska merge / delete / weed / map
def matches(out, golden):
    """snapbox stdout_matches_path: `[..]` in the golden matches anything up to the end of the line"""
    import re
    pat = "".join(".*" if part == "[..]" else re.escape(part) for part in re.split(r"(\[\.\.\])", golden.decode()))
    return re.fullmatch(pat, out.decode(), flags=re.S) is not None


def nk_sets(text):
    """`ska nk --full-info`: header lines exactly, k-mer lines as a set (row order is unspecified in the reference)"""
    head, _, body = text.partition("\n\n")
    return head, sorted(l for l in body.splitlines() if l)


@pytest.mark.gpu
def test_cli_merge_delete(gold_in, gold_ok, tmp_path):
    """tests/skf_ops.rs:11-92 merge_delete"""
    wd = str(tmp_path)
    run(["build", gold_in("test_1.fa"), "-o", "test_1"], cwd=wd)
    run(["build", gold_in("test_2.fa"), "-o", "test_2"], cwd=wd)
    run(["merge", "test_1.skf", "test_2.skf", "-o", "merge"], cwd=wd)
    assert os.path.exists(tmp_path / "merge.skf")
    assert matches(run(["nk", "merge.skf"], cwd=wd).stdout, gold_ok("merge_nk.stdout"))
    # merged file == a joint build, row for row
    joint = O.OracleArray.build_from_files([gold_in("test_1.fa"), gold_in("test_2.fa")], k=31)
    d = skf_py.load_skf(str(tmp_path / "merge.skf"))
    keys, var, cnt = joint.export()
    assert d["split_kmers"] == keys and d["variants"] == var.tobytes() and d["variant_count"] == [int(c) for c in cnt]
    assert run(["delete", "-s", "merge.skf", "test_3"], cwd=wd, check=False).returncode != 0
    test1_nk = run(["nk", "test_1.skf"], cwd=wd).stdout
    run(["delete", "-s", "merge.skf", "test_2"], cwd=wd)
    assert run(["nk", "merge.skf"], cwd=wd).stdout == test1_nk
    assert run(["delete", "-s", "merge.skf", "-f", gold_in("missing_delete.txt")], cwd=wd, check=False).returncode != 0


@pytest.mark.gpu
def test_cli_weed(gold_in, gold_ok, tmp_path):
    """tests/skf_ops.rs:165-257 weed (k <= 31 part)"""
    import shutil
    wd = str(tmp_path)
    shutil.copy(gold_in("merge.skf"), tmp_path / "merge.skf")
    run(["weed", "merge.skf", gold_in("weed.fa"), "-v"], cwd=wd)
    assert run(["align", "merge.skf"], cwd=wd).stdout == gold_ok("weed_align.stdout")
    run(["weed", "merge.skf", "--filter", "no-const", "--min-freq", "1"], cwd=wd)
    out = run(["nk", "merge.skf", "--full-info"], cwd=wd).stdout
    assert matches(out, gold_ok("weed_nk.stdout"))
    shutil.copy(gold_in("merge_k9.skf"), tmp_path / "merge_k9.skf")
    run(["weed", "merge_k9.skf", "--ambig-mask"], cwd=wd)
    assert matches(run(["nk", "merge_k9.skf"], cwd=wd).stdout, gold_ok("weed_nk_k9.stdout"))
    shutil.copy(gold_in("merge.skf"), tmp_path / "merge.skf")
    run(["weed", "merge.skf", gold_in("weed.fa"), "--reverse", "-v"], cwd=wd)
    assert run(["align", "merge.skf"], cwd=wd).stdout == gold_ok("weed_align_reverse.stdout")


@pytest.mark.gpu
def test_cli_map(gold_in, gold_ok, tmp_path):
    """tests/map.rs: map_aln, map_vcf, map_rev_comp, repeat_mask (k <= 31 cases)"""
    wd = str(tmp_path)
    ref, ref2 = gold_in("test_ref.fa"), gold_in("test_ref_two_chrom.fa")
    run(["map", ref, gold_in("test_1.fa"), gold_in("test_2.fa"), "-o", "map.aln"], cwd=wd)
    assert os.path.exists(tmp_path / "map.aln")
    assert run(["map", ref, gold_in("merge.skf"), "-v", "--threads", "2"], cwd=wd).stdout == gold_ok("map_aln.stdout")
    assert run(["map", ref, gold_in("merge_k9.skf")], cwd=wd).stdout == gold_ok("map_aln_k9.stdout")
    assert run(["map", ref, gold_in("merge_k9.skf"), "--ambig-mask", "-v"], cwd=wd).stdout == gold_ok("map_aln_k9_filter.stdout")
    assert run(["map", ref2, gold_in("merge.skf")], cwd=wd).stdout == gold_ok("map_aln_two_chrom.stdout")
    assert run(["map", ref, gold_in("test_1.fa"), gold_in("indel_test.fa")], cwd=wd).stdout == gold_ok("map_aln_indels.stdout")
    run(["build", "-k", "17", gold_in("ambig_test_1.fa"), gold_in("ambig_test_2.fa"), "-o", "ambig_map", "--single-strand"], cwd=wd)
    assert run(["map", gold_in("ambig_test_ref.fa"), "ambig_map.skf"], cwd=wd).stdout == gold_ok("map_aln_ambig.stdout")
    # vcf
    assert matches(run(["map", ref, gold_in("merge.skf"), "-f", "vcf"], cwd=wd).stdout, gold_ok("map_vcf.stdout"))
    assert matches(run(["map", ref2, gold_in("merge.skf"), "-f", "vcf"], cwd=wd).stdout, gold_ok("map_vcf_two_chrom.stdout"))
    assert matches(run(["map", ref, gold_in("test_1.fa"), gold_in("indel_test.fa"), "-f", "vcf"], cwd=wd).stdout, gold_ok("map_vcf_indels.stdout"))
    # reverse complement input == forward input; single strand
    run(["build", "-o", "rc_build", "-k", "9", gold_in("test_1.fa"), gold_in("test_2_rc.fa")], cwd=wd)
    rc_map = run(["map", ref, "rc_build.skf"], cwd=wd).stdout                 # (the reference's own comparison maps a file that does not
    fwd_map = run(["map", ref, gold_in("merge_k9.skf")], cwd=wd).stdout       # exist, tests/map.rs:187-196; here: same alignment but for the
    a, b = rc_map.split(b"\n")[1::2], fwd_map.split(b"\n")[1::2]             # record-end window that the reverse strand drops at the other end)
    assert a[0] == b[0] and a[1][:70] == b[1][:70]
    run(["build", "-o", "ss_map", "-k", "9", "--single-strand", gold_in("test_1.fa"), gold_in("test_2_rc.fa")], cwd=wd)
    assert run(["map", ref, "ss_map.skf"], cwd=wd).stdout == gold_ok("map_ss.stdout")
    assert matches(run(["map", ref, "ss_map.skf", "-f", "vcf"], cwd=wd).stdout, gold_ok("map_vcf_ss.stdout"))
    # repeat mask
    assert run(["map", ref, gold_in("merge_k9.skf"), "-v", "--repeat-mask"], cwd=wd).stdout == gold_ok("map_aln_k9.masked.stdout")
    assert matches(run(["map", ref, gold_in("merge_k9.skf"), "--repeat-mask", "--format", "vcf"], cwd=wd).stdout, gold_ok("map_vcf_k9.masked.stdout"))
    assert run(["map", ref2, gold_in("merge_k9.skf"), "--repeat-mask"], cwd=wd).stdout == gold_ok("map_all_repeats.masked.stdout")
    ref3 = gold_in("test_ref_two_chrom_repeats.fa")
    assert run(["map", ref3, gold_in("merge_k9.skf"), "--repeat-mask"], cwd=wd).stdout == gold_ok("map_aln_two_chrom.masked.stdout")
    assert matches(run(["map", ref3, gold_in("merge_k9.skf"), "--repeat-mask", "--format", "vcf"], cwd=wd).stdout,
                   gold_ok("map_vcf_two_chrom.masked.stdout"))


@pytest.mark.gpu
def test_cli_u128_build_merge_delete_weed_map(gold_in, gold_ok, tmp_path):
    """k > 31 end to end: tests/skf_ops.rs:94-163 merge_delete_u128, :259-289 weed at k = 41, tests/map.rs:93-117 map_u128,
    tests/align.rs k = 33 (k33.stdout)"""
    import shutil
    wd = str(tmp_path)
    run(["build", gold_in("test_1.fa"), "-o", "test_1", "-k", "41"], cwd=wd)
    run(["build", gold_in("test_2.fa"), "-o", "test_2", "-k", "41"], cwd=wd)
    run(["merge", "test_1.skf", "test_2.skf", "-o", "merge", "-v"], cwd=wd)
    # the merged file holds what the reference's merge_k41.skf holds
    a, b = skf_py.load_skf(str(tmp_path / "merge.skf")), skf_py.load_skf(gold_in("merge_k41.skf"))
    rows = lambda d: {k_: (d["variants"][i * d["n_samples"]:(i + 1) * d["n_samples"]], d["variant_count"][i]) for i, k_ in enumerate(d["split_kmers"])}
    assert rows(a) == rows(b) and a["k"] == 41 and a["k_bits"] == 128
    assert run(["delete", "-s", "merge.skf", "-f", gold_in("missing_delete.txt"), "-v"], cwd=wd, check=False).returncode != 0
    test1_nk = run(["nk", "test_1.skf"], cwd=wd).stdout
    run(["delete", "-s", "merge.skf", "-o", "merge_delete", "test_2", "-v"], cwd=wd)
    assert run(["nk", "merge_delete.skf"], cwd=wd).stdout == test1_nk
    # weed at k = 41
    run(["build", "-o", "build_k41", "-k", "41", gold_in("test_1.fa"), gold_in("test_2.fa"), "-v"], cwd=wd)
    run(["weed", "build_k41.skf", "--filter", "no-ambig-or-const", "--min-freq", "1", "-v"], cwd=wd)
    assert matches(run(["nk", "build_k41.skf", "--full-info", "-v"], cwd=wd).stdout, gold_ok("weed_nk_k41.stdout"))
    shutil.copy(gold_in("merge_k41.skf"), tmp_path / "m41.skf")
    # weed.fa holds nothing as long as 41 bases: RefSka::new panics "has no valid sequence" (ska_ref.rs:253-255)
    assert run(["weed", "m41.skf", gold_in("weed.fa")], cwd=wd, check=False).returncode == 101
    run(["weed", "m41.skf", gold_in("test_ref.fa")], cwd=wd)                # key lookups on a loaded (unsorted) u128 array
    # map against u128 arrays
    ref = gold_in("test_ref.fa")
    assert run(["map", ref, gold_in("merge_k41.skf"), "-v", "--threads", "2"], cwd=wd).stdout == gold_ok("map_aln_k41.stdout")
    assert matches(run(["map", ref, gold_in("merge_k41.skf"), "-f", "vcf"], cwd=wd).stdout, gold_ok("map_vcf_k41.stdout"))
    # k = 33 build + nk (tests/align.rs:116-153)
    run(["build", "-o", "k33", "-k", "33", gold_in("test_1.fa"), gold_in("test_2.fa")], cwd=wd)
    assert matches(run(["nk", "k33.skf"], cwd=wd).stdout, gold_ok("k33.stdout"))


# ------------------------------------------------------------------ ska cov
EXAMPLE_COUNTS = [44633459, 950672, 104410, 44137, 24170, 21232, 21699, 24145, 30696, 39210, 49878, 63683, 77690, 95147, 112416, 130307,
                  146531, 160932, 175130, 185113, 193149, 197468, 199189, 198235, 192150, 185565, 176362, 165455, 152487, 139495, 127036,
                  112803, 103080, 90425, 80637, 70960, 62698, 54949, 46744, 41240, 35591, 30025, 25856, 22105, 19405, 16668, 14780, 12620,
                  11074, 9807, 8517, 7731, 7112, 6846, 6126, 5696, 5233, 4779, 4288, 3873, 3519, 3406, 2994, 2859, 2650, 2394, 2376, 2260,
                  2233, 2050, 1859, 1863, 1792, 1777, 1773, 1738, 1648]


def test_mixture_fit_reproduces_reference_unit_test():
    """src/coverage.rs:369-397 test_fit_histogram: the 77-bin example histogram gives cutoff 9"""
    out = run(["fit-selftest"] + [str(c) for c in EXAMPLE_COUNTS]).stdout.decode()
    assert "cutoff 9 " in out, out
    assert "cutoff 1 " in run(["fit-selftest"]).stdout.decode()          # empty histogram: converged at once, cutoff 1 (what cov_check relies on)


@pytest.mark.gpu
def test_cli_cov_and_auto_min_count(gold_in, tmp_path):
    """tests/fastq_input.rs:473-509 cov_check, :512-560 build_auto_check"""
    wd = str(tmp_path)
    p = run(["cov", gold_in("test_1_fwd.fastq.gz"), gold_in("test_1_rev.fastq.gz"), "-k", "9", "-v"], cwd=wd)
    assert p.stdout.startswith(b"Count\tK_mers\tMixture_density\tComponent\n") and b"Estimated cutoff\t" in p.stderr
    run(["cov", gold_in("test_long_1_fwd.fastq.gz"), gold_in("test_long_1_rev.fastq.gz"), "-k", "33", "-v"], cwd=wd)
    assert run(["cov", gold_in("test_1.fa"), gold_in("test_2.fa"), "-k", "9", "-v"], cwd=wd, check=False).returncode != 0
    rfile = tmp_path / "rfile"
    rfile.write_text("".join(f"test_{i}\t{gold_in(f'test_{i}_fwd.fastq.gz')}\t{gold_in(f'test_{i}_rev.fastq.gz')}\n" for i in (1, 2)))
    run(["build", "-f", str(rfile), "-o", "auto", "--min-count", "auto", "-k", "9", "-v"], cwd=wd)
    assert os.path.exists(tmp_path / "auto.skf")
    assert run(["build", "-f", str(rfile), "-o", "bad", "--min-count", "70000", "-k", "9"], cwd=wd, check=False).returncode == 2


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 4])
def test_cli_build_sharded_over_gpus_writes_the_same_skf(tmp_path, gpus):
    """`ska build --gpus N` (one host thread per GPU, rows exchanged inside libskagpu over NCCL) writes byte for byte the
    .skf of the single-GPU build; `ska align --gpus N` straight from the FASTA files gives the same alignment"""
    import torch
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    from ska_rust_b200 import synth
    genomes = synth.tree_genomes(10, 50_000, 10, seed=0x5CA20012)
    files = []
    for i, g in enumerate(genomes):
        f = tmp_path / f"g{i}.fa"
        synth.write_fasta(str(f), f"g{i}", g)
        files.append(str(f))
    run(["build", "-o", str(tmp_path / "one"), "-k", "31"] + files)
    run(["build", "-o", str(tmp_path / "many"), "-k", "31", "--gpus", str(gpus)] + files)
    assert (tmp_path / "one.skf").read_bytes() == (tmp_path / "many.skf").read_bytes()
    a1 = run(["align"] + files).stdout
    a2 = run(["align", "--gpus", str(gpus)] + files).stdout
    assert a1 == a2 and len(a1) > 0
    # a sample without sequence on one of the ranks: the reference's panic text, no hang
    bad = tmp_path / "empty.fa"
    bad.write_text(">x\nNNNNNNNN\n")
    p = run(["build", "-o", str(tmp_path / "bad"), "--gpus", str(gpus)] + files + [str(bad)], check=False)
    assert p.returncode == 101 and b"has no valid sequence" in p.stderr
