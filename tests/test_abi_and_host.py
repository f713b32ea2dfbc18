"""CPU-side checks: the C-ABI library loads and exports every symbol include/skagpu.h declares (no
compute without a GPU), compute calls fail loudly without a device (no CPU fallback), and the
host-side helpers mirror the reference's io_utils."""
import ctypes
import os
import re

import numpy as np
import pytest

import ska_rust_b200 as S
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "skagpu.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skagpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(S.api.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in skagpu.h but not exported"
    assert sorted(S.api.EXPORTS) == names, "api.py binds a different symbol set than the header declares"


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    s = S.Sample.from_records([b"ACGTACGTACGTACGTACGT"])
    with pytest.raises(S.SkaError, match="no CUDA device|CUDA"):
        S.build_from_samples([s], ["x"], 9, True)
    with pytest.raises(S.SkaError):
        S.MergeSkaArray.from_arrays(9, True, 64, [1], np.array([[65]], dtype=np.uint8), [1], ["x"])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ska.rust_b200")
    for root, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                with open(os.path.join(root, f)) as fh:
                    txt = fh.read()
                assert "oracle" not in txt.replace("CPU oracle lives in oracle/", "").replace("(The CPU oracle", ""), \
                    f"{f} mentions the oracle: the product path must not depend on it"


def test_read_input_fastas_names():
    """io_utils.rs:31-46"""
    got = S.api.read_input_fastas(["a/b/test_1.fa", "x.FASTA", "reads.fastq.gz", "dir/odd.txt", "plain", "p/q.fa.gz"])
    assert [g[0] for g in got] == ["test_1", "x", "reads", "dir/odd.txt", "plain", "p/q.fa.gz"]
    assert [O.sample_name(p) for p in ["a/b/test_1.fa", "x.FASTA", "reads.fastq.gz", "dir/odd.txt"]] == \
           ["test_1", "x", "reads", "dir/odd.txt"]


def test_get_input_list(tmp_path):
    """io_utils.rs:116-146"""
    p = tmp_path / "rfile"
    p.write_text("s1\ta.fa\ns2 b_1.fastq.gz   b_2.fastq.gz\n")
    assert S.api.get_input_list(str(p), None) == [("s1", "a.fa", None), ("s2", "b_1.fastq.gz", "b_2.fastq.gz")]
    p.write_text("only_one_field\n")
    with pytest.raises(S.SkaError):
        S.api.get_input_list(str(p), None)


def test_parse_fastx_matches_oracle_reader(gold_in):
    for name in ["test_1.fa", "ambig_test_2.fa", "par_test/test_1.fa.gz", "test_1_fwd.fastq.gz", "test_long_2_rev.fastq.gz"]:
        fq, seq, qual, off = S.api.parse_fastx(gold_in(name))
        ofq, recs = O.parse_fastx(gold_in(name))
        assert fq == ofq and len(off) - 1 == len(recs)
        for i, (s, q) in enumerate(recs):
            assert seq[int(off[i]):int(off[i + 1])].tobytes() == s
            if q is not None:
                assert qual[int(off[i]):int(off[i + 1])].tobytes() == q


def test_parse_wrapped_fasta(tmp_path):
    p = tmp_path / "w.fa"
    p.write_bytes(b">r1 desc\nACGT\nAC\r\nGT\n>r2\n\nTTTT\n")
    fq, seq, qual, off = S.api.parse_fastx(str(p))
    assert not fq and qual is None
    assert [seq[int(off[i]):int(off[i + 1])].tobytes() for i in range(2)] == [b"ACGTACGT", b"TTTT"]


def test_synth_is_deterministic_and_shardable():
    from ska_rust_b200 import synth
    a = synth.star_genomes(4, 10_000, seed=5)
    b = synth.star_genomes(2, 10_000, seed=5, first=2)
    assert all(np.array_equal(x, y) for x, y in zip(a[2:], b))
    diff = (a[0] != a[1]).mean()
    assert 0.0005 < diff < 0.005


def test_pack_bases_matches_numpy():
    """skagpu_pack_bases (host helper, no device work) against a numpy restatement of encode_base / valid_base
    (bit_encoding.rs:34-36, 52-54), including unaligned record starts"""
    import numpy as np
    import ska_rust_b200 as S
    L = S.api.lib()
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKM-.>~^", dtype=np.uint8)
    total = 10007
    seq = rng.choice(alphabet, size=total)
    p2 = np.zeros((total + 3) // 4, dtype=np.uint8)
    inv = np.full((total + 7) // 8, 0xFF, dtype=np.uint8)
    cuts = [0, 1, 8, 13, 14, 100, 1003, 4096, 4099, 9000, total]      # records packed back to back, arbitrary alignment
    for a, b in zip(cuts[:-1], cuts[1:]):
        rec = np.ascontiguousarray(seq[a:b])
        L.skagpu_pack_bases(rec.ctypes.data, b - a, a, p2.ctypes.data, inv.ctypes.data)
    codes = (seq >> 1) & 3
    exp2 = np.zeros(len(p2) * 4, dtype=np.uint8)
    exp2[:total] = codes
    exp2 = (exp2[0::4] | (exp2[1::4] << 2) | (exp2[2::4] << 4) | (exp2[3::4] << 6)).astype(np.uint8)
    bad = np.ones(len(inv) * 8, dtype=np.uint8)
    bad[:total] = (seq & 0xF) == 14
    expi = np.packbits(bad, bitorder="little")
    assert np.array_equal(p2, exp2)
    assert np.array_equal(inv, expi)


def test_owner_plan_matches_python_plan():
    """the owner cut inside libskagpu (csrc/comm.cu, skagpu_plan_owners) is the arithmetic of distributed.plan_owners"""
    import ska_rust_b200 as S
    from ska_rust_b200 import distributed as D
    rng = np.random.default_rng(7)
    for world in (1, 2, 3, 8):
        for case in range(20):
            c = rng.integers(0, 1000, size=4096).astype(np.uint64)
            if case % 3 == 0:
                c[rng.integers(0, 4096, size=3500)] = 0
            if case == 1:
                c[:] = 0
            assert np.array_equal(S.api.plan_owners(c, world), D.plan_owners(c, world))
