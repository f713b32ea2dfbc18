"""ctypes front-end of the CPU oracle (oracle/ska_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Nothing under ska.rust_b200/ imports this module.

Mirrors the reference's call sequence so tests read like tests/*.rs of bacpop/ska.rust:
  build_and_merge (merge_ska_dict.rs:354) -> MergeSkaArray::new (merge_ska_array.rs:166)
  apply_filters / align / distance       (generic_modes.rs:22-50, 112-189)
"""
import ctypes as C
import gzip
import math
import os
import subprocess

import numpy as np

from . import skf_py

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libska_oracle.so")
_lib = None

FILTERS = {"no-filter": 0, "no-const": 1, "no-ambig": 2, "no-ambig-or-const": 3}  # cli.rs:129-138
QUAL_FILTERS = {"no-filter": 0, "middle": 1, "strict": 2}                           # cli.rs:112-123


class _Sample(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("qual", C.c_void_p), ("rec_off", C.c_void_p),
                ("n_rec", C.c_uint64), ("n_rec_file1", C.c_uint64)]


class _Dist(C.Structure):
    _fields_ = [("distance", C.c_double), ("mismatch_prop", C.c_double),
                ("match_count", C.c_uint64), ("mismatch_count", C.c_uint64)]


def build_lib(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "ska_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build_lib()
        L = C.CDLL(_LIB_PATH)
        L.ora_last_error.restype = C.c_char_p
        L.ora_build.restype = C.c_void_p
        L.ora_build.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_uint16, C.c_uint8, C.c_int,
                                C.c_double, C.c_int]
        L.ora_from_host.restype = C.c_void_p
        L.ora_from_host.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
        L.ora_dims.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.ora_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_export_transposed.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_filter.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ora_distance.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int]
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_windows.restype = C.c_uint64
        L.ora_windows.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int, C.c_uint8, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ora_iupac.restype = C.c_uint8
        L.ora_iupac.argtypes = [C.c_uint8, C.c_uint8]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


# ------------------------------------------------------------------------------- input parsing
def parse_fastx(path):
    """needletail-like record reader: -> (is_fastq, [(seq bytes, qual bytes|None)]).
    gz by magic bytes; FASTA sequences may be wrapped (newlines stripped); FASTQ strict 4-line."""
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    recs = []
    if not data:
        raise OracleError("Invalid FASTA/Q record")
    if data[:1] == b">":
        for block in data.split(b"\n>"):
            lines = block.split(b"\n")
            seq = b"".join(l.rstrip(b"\r") for l in lines[1:])
            recs.append((seq, None))
        return False, recs
    if data[:1] == b"@":
        lines = data.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        if len(lines) % 4:
            raise OracleError("Invalid FASTA/Q record")
        for i in range(0, len(lines), 4):
            seq, qual = lines[i + 1].rstrip(b"\r"), lines[i + 3].rstrip(b"\r")
            if len(seq) != len(qual):
                raise OracleError("Invalid FASTA/Q record")
            recs.append((seq, qual))
        return True, recs
    raise OracleError("Invalid FASTA/Q record")


class SampleData:
    """One sample flattened for the C side: concatenated seq (+qual) and record offsets."""

    def __init__(self, records_file1, records_file2=None, is_reads=None):
        recs = list(records_file1) + list(records_file2 or [])
        self.n_rec = len(recs)
        self.n_rec_file1 = len(records_file1)
        if is_reads is None:
            is_reads = bool(recs) and recs[0][1] is not None
        self.is_reads = is_reads
        lens = np.array([len(r[0]) for r in recs], dtype=np.uint64)
        self.rec_off = np.zeros(self.n_rec + 1, dtype=np.uint64)
        np.cumsum(lens, out=self.rec_off[1:])
        self.seq = np.frombuffer(b"".join(r[0] for r in recs), dtype=np.uint8).copy()
        self.qual = (np.frombuffer(b"".join(r[1] for r in recs), dtype=np.uint8).copy() if is_reads else None)

    @classmethod
    def from_files(cls, file1, file2=None):
        is_fq, r1 = parse_fastx(file1)
        r2 = parse_fastx(file2)[1] if file2 else None
        return cls(r1, r2, is_reads=is_fq)


def sample_name(path):
    """io_utils.rs:31-46 -- strip dir and one of .fa/.fasta/.fastq/.fastq.gz (case-insensitive)."""
    import re
    m = re.match(r"^.+/(.+)\.(?i:fa|fasta|fastq|fastq\.gz)$", path) or re.match(r"^(.+)\.(?i:fa|fasta|fastq|fastq\.gz)$", path)
    return m.group(1) if m else path


# ------------------------------------------------------------------------------- the array
class OracleArray:
    """Restatement of MergeSkaArray<IntT> (merge_ska_array.rs:108-126) behind a C handle."""

    def __init__(self, handle, names, ska_version="0.5.2"):
        if not handle:
            raise OracleError(lib().ora_last_error().decode())
        self._h = handle
        self.names = list(names)
        self.ska_version = ska_version

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_free(self._h)
            self._h = None

    # ---- constructors
    @classmethod
    def build(cls, samples, names, k=31, rc=True, min_count=5, min_qual=20, qual_filter="strict",
              proportion_reads=None, threads=1):
        """build_and_merge + MergeSkaArray::new. `samples`: list[SampleData]."""
        arr = (_Sample * len(samples))()
        for i, s in enumerate(samples):
            arr[i].seq = s.seq.ctypes.data
            arr[i].qual = s.qual.ctypes.data if s.qual is not None else None
            arr[i].rec_off = s.rec_off.ctypes.data
            arr[i].n_rec = s.n_rec
            arr[i].n_rec_file1 = s.n_rec_file1
        qf = QUAL_FILTERS[qual_filter] if isinstance(qual_filter, str) else qual_filter
        h = lib().ora_build(arr, len(samples), k, int(rc), min_count, min_qual, qf,
                            float(proportion_reads) if proportion_reads else 0.0, threads)
        return cls(h, names)

    @classmethod
    def build_from_files(cls, files, k=31, rc=True, names=None, **kw):
        """files: list of path | (path1, path2)."""
        samples, nm = [], []
        for f in files:
            f1, f2 = (f, None) if isinstance(f, str) else f
            samples.append(SampleData.from_files(f1, f2))
            nm.append(sample_name(f1))
        return cls.build(samples, names or nm, k=k, rc=rc, **kw)

    @classmethod
    def from_arrays(cls, k, rc, key_bits, keys, variants, counts, names, ska_version="0.5.2"):
        """keys: list[int] (python ints, may exceed 64 bits); variants: (R,N) uint8; counts: (R,)"""
        variants = np.ascontiguousarray(variants, dtype=np.uint8)
        n_rows = len(keys)
        n_samples = len(names)
        lohi = np.zeros((n_rows, 2), dtype=np.uint64)
        for i, v in enumerate(keys):
            lohi[i, 0] = v & 0xFFFFFFFFFFFFFFFF
            lohi[i, 1] = v >> 64
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        h = lib().ora_from_host(k, int(rc), key_bits, n_rows, n_samples, lohi.ctypes.data, variants.ctypes.data,
                                counts.ctypes.data)
        return cls(h, names, ska_version)

    @classmethod
    def load(cls, path):
        """MergeSkaArray::load (merge_ska_array.rs:199-204): u64 first, else u128 (lib.rs:635-661)."""
        d = skf_py.load_skf(path)
        key_bits = 64 if all(v < (1 << 64) for v in d["split_kmers"]) else 128
        var = np.frombuffer(d["variants"], dtype=np.uint8).reshape(d["n_rows"], d["n_samples"])
        return cls.from_arrays(d["k"], d["rc"], key_bits, d["split_kmers"], var, d["variant_count"], d["names"],
                               d["ska_version"])

    # ---- accessors
    def dims(self):
        v = [C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_int(), C.c_uint32()]
        lib().ora_dims(self._h, *[C.byref(x) for x in v])
        return dict(n_rows=v[0].value, n_keys=v[1].value, n_samples=v[2].value, k=v[3].value, rc=bool(v[4].value),
                    key_bits=v[5].value)

    def export(self):
        """-> keys (list[int]), variants (R,N) uint8, counts (R,) uint64"""
        d = self.dims()
        lohi = np.zeros((d["n_keys"], 2), dtype=np.uint64)
        var = np.zeros((d["n_rows"], d["n_samples"]), dtype=np.uint8)
        cnt = np.zeros(d["n_rows"], dtype=np.uint64)
        lib().ora_export(self._h, lohi.ctypes.data, var.ctypes.data, cnt.ctypes.data)
        keys = [int(lo) | (int(hi) << 64) for lo, hi in lohi]
        return keys, var, cnt

    def row_map(self):
        """{key -> row bytes}: the order-free form in which .skf content is compared."""
        keys, var, _ = self.export()
        return {k: bytes(var[i]) for i, k in enumerate(keys)}

    # ---- operations
    def filter(self, min_count, filter_ambig_as_missing=False, filter_type="no-filter", mask_ambig=False,
               ignore_const_gaps=False, update_kmers=True):
        ft = FILTERS[filter_type] if isinstance(filter_type, str) else filter_type
        removed = C.c_int32()
        if lib().ora_filter(self._h, int(min_count), int(filter_ambig_as_missing), ft, int(mask_ambig),
                            int(ignore_const_gaps), int(update_kmers), C.byref(removed)):
            raise OracleError(lib().ora_last_error().decode())
        return removed.value

    def apply_filters(self, min_freq, filter_ambig_as_missing, filter_type, ambig_mask, ignore_const_gaps):
        """generic_modes.rs:112-131"""
        thr = int(math.ceil(self.dims()["n_samples"] * min_freq))
        return self.filter(thr, filter_ambig_as_missing, filter_type, ambig_mask, ignore_const_gaps, update_kmers=False)

    def distance_raw(self, constant, filt_ambig=True, threads=1):
        n = self.dims()["n_samples"]
        out = (_Dist * (n * (n - 1) // 2))()
        if lib().ora_distance(self._h, float(constant), int(filt_ambig), out, threads):
            raise OracleError(lib().ora_last_error().decode())
        return [(d.distance, d.mismatch_prop, d.match_count, d.mismatch_count) for d in out]

    def write_fasta(self):
        """merge_ska_array.rs:499-517 -> bytes"""
        d = self.dims()
        t = np.zeros((d["n_samples"], d["n_rows"]), dtype=np.uint8)
        lib().ora_export_transposed(self._h, t.ctypes.data)
        return b"".join(b">" + nm.encode() + b"\n" + bytes(t[i]) + b"\n" for i, nm in enumerate(self.names))

    # ---- mode drivers (generic_modes.rs)
    def align(self, filter_type="no-const", mask_ambig=False, ignore_const_gaps=False, min_freq=0.9,
              filter_ambig_as_missing=False):
        """generic_modes.rs:22-50 with cli.rs defaults -> FASTA bytes"""
        self.apply_filters(min_freq, filter_ambig_as_missing, filter_type, mask_ambig, ignore_const_gaps)
        return self.write_fasta()

    def distance(self, min_freq=0.0, filt_ambig=True, threads=1):
        """generic_modes.rs:136-189 -> TSV bytes"""
        n = self.dims()["n_samples"]
        if min_freq * n >= 1.0:
            self.apply_filters(min_freq, False, "no-filter", False, False)
        constant = self.apply_filters(0.0, False, "no-const", False, False)
        dists = self.distance_raw(float(constant), filt_ambig, threads)
        return format_distances(self.names, dists)

    def nk(self, full_info=False):
        """Display / Debug of MergeSkaArray (merge_ska_array.rs:649-698) -> str, as `ska nk` prints it"""
        d = self.dims()
        keys, var, _ = self.export()
        sample_kmers = [int((var[:, s] != ord("-")).sum()) for s in range(d["n_samples"])]
        out = (f"ska_version={self.ska_version}\nk={d['k']}\nk_bits={d['key_bits']}\nrc={'true' if d['rc'] else 'false'}\n"
               f"k-mers={d['n_rows']}\nsamples={d['n_samples']}\n"
               "sample_names=[" + ", ".join('"' + n + '"' for n in self.names) + "]\n"
               "sample_kmers=[" + ", ".join(str(c) for c in sample_kmers) + "]\n")
        out += "\n"  # println!("{ska_array}")  (lib.rs:813)
        if full_info:  # println!("{ska_array:?}")  (lib.rs:816)
            for i, key in enumerate(keys):
                up, lo = decode_kmer(d["k"], key)
                out += f"{up}\t{lo}\t" + ",".join(chr(b) for b in var[i]) + "\n"
            out += "\n"
        return out


def decode_kmer(k, key):
    """bit_encoding.rs:307-335"""
    h = (k - 1) // 2
    letters = "ACTG"
    up = "".join(letters[(key >> (2 * h + 2 * (h - 1 - i))) & 3] for i in range(h))
    lo = "".join(letters[(key >> (2 * (h - 1 - i))) & 3] for i in range(h))
    return up, lo


def format_distances(names, dists):
    """generic_modes.rs:175-188 + VariantDist Display (merge_ska_array.rs:55-65)"""
    lines = ["Sample1\tSample2\tDistance\tMismatches (proportion)\tMatch count\tMismatch count"]
    n = len(names)
    it = iter(dists)
    for i in range(n):
        for j in range(i + 1, n):
            dist, prop, m, mm = next(it)
            lines.append(f"{names[i]}\t{names[j]}\t{rust_fixed(dist, 2)}\t{rust_fixed(prop, 5)}\t{m}\t{mm}")
    return ("\n".join(lines) + "\n").encode()


def rust_fixed(x, nd):
    """Rust `{:.N}`: correctly-rounded decimal (round-half-even on the exact binary value), like Python's."""
    return f"{x:.{nd}f}"


def var_hash(aln_bytes):
    """tests/common/mod.rs:165-189 -- order-insensitive set of alignment columns"""
    lines = aln_bytes.decode().splitlines()
    seqs = [l for i, l in enumerate(lines) if i % 2 == 1]
    if not seqs:
        return set()
    return {tuple(s[i] for s in seqs) for i in range(len(seqs[0]))}


def aln_length(aln_bytes):
    lines = aln_bytes.decode().splitlines()
    return [len(l) for i, l in enumerate(lines) if i % 2 == 1]


def windows(seq, k, rc=True, qual=None, min_qual=0, qual_filter="no-filter", is_reads=False):
    """All windows of one record as the reference's rolling iterator yields them ->
    dict of numpy arrays (key as python ints)."""
    seq_a = np.frombuffer(bytes(seq), dtype=np.uint8).copy()
    n = len(seq_a)
    qual_a = np.frombuffer(bytes(qual), dtype=np.uint8).copy() if qual is not None else None
    lohi = np.zeros((max(n, 1), 2), dtype=np.uint64)
    mid = np.zeros(max(n, 1), dtype=np.uint8)
    isrc = np.zeros(max(n, 1), dtype=np.uint8)
    ispal = np.zeros(max(n, 1), dtype=np.uint8)
    midpos = np.zeros(max(n, 1), dtype=np.uint64)
    midok = np.zeros(max(n, 1), dtype=np.uint8)
    hsh = np.zeros(max(n, 1), dtype=np.uint64)
    qf = QUAL_FILTERS[qual_filter] if isinstance(qual_filter, str) else qual_filter
    cnt = lib().ora_windows(seq_a.ctypes.data, n, qual_a.ctypes.data if qual_a is not None else None, k, int(rc),
                            min_qual, qf, int(is_reads), lohi.ctypes.data, mid.ctypes.data, isrc.ctypes.data,
                            ispal.ctypes.data, midpos.ctypes.data, midok.ctypes.data, hsh.ctypes.data)
    keys = [int(lo) | (int(hi) << 64) for lo, hi in lohi[:cnt]]
    return dict(n=cnt, key=keys, mid=mid[:cnt], is_rc=isrc[:cnt], is_pal=ispal[:cnt], mid_pos=midpos[:cnt],
                mid_ok=midok[:cnt], hash=hsh[:cnt])
