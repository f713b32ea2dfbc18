"""Host-side mirror of the reference's hot-path interface, bound to libskagpu.so over its C ABI.

Names, argument meaning and error behaviour follow the Rust API this path replaces:

  build_and_merge(...)            src/merge_ska_dict.rs:354-417 (+ MergeSkaArray::new, merge_ska_array.rs:166)
  MergeSkaArray.filter / distance / write_fasta / load / save
                                  src/merge_ska_array.rs:289-402, 416-438, 499-517, 191-204
  apply_filters / align / distance  src/generic_modes.rs:112-131, 22-50, 136-189
  read_input_fastas / get_input_list  src/io_utils.rs:31-46, 116-146

There is NO CPU fallback: if libskagpu.so is missing or no B200 is visible every compute call
raises SkaError.  (The CPU oracle lives in oracle/ and is test infrastructure only.)
"""
import ctypes as C
import math
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SKAGPU_LIB: another build of the same library (kernel tuning experiments, profiles/); never a different implementation
LIB_PATH = os.environ.get("SKAGPU_LIB") or os.path.join(_HERE, "libskagpu.so")
_lib = None

FILTER_TYPES = {"no-filter": 0, "no-const": 1, "no-ambig": 2, "no-ambig-or-const": 3}   # cli.rs:129-138
QUAL_FILTERS = {"no-filter": 0, "middle": 1, "strict": 2}                                # cli.rs:112-123
SKA_VERSION = "0.5.2"   # version stamped into arrays we build (merge_ska_array.rs:183, Cargo.toml:3)

EXPORTS = [
    "skagpu_last_error", "skagpu_set_device", "skagpu_build", "skagpu_seqbatch_upload", "skagpu_build_from_batch",
    "skagpu_seqbatch_free", "skagpu_array_from_host", "skagpu_dims", "skagpu_filter", "skagpu_distance", "skagpu_export",
    "skagpu_export_transposed", "skagpu_free", "skagpu_host_alloc", "skagpu_host_free", "skagpu_last_timings",
    "skagpu_launch_count", "skagpu_k1_extract", "skagpu_stream_timer", "skagpu_extract_records", "skagpu_records_info",
    "skagpu_records_free", "skagpu_merge_records_dev", "skagpu_distance_counters", "skagpu_merge_prefix_bits",
    "skagpu_array_dev", "skagpu_key_cumulative", "skagpu_merge_partial_rows_dev", "skagpu_array_checksum",
    "skagpu_pack_bases", "skagpu_build_packed", "skagpu_seqbatch_upload_packed", "skagpu_export_transposed_async",
    "skagpu_export_wait", "skagpu_seqbatch_upload_packed_async", "skagpu_merge_arrays", "skagpu_delete_samples", "skagpu_weed",
    "skagpu_map", "skagpu_map_dims", "skagpu_map_export", "skagpu_map_free", "skagpu_kmer_count_hist",
    "skagpu_comm_unique_id", "skagpu_comm_init", "skagpu_comm_destroy", "skagpu_exchange_rows", "skagpu_allreduce_u64",
    "skagpu_plan_owners", "skagpu_distance_sharded",
]


class SkaError(RuntimeError):
    """Raised where the reference panics / returns Err (exit code 101 at the CLI)."""


class _Input(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("qual", C.c_void_p), ("rec_off", C.c_void_p), ("n_rec", C.c_uint64),
                ("n_rec_file1", C.c_uint64)]


class _Dist(C.Structure):
    _fields_ = [("distance", C.c_double), ("mismatch_prop", C.c_double), ("match_count", C.c_uint64),
                ("mismatch_count", C.c_uint64)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SkaError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.skagpu_last_error.restype = C.c_char_p
        L.skagpu_set_device.argtypes = [C.c_int]
        L.skagpu_build.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_uint16, C.c_uint8, C.c_int, C.c_double,
                                   C.c_void_p]
        L.skagpu_seqbatch_upload.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_void_p]
        L.skagpu_build_from_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint16, C.c_uint8, C.c_int, C.c_void_p]
        L.skagpu_seqbatch_free.argtypes = [C.c_void_p]
        L.skagpu_array_from_host.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
        L.skagpu_dims.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.skagpu_filter.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.skagpu_distance.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p]
        L.skagpu_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.skagpu_export_transposed.argtypes = [C.c_void_p, C.c_void_p]
        L.skagpu_export_transposed_async.argtypes = [C.c_void_p, C.c_void_p]
        L.skagpu_free.argtypes = [C.c_void_p]
        L.skagpu_host_alloc.restype = C.c_void_p
        L.skagpu_host_alloc.argtypes = [C.c_size_t]
        L.skagpu_host_free.argtypes = [C.c_void_p]
        L.skagpu_last_timings.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.skagpu_launch_count.restype = C.c_uint64
        L.skagpu_stream_timer.argtypes = [C.c_int, C.c_void_p]
        L.skagpu_extract_records.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
        L.skagpu_records_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.skagpu_records_free.argtypes = [C.c_void_p]
        L.skagpu_merge_records_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                               C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        L.skagpu_merge_prefix_bits.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.skagpu_array_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.skagpu_key_cumulative.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.skagpu_merge_partial_rows_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        L.skagpu_distance_counters.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.skagpu_array_checksum.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.skagpu_merge_arrays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.skagpu_delete_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.skagpu_weed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.skagpu_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.skagpu_map_dims.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.skagpu_map_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.skagpu_map_free.argtypes = [C.c_void_p]
        L.skagpu_kmer_count_hist.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.skagpu_comm_unique_id.argtypes = [C.c_void_p]
        L.skagpu_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.skagpu_exchange_rows.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.skagpu_allreduce_u64.argtypes = [C.c_void_p, C.c_size_t]
        L.skagpu_plan_owners.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.skagpu_distance_sharded.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p]
        L.skagpu_pack_bases.restype = None
        L.skagpu_pack_bases.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.skagpu_build_packed.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, C.c_void_p]
        L.skagpu_seqbatch_upload_packed.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.skagpu_seqbatch_upload_packed_async.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.skagpu_k1_extract.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint8, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc:
        raise SkaError(lib().skagpu_last_error().decode())


def set_device(device):
    _check(lib().skagpu_set_device(int(device)))


def export_wait():
    _check(lib().skagpu_export_wait())


def timer_start():
    _check(lib().skagpu_stream_timer(0, None))


def timer_stop():
    """-> device milliseconds since timer_start(), from CUDA events on the library's stream"""
    ms = C.c_float()
    _check(lib().skagpu_stream_timer(1, C.byref(ms)))
    return float(ms.value)


def launch_count():
    return int(lib().skagpu_launch_count())


def last_timings():
    """[(stage, ms, kernel launches)] of the most recent build, from CUDA events on the library stream."""
    names = (C.c_char_p * 32)()
    ms = (C.c_float * 32)()
    ln = (C.c_uint64 * 32)()
    n = lib().skagpu_last_timings(names, ms, ln, 32)
    return [(names[i].decode(), float(ms[i]), int(ln[i])) for i in range(n)]


# ------------------------------------------------------------------------------------ input side
class QualOpts:
    """lib.rs:533-540"""

    def __init__(self, min_count=5, min_qual=20, qual_filter="strict"):
        self.min_count, self.min_qual, self.qual_filter = min_count, min_qual, qual_filter


def read_input_fastas(seq_files):
    """io_utils.rs:31-46 -> [(name, file1, None)]"""
    out = []
    for f in seq_files:
        m = (re.match(r"^.+/(.+)\.(?i:fa|fasta|fastq|fastq\.gz)$", f) or
             re.match(r"^(.+)\.(?i:fa|fasta|fastq|fastq\.gz)$", f))
        out.append((m.group(1) if m else f, f, None))
    return out


def get_input_list(file_list, seq_files):
    """io_utils.rs:116-146"""
    if file_list is None:
        return read_input_fastas(seq_files)
    out = []
    with open(file_list) as fh:
        for line in fh.read().splitlines():
            fields = line.split()
            if len(fields) == 2:
                out.append((fields[0], fields[1], None))
            elif len(fields) == 3:
                out.append((fields[0], fields[1], fields[2]))
            else:
                raise SkaError("Unable to parse line in file_list")
    return out


def parse_fastx(path):
    """Record reader standing in for needletail (ska_dict.rs:131-147):
    -> (is_fastq, seq uint8[], qual uint8[]|None, rec_off uint64[n_rec+1])"""
    import gzip
    try:
        with open(path, "rb") as fh:
            data = fh.read()
    except OSError:
        raise SkaError(f"Invalid path/file: {path}")
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    if data[:1] == b">":
        seqs = []
        for block in data.split(b"\n>"):
            nl = block.find(b"\n")
            body = block[nl + 1:] if nl >= 0 else b""
            seqs.append(body.replace(b"\n", b"").replace(b"\r", b""))
        quals = None
        is_fq = False
    elif data[:1] == b"@":
        lines = data.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        if len(lines) % 4:
            raise SkaError("Invalid FASTA/Q record")
        seqs = [l.rstrip(b"\r") for l in lines[1::4]]
        quals = [l.rstrip(b"\r") for l in lines[3::4]]
        if any(len(a) != len(b) for a, b in zip(seqs, quals)):
            raise SkaError("Invalid FASTA/Q record")
        is_fq = True
    else:
        raise SkaError("Invalid FASTA/Q record")
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    seq = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    qual = np.frombuffer(b"".join(quals), dtype=np.uint8) if quals is not None else None
    return is_fq, seq, qual, off


class Sample:
    """One InputFastx after parsing: what the C ABI's ska_input_t points at."""

    def __init__(self, seq, rec_off, qual=None, n_rec_file1=None):
        self.seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self.rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        self.qual = None if qual is None else np.ascontiguousarray(qual, dtype=np.uint8)
        self.n_rec = len(self.rec_off) - 1
        self.n_rec_file1 = self.n_rec if n_rec_file1 is None else n_rec_file1

    @classmethod
    def from_files(cls, file1, file2=None):
        fq, seq, qual, off = parse_fastx(file1)
        n1 = len(off) - 1
        if file2 is not None:
            _fq2, seq2, qual2, off2 = parse_fastx(file2)
            seq = np.concatenate([seq, seq2])
            if fq:
                qual = np.concatenate([qual, qual2 if qual2 is not None else np.full(len(seq2), 0xFF, np.uint8)])   # 0xFF: no quality, always valid
            off = np.concatenate([off, off2[1:] + off[-1]])
        return cls(seq, off, qual if fq else None, n1)

    @classmethod
    def from_records(cls, seqs):
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in seqs], out=off[1:])
        return cls(np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8), off)


class _InputPacked(C.Structure):
    _fields_ = [("packed2", C.c_void_p), ("invalid", C.c_void_p), ("rec_off", C.c_void_p), ("n_rec", C.c_uint64),
                ("n_rec_file1", C.c_uint64)]


class PackedSample:
    """One assembly as the packing FASTA reader hands it over (ska_input_packed_t): 2-bit codes, 4 bases per byte, and
    the invalid-base mask, 8 bases per byte; records back to back in base space."""

    def __init__(self, packed2, invalid, rec_off):
        self.packed2, self.invalid = packed2, invalid
        self.rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        self.n_rec = len(self.rec_off) - 1
        self.n_rec_file1 = self.n_rec
        self.n_bases = int(self.rec_off[-1])

    @classmethod
    def from_sample(cls, sample, packed2=None, invalid=None):
        """pack an ASCII Sample with the library's host helper (skagpu_pack_bases); optional preallocated (pinned) outputs"""
        n = int(sample.rec_off[-1] - sample.rec_off[0])
        p2 = np.zeros((n + 3) // 4, dtype=np.uint8) if packed2 is None else packed2
        inv = np.empty((n + 7) // 8, dtype=np.uint8) if invalid is None else invalid
        inv[:] = 0xFF
        seq = sample.seq[int(sample.rec_off[0]):]
        lib().skagpu_pack_bases(seq.ctypes.data, n, 0, p2.ctypes.data, inv.ctypes.data)
        return cls(p2, inv, sample.rec_off - sample.rec_off[0])


def _packed_inputs_array(samples):
    arr = (_InputPacked * len(samples))()
    for i, s in enumerate(samples):
        arr[i].packed2 = s.packed2.ctypes.data
        arr[i].invalid = s.invalid.ctypes.data
        arr[i].rec_off = s.rec_off.ctypes.data
        arr[i].n_rec = s.n_rec
        arr[i].n_rec_file1 = s.n_rec_file1
    return arr


def _inputs_array(samples):
    arr = (_Input * len(samples))()
    for i, s in enumerate(samples):
        arr[i].seq = s.seq.ctypes.data
        arr[i].qual = s.qual.ctypes.data if s.qual is not None else None
        arr[i].rec_off = s.rec_off.ctypes.data
        arr[i].n_rec = s.n_rec
        arr[i].n_rec_file1 = s.n_rec_file1
    return arr


class SeqBatch:
    """Sequence bytes of a set of samples resident in HBM (skagpu_seqbatch_upload)."""

    def __init__(self, samples, proportion_reads=None, asynchronous=False):
        h = C.c_void_p()
        self._keep = samples
        if samples and isinstance(samples[0], PackedSample):
            fn = lib().skagpu_seqbatch_upload_packed_async if asynchronous else lib().skagpu_seqbatch_upload_packed
            self._inputs = _packed_inputs_array(samples)
            _check(fn(self._inputs, len(samples), C.byref(h)))
        else:
            _check(lib().skagpu_seqbatch_upload(_inputs_array(samples), len(samples),
                                                float(proportion_reads) if proportion_reads else 0.0, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().skagpu_seqbatch_free(self._h)
            self._h = None

    def n_samples(self):
        return len(self._keep)

    def k1_extract(self, k, rc=True, min_qual=0, qual_filter="no-filter", iters=1, want_records=True):
        """Raw K1 output: one u64 per base position (key<<4 | one-hot middle; all-ones = no window)."""
        n = C.c_uint64()
        ms = C.c_float()
        qf = QUAL_FILTERS[qual_filter] if isinstance(qual_filter, str) else qual_filter
        if self._keep and isinstance(self._keep[0], PackedSample):      # position space with 32-aligned sample starts
            total = int(sum((s.n_bases + 31) // 32 * 32 for s in self._keep))
        else:
            total = int(sum(len(s.seq) for s in self._keep))
        rec = np.empty(total * (2 if k > 31 else 1), dtype=np.uint64) if want_records else None   # k > 31: (lo, hi) per position
        _check(lib().skagpu_k1_extract(self._h, k, int(rc), min_qual, qf, iters,
                                       rec.ctypes.data if rec is not None else None, C.byref(n), C.byref(ms)))
        return rec, float(ms.value)


# ------------------------------------------------------------------------------------ the array
class MergeSkaArray:
    """Device-resident MergeSkaArray<IntT> (merge_ska_array.rs:108-126)."""

    def __init__(self, handle, names, ska_version=SKA_VERSION):
        self._h = handle
        self._names = list(names)
        self.ska_version = ska_version

    def __del__(self):
        if getattr(self, "_h", None):
            lib().skagpu_free(self._h)
            self._h = None

    # ---- constructors
    @classmethod
    def from_arrays(cls, k, rc, key_bits, keys, variants, counts, names, ska_version=SKA_VERSION):
        """skagpu_array_from_host. keys: python ints or uint64 array (key_bits 64) / ints (128)."""
        variants = np.ascontiguousarray(variants, dtype=np.uint8)
        n_rows, n_samples = len(keys), len(names)
        if key_bits == 64:
            karr = np.array(keys, dtype=np.uint64) if not isinstance(keys, np.ndarray) else keys.astype(np.uint64)
        else:
            karr = np.zeros((n_rows, 2), dtype=np.uint64)
            for i, v in enumerate(keys):
                karr[i, 0] = int(v) & 0xFFFFFFFFFFFFFFFF
                karr[i, 1] = int(v) >> 64
        karr = np.ascontiguousarray(karr)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        h = C.c_void_p()
        _check(lib().skagpu_array_from_host(k, int(rc), key_bits, n_rows, n_samples, karr.ctypes.data,
                                            variants.ctypes.data, counts.ctypes.data, C.byref(h)))
        return cls(h, names, ska_version)

    # ---- accessors (merge_ska_array.rs:519-585)
    def dims(self):
        v = [C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_int(), C.c_uint32()]
        _check(lib().skagpu_dims(self._h, *[C.byref(x) for x in v]))
        return dict(n_rows=v[0].value, n_keys=v[1].value, n_samples=v[2].value, k=v[3].value, rc=bool(v[4].value),
                    key_bits=v[5].value)

    def names(self):
        return self._names

    def nsamples(self):
        return self.dims()["n_samples"]

    def ksize(self):
        return self.dims()["n_rows"]

    def kmer_len(self):
        return self.dims()["k"]

    def rc(self):
        return self.dims()["rc"]

    def export(self):
        """-> keys (uint64[n_keys] or list[int] for 128-bit), variants (R,N) uint8, counts uint64[R]"""
        d = self.dims()
        kw = d["key_bits"] // 64
        karr = np.zeros((d["n_keys"], kw), dtype=np.uint64)
        var = np.zeros((d["n_rows"], d["n_samples"]), dtype=np.uint8)
        cnt = np.zeros(d["n_rows"], dtype=np.uint64)
        _check(lib().skagpu_export(self._h, karr.ctypes.data, var.ctypes.data, cnt.ctypes.data))
        if kw == 1:
            keys = karr[:, 0]
        else:
            keys = [int(lo) | (int(hi) << 64) for lo, hi in karr]
        return keys, var, cnt

    def row_map(self):
        keys, var, _ = self.export()
        return {int(k): bytes(var[i]) for i, k in enumerate(keys)}

    # ---- MergeSkaArray::filter (merge_ska_array.rs:289-402)
    def filter(self, min_count, filter_ambig_as_missing, filter_type, mask_ambig, ignore_const_gaps, update_kmers):
        ft = FILTER_TYPES[filter_type] if isinstance(filter_type, str) else filter_type
        removed = C.c_int32()
        _check(lib().skagpu_filter(self._h, int(min_count), int(filter_ambig_as_missing), ft, int(mask_ambig),
                                   int(ignore_const_gaps), int(update_kmers), C.byref(removed)))
        return removed.value

    # ---- MergeSkaArray::distance (merge_ska_array.rs:416-438) -> [(distance, mismatch_prop, matches, mismatches)]
    def distance(self, constant, filt_ambig):
        n = self.nsamples()
        out = (_Dist * max(1, n * (n - 1) // 2))()
        _check(lib().skagpu_distance(self._h, float(constant), int(filt_ambig), out))
        return [(d.distance, d.mismatch_prop, d.match_count, d.mismatch_count) for d in out[:n * (n - 1) // 2]]

    def distance_sharded(self, constant_local, filt_ambig):
        """collective (library communicator): skagpu_distance on this rank's slab, counters and the slabs' constant-site counts summed
        over the ranks -> numpy record array (distance, mismatch_prop, matches, mismatches) of the whole job's pairs"""
        n = self.nsamples()
        npairs = n * (n - 1) // 2
        out = np.zeros(max(1, npairs), dtype=[("distance", "f8"), ("mismatch_prop", "f8"), ("matches", "u8"), ("mismatches", "u8")])
        _check(lib().skagpu_distance_sharded(self._h, float(constant_local), int(filt_ambig), out.ctypes.data))
        return out[:npairs]

    def distance_counters(self, filt_ambig):
        """(n_pairs, 4) uint64 partial sums {mismatches, matches, dist_int, n_both} over this array's rows"""
        n = self.nsamples()
        out = np.zeros((max(1, n * (n - 1) // 2), 4), dtype=np.uint64)
        _check(lib().skagpu_distance_counters(self._h, int(filt_ambig), out.ctypes.data))
        return out[:n * (n - 1) // 2]

    # ---- joins on the key
    def delete_samples(self, del_names):
        """MergeSkaArray::delete_samples (merge_ska_array.rs:231-271)"""
        del_names = list(del_names)
        if not del_names or len(del_names) == self.nsamples():
            raise SkaError("Invalid number of samples to remove")
        missing = [n for n in del_names if n not in self._names]
        if missing:
            raise SkaError(f"Could not find sample(s): {set(missing)}")
        idx = np.array([self._names.index(n) for n in del_names], dtype=np.uint32)
        _check(lib().skagpu_delete_samples(self._h, idx.ctypes.data, len(idx)))
        self._names = [n for n in self._names if n not in del_names]

    def weed(self, weed_sample, reverse=False):
        """RefSka::new + MergeSkaArray::weed (merge_ska_array.rs:452-487); weed_sample: a Sample holding the weed FASTA's records
        -> number of split k-mers in the weed sequences"""
        n = C.c_uint64()
        _check(lib().skagpu_weed(self._h, _inputs_array([weed_sample]), int(reverse), C.byref(n)))
        return int(n.value)

    def map(self, ref_sample, want_records=False):
        """RefSka::new + RefSka::map (ska_ref.rs:189-311, 508-533); ref_sample: a Sample holding the reference's records ->
        (middle positions uint64[M] in record-concatenated coordinates, mapped variants (M, N) uint8, n_ref_kmers[, K1 records])"""
        h = C.c_void_p()
        _check(lib().skagpu_map(self._h, _inputs_array([ref_sample]), C.byref(h)))
        try:
            v = [C.c_uint64() for _ in range(4)]
            _check(lib().skagpu_map_dims(h, *[C.byref(x) for x in v]))
            m, n_ref, n, n_pos = (x.value for x in v)
            pos = np.zeros(m, dtype=np.uint64)
            var = np.zeros((m, n), dtype=np.uint8)
            rec = np.zeros(n_pos * (2 if self.kmer_len() > 31 else 1), dtype=np.uint64) if want_records else None
            _check(lib().skagpu_map_export(h, pos.ctypes.data, var.ctypes.data, rec.ctypes.data if rec is not None else None))
        finally:
            lib().skagpu_map_free(h)
        return (pos, var, n_ref, rec) if want_records else (pos, var, n_ref)

    def checksum(self, col_lo=0, col_hi=None):
        """-> (cell checksum over columns [col_lo, col_hi), rows with a cell there, rows with a wrong count, keys out of order)"""
        out = np.zeros(4, dtype=np.uint64)
        hi = self.nsamples() if col_hi is None else col_hi
        _check(lib().skagpu_array_checksum(self._h, int(col_lo), int(hi), out.ctypes.data))
        return tuple(int(x) for x in out)

    # ---- MergeSkaArray::write_fasta (merge_ska_array.rs:499-517)
    def transpose_on_device(self):
        """device-only transpose (no D2H), for kernel timing"""
        _check(lib().skagpu_export_transposed(self._h, None))

    def transposed(self, out=None):
        d = self.dims()
        if out is not None:
            _check(lib().skagpu_export_transposed(self._h, out.ctypes.data))
            return out
        t = np.zeros((d["n_samples"], d["n_rows"]), dtype=np.uint8)
        _check(lib().skagpu_export_transposed(self._h, t.ctypes.data))
        return t

    def transposed_async(self, out):
        """queue transpose + copy to the pinned array `out`; export_wait() completes it (skagpu_export_transposed_async)"""
        _check(lib().skagpu_export_transposed_async(self._h, out.ctypes.data))

    def write_fasta(self, f):
        t = self.transposed()
        for i, name in enumerate(self._names):
            f.write(b">" + name.encode() + b"\n")
            f.write(t[i].tobytes())
            f.write(b"\n")

    # ---- Display / Debug (merge_ska_array.rs:649-698), as `ska nk [--full-info]` prints them
    def n_sample_kmers(self):
        _, var, _ = self.export()
        return [int(c) for c in (var != ord("-")).sum(axis=0)] if var.size else [0] * self.nsamples()

    def display(self, full_info=False):
        d = self.dims()
        out = (f"ska_version={self.ska_version}\nk={d['k']}\nk_bits={d['key_bits']}\nrc={'true' if d['rc'] else 'false'}\n"
               f"k-mers={d['n_rows']}\nsamples={d['n_samples']}\n"
               "sample_names=[" + ", ".join('"' + n + '"' for n in self._names) + "]\n"
               "sample_kmers=[" + ", ".join(str(c) for c in self.n_sample_kmers()) + "]\n\n")
        if full_info:
            keys, var, _ = self.export()
            h = (d["k"] - 1) // 2
            for i in range(d["n_rows"]):
                key = int(keys[i])
                up = "".join("ACTG"[(key >> (2 * h + 2 * (h - 1 - j))) & 3] for j in range(h))
                lo = "".join("ACTG"[(key >> (2 * (h - 1 - j))) & 3] for j in range(h))
                out += f"{up}\t{lo}\t" + ",".join(chr(b) for b in var[i]) + "\n"
            out += "\n"
        return out


def build_and_merge(input_files, k, rc, qual=None, threads=1, proportion_reads=None):
    """merge_ska_dict.rs:354-417 followed by MergeSkaArray::new (generic_modes.rs:279).
    input_files: [(name, file1, file2|None)] or [(name, Sample)].  `threads` is accepted for
    interface parity; the device schedules its own parallelism."""
    qual = qual or QualOpts()
    names, samples = [], []
    for item in input_files:
        if len(item) == 2 and isinstance(item[1], Sample):
            names.append(item[0])
            samples.append(item[1])
        else:
            name, f1, f2 = item
            names.append(name)
            samples.append(Sample.from_files(f1, f2))
    return build_from_samples(samples, names, k, rc, qual, proportion_reads)


def build_from_samples(samples, names, k, rc, qual=None, proportion_reads=None):
    qual = qual or QualOpts()
    h = C.c_void_p()
    qf = QUAL_FILTERS[qual.qual_filter] if isinstance(qual.qual_filter, str) else qual.qual_filter
    _check(lib().skagpu_build(_inputs_array(samples), len(samples), k, int(rc), qual.min_count, qual.min_qual, qf,
                              float(proportion_reads) if proportion_reads else 0.0, C.byref(h)))
    return MergeSkaArray(h, names)


def kmer_count_hist(sample, k, rc=True):
    """counts[c - 1] = distinct split k-mers of the read set seen c times, c = 1..1000 (coverage.rs:70-135, 156-163)"""
    hist = np.zeros(1000, dtype=np.uint64)
    _check(lib().skagpu_kmer_count_hist(_inputs_array([sample]), k, int(rc), hist.ctypes.data))
    return hist


# ---- the library's own multi-GPU transport (csrc/comm.cu): one host thread or process per GPU, no torch involved
COMM_ID_BYTES = 128


def comm_unique_id():
    """rank 0: the NCCL id every rank passes to comm_init (hand it over by any means: a shared variable, a file, a store)"""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    _check(lib().skagpu_comm_unique_id(buf))
    return bytes(buf)


def comm_init(comm_id, world, rank):
    buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id)
    _check(lib().skagpu_comm_init(buf, world, rank))


def comm_destroy():
    _check(lib().skagpu_comm_destroy())


def exchange_rows(local, sample_base, names_all):
    """collective: the array of this rank's samples -> this rank's slab of the whole job's array (skagpu_exchange_rows).
    -> (slab, {rows_sent, rows_received, bytes_sent, slab_rows})"""
    h = C.c_void_p()
    st = (C.c_uint64 * 4)()
    _check(lib().skagpu_exchange_rows(local._h, sample_base, len(names_all), C.byref(h), st))
    return MergeSkaArray(h, list(names_all), local.ska_version), dict(zip(("rows_sent", "rows_received", "bytes_sent", "slab_rows"), map(int, st)))


def allreduce_u64(values):
    v = np.ascontiguousarray(values, dtype=np.uint64).copy()
    _check(lib().skagpu_allreduce_u64(v.ctypes.data, v.size))
    return v


def plan_owners(region_rows, world):
    """the owner cut of the row exchange (host arithmetic only): cuts[world + 1] region indices"""
    c = np.ascontiguousarray(region_rows, dtype=np.uint64)
    cuts = np.zeros(world + 1, dtype=np.uint64)
    _check(lib().skagpu_plan_owners(c.ctypes.data, len(c), world, cuts.ctypes.data))
    return cuts.astype(np.int64)


def merge_arrays(a, b):
    """MergeSkaDict::extend (merge_ska_dict.rs:160-193): samples of b appended to those of a, rows joined on the key"""
    h = C.c_void_p()
    _check(lib().skagpu_merge_arrays(a._h, b._h, C.byref(h)))
    return MergeSkaArray(h, a.names() + b.names(), a.ska_version)


def build_from_packed(samples, names, k, rc):
    """skagpu_build_packed: host-packed assemblies (PackedSample) -> array; K1 runs behind the (4x smaller) upload"""
    h = C.c_void_p()
    _check(lib().skagpu_build_packed(_packed_inputs_array(samples), len(samples), k, int(rc), C.byref(h)))
    return MergeSkaArray(h, names)


def build_from_batch(batch, names, k, rc, qual=None):
    qual = qual or QualOpts()
    h = C.c_void_p()
    qf = QUAL_FILTERS[qual.qual_filter] if isinstance(qual.qual_filter, str) else qual.qual_filter
    _check(lib().skagpu_build_from_batch(batch._h, k, int(rc), qual.min_count, qual.min_qual, qf, C.byref(h)))
    return MergeSkaArray(h, names)


# ------------------------------------------------------------------------------------ generic_modes.rs
def apply_filters(ska_array, min_freq, filter_ambig_as_missing, filter_type, ambig_mask, ignore_const_gaps):
    """generic_modes.rs:112-131"""
    threshold = int(math.ceil(ska_array.nsamples() * min_freq))
    return ska_array.filter(threshold, filter_ambig_as_missing, filter_type, ambig_mask, ignore_const_gaps, False)


def align(ska_array, out, filter_type="no-const", mask_ambig=False, ignore_const_gaps=False, min_freq=0.9,
          filter_ambig_as_missing=False):
    """generic_modes.rs:22-50 (defaults: cli.rs:15,227-244); `out` is a binary file object"""
    apply_filters(ska_array, min_freq, filter_ambig_as_missing, filter_type, mask_ambig, ignore_const_gaps)
    ska_array.write_fasta(out)


def distance(ska_array, out, min_freq=0.0, filt_ambig=True):
    """generic_modes.rs:136-189 (defaults: cli.rs:287-298, lib.rs:715)"""
    if min_freq * ska_array.nsamples() >= 1.0:
        apply_filters(ska_array, min_freq, False, "no-filter", False, False)
    constant = apply_filters(ska_array, 0.0, False, "no-const", False, False)
    dists = ska_array.distance(float(constant), filt_ambig)
    names = ska_array.names()
    out.write(b"Sample1\tSample2\tDistance\tMismatches (proportion)\tMatch count\tMismatch count\n")
    it = iter(dists)
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            d, p, m, mm = next(it)
            out.write(f"{names[i]}\t{names[j]}\t{d:.2f}\t{p:.5f}\t{m}\t{mm}\n".encode())
