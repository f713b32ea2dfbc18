"""Row-sharded build / align / distance over several GPUs: one process per GPU, torch.distributed for
the transport (NCCL on GPUs; gloo in the CPU tests).

The reference is single-node shared memory (README.md:124); its build_and_merge
(src/merge_ska_dict.rs:354-417) has exactly one point where data from different samples must meet --
rows of the merged dictionary are keyed by split k-mer.  Here:

Two exchange strategies give the same slabs (tests compare both with the single-GPU array):

"rows" (default; SURVEY 8e "local pre-merge"): every rank builds the array of ITS samples (all its rows, its own
columns) exactly as a single GPU would; the ranks' cumulative key histograms (2^12 key ranges) are all-gathered and
cut into `world` contiguous owner ranges of equal row count; ONE exchange (three all-to-alls: keys, variant rows,
counts) moves every partial row to its owner; the owner joins the runs on the key (csrc/rowmerge.cu), cells of
samples without the k-mer stay '-'.  Related genomes share most split k-mers, so a rank ships one (8 + columns + 4)-byte
row per DISTINCT key instead of one 8-byte record per sample and key: 3-10x fewer bytes, and the merge of wide rows
never happens on records.

"records": the exchange sits between the two partition passes of one big merge:
  1. every rank extracts the records of ITS samples (K1) and runs the FIRST of the two partition passes:
     records grouped by the top `region_bits` bits of the key ("regions").  The ranks' region offsets are
     all-gathered (a few KB); summed they are cut into `world` contiguous key ranges of equal record count
     -> owner of every region; every rank derives all send/receive counts and run offsets from the same table;
  2. ONE all-to-all moves each record range to its owner (one packed u64 per record, sample id inside);
  3. the owner runs the SECOND partition pass over the runs it received (this also concatenates them) and
     merges the buckets into its slab of rows with ALL samples' columns.  Filters are row-local; distances
     are per-pair integer sums over rows, so slabs add up with one all-reduce.

Slabs are ascending key ranges in rank order: concatenating them gives the canonical array, which is
how results are compared with the single-GPU path (identical, not merely equivalent).

`backend` abstracts the three device steps so the same orchestration runs on CPU in the tests
(tests/test_distributed_gloo.py injects a CPU stand-in); the product backend is GpuBackend.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import api


class _CudaView:
    """zero-copy torch view of library-owned device memory"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class GpuBackend:
    """The product backend: libskagpu through its C ABI."""
    device = "cuda"

    def extract(self, samples, k, rc, sample_base, n_samples_total, prefix_bits):
        """-> (packed records int64 tensor grouped by region, region offsets uint64[2^region_bits + 1])"""
        import ctypes as C
        batch = samples if isinstance(samples, api.SeqBatch) else api.SeqBatch(samples)
        h = C.c_void_p()
        api._check(api.lib().skagpu_extract_records(batch._h, k, int(rc), sample_base, n_samples_total, prefix_bits, C.byref(h)))
        n = C.c_uint64()
        rec_p = C.c_void_p()
        rb = C.c_int()
        api._check(api.lib().skagpu_records_info(h, C.byref(n), C.byref(rec_p), C.byref(rb), None))
        off = np.zeros((1 << rb.value) + 1, dtype=np.uint64)
        api._check(api.lib().skagpu_records_info(h, None, None, None, off.ctypes.data))
        api.lib().skagpu_records_free(h)        # frees only the handle; buffers live in the arena until the next build
        n = n.value
        rec = (torch.as_tensor(_CudaView(rec_p.value, n, "<i8"), device="cuda") if n
               else torch.empty(0, dtype=torch.int64, device="cuda"))
        return rec, off

    def merge(self, rec, run_offsets, region_lo, region_hi, prefix_bits, n_samples_total, k, rc, names):
        """second partition pass over the received runs + merge; `rec` is consumed (scratch)"""
        import ctypes as C
        h = C.c_void_p()
        ro = np.ascontiguousarray(run_offsets, dtype=np.uint64)
        api._check(api.lib().skagpu_merge_records_dev(rec.data_ptr(), ro.shape[0], ro.ctypes.data, region_lo, region_hi,
                                                      prefix_bits, n_samples_total, k, int(rc), C.byref(h)))
        return api.MergeSkaArray(h, names)

    # ---- "rows" strategy
    region_bits_max = 12          # 4096 key ranges to cut owners from: fine enough to balance 8-16 ranks, small enough to plan in microseconds

    def build_local(self, samples, names_local, k, rc, qual=None):
        """the rank's own samples -> its array; FASTQ samples take `qual` (api.QualOpts: min_count, min_qual, qual_filter)"""
        if isinstance(samples, api.SeqBatch):                    # already resident in HBM
            return api.build_from_batch(samples, names_local, k, rc, qual)
        if samples and isinstance(samples[0], api.PackedSample):    # host-packed assemblies: 0.375 B/base across PCIe
            return api.build_from_packed(samples, names_local, k, rc)
        return api.build_from_samples(samples, names_local, k, rc, qual)   # host buffers: K1 overlaps the upload (skagpu_build)

    def key_cumulative(self, arr, region_bits):
        cum = np.zeros((1 << region_bits) + 1, dtype=np.uint64)
        api._check(api.lib().skagpu_key_cumulative(arr._h, region_bits, cum.ctypes.data))
        return cum.astype(np.int64)

    def array_tensors(self, arr):
        """zero-copy device views (keys int64[R], variants uint8[R * cols], counts int32[R]) of a local array"""
        import ctypes as C
        kp, vp, cp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        api._check(api.lib().skagpu_array_dev(arr._h, C.byref(kp), C.byref(vp), C.byref(cp)))
        d = arr.dims()
        r, n = d["n_rows"], d["n_samples"]
        if r == 0:
            return (torch.empty(0, dtype=torch.int64, device="cuda"), torch.empty(0, dtype=torch.uint8, device="cuda"),
                    torch.empty(0, dtype=torch.int32, device="cuda"))
        return (torch.as_tensor(_CudaView(kp.value, r, "<i8"), device="cuda"),
                torch.as_tensor(_CudaView(vp.value, r * n, "|u1"), device="cuda"),
                torch.as_tensor(_CudaView(cp.value, r, "<i4"), device="cuda"))

    def merge_partial(self, keys, var, cnt, run_rows, run_cols, run_col_base, n_samples_total, k, rc, names):
        import ctypes as C
        h = C.c_void_p()
        rr = np.ascontiguousarray(run_rows, dtype=np.uint64)
        rc_ = np.ascontiguousarray(run_cols, dtype=np.uint32)
        rb = np.ascontiguousarray(run_col_base, dtype=np.uint32)
        api._check(api.lib().skagpu_merge_partial_rows_dev(keys.data_ptr(), var.data_ptr(), cnt.data_ptr(), len(rr), rr.ctypes.data,
                                                           rc_.ctypes.data, rb.ctypes.data, n_samples_total, k, int(rc), C.byref(h)))
        return api.MergeSkaArray(h, names)

    def n_positions(self, samples):
        if isinstance(samples, api.SeqBatch):
            samples = samples._keep
        return int(sum(s.n_bases if isinstance(s, api.PackedSample) else len(s.seq) for s in samples))

    def prefix_bits(self, n_records_total, k, n_samples_total):
        p = int(api.lib().skagpu_merge_prefix_bits(int(n_records_total), k, n_samples_total))
        if p < 0:
            raise api.SkaError(api.lib().skagpu_last_error().decode())
        return p


def plan_owners(global_bucket_counts, world):
    """Cut the buckets (ascending key prefix) into `world` contiguous ranges of near-equal record count.
    -> int array of world+1 bucket indices.  Deterministic given the counts."""
    c = np.asarray(global_bucket_counts, dtype=np.uint64)
    cum = np.concatenate([[0], np.cumsum(c)])
    total = int(cum[-1])
    cuts = [0]
    for g in range(1, world):
        target = total * g // world
        b = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(max(b, cuts[-1]), len(c)))
    cuts.append(len(c))
    return np.array(cuts, dtype=np.int64)


def exchange_plan(all_off, world, rank):
    """all_off: (world, n_regions + 1) region offsets of every rank's first-pass output.
    -> (cuts, send_counts, recv_counts, run_offsets) for `rank`; identical arithmetic on every rank."""
    all_off = np.asarray(all_off, dtype=np.int64)
    counts = np.diff(all_off, axis=1)                      # (world, n_regions)
    cuts = plan_owners(counts.sum(axis=0), world)          # region index where each owner's range starts
    bounds = all_off[:, cuts]                              # (world, world + 1): record index, per source rank
    send_counts = np.diff(bounds[rank])
    recv_counts = bounds[:, rank + 1] - bounds[:, rank]
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    run_base = np.concatenate([[0], np.cumsum(recv_counts)[:-1]])
    run_offsets = all_off[:, lo:hi + 1] - bounds[:, rank][:, None] + run_base[:, None]
    return cuts, send_counts, recv_counts, run_offsets.astype(np.uint64)


def rows_exchange_plan(all_cum, world, rank):
    """all_cum: (world, n_regions + 1) cumulative row counts of every rank's local array over the key regions.
    -> (cuts, send_rows, recv_rows): region index where each owner starts, rows this rank sends to every owner,
    rows it receives from every source.  Identical arithmetic on every rank."""
    all_cum = np.asarray(all_cum, dtype=np.int64)
    cuts = plan_owners(np.diff(all_cum, axis=1).sum(axis=0), world)
    bounds = all_cum[:, cuts]                              # (world, world + 1): row index per source rank
    send_rows = np.diff(bounds[rank])
    recv_rows = bounds[:, rank + 1] - bounds[:, rank]
    return cuts, send_rows, recv_rows


def build_and_merge_sharded_rows(backend, samples_local, names_all, sample_base, k, rc, group=None, timings=None, qual=None):
    """"rows" strategy (module docstring).  -> this rank's slab.  Collective over `group`."""
    import time
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = backend.device
    t_last = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            if dev == "cuda":
                torch.cuda.synchronize()
            now = time.perf_counter()
            timings[name] = (now - t_last[0]) * 1e3
            t_last[0] = now

    n_local = samples_local.n_samples() if hasattr(samples_local, "n_samples") else len(samples_local)
    if qual is not None:
        local = backend.build_local(samples_local, names_all[sample_base:sample_base + n_local], k, rc, qual)
    else:
        local = backend.build_local(samples_local, names_all[sample_base:sample_base + n_local], k, rc)
    lap("local_build_ms")
    region_bits = min(backend.region_bits_max, 2 * (k - 1))
    cum = backend.key_cumulative(local, region_bits)
    mine = torch.from_numpy(np.concatenate([cum, np.array([sample_base, n_local], dtype=np.int64)])).to(dev)
    gathered = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    table = gathered.cpu().numpy().reshape(world, -1)
    all_cum, col_base, cols = table[:, :-2], table[:, -2], table[:, -1]
    cuts, send_rows, recv_rows = rows_exchange_plan(all_cum, world, rank)
    lap("owners_ms")
    keys, var, cnt = backend.array_tensors(local)
    first = int(all_cum[rank, cuts[0]])                   # 0: the owners tile the whole key space
    n_send = int(send_rows.sum())
    n_recv = int(recv_rows.sum())
    recv_keys = torch.empty(n_recv, dtype=torch.int64, device=dev)
    recv_cnt = torch.empty(n_recv, dtype=torch.int32, device=dev)
    recv_var = torch.empty(int((recv_rows * cols).sum()), dtype=torch.uint8, device=dev)
    rl, sl = list(map(int, recv_rows)), list(map(int, send_rows))
    dist.all_to_all_single(recv_keys, keys[first:first + n_send], rl, sl, group=group)
    dist.all_to_all_single(recv_cnt, cnt[first:first + n_send], rl, sl, group=group)
    dist.all_to_all_single(recv_var, var[first * n_local:(first + n_send) * n_local], [int(r * c) for r, c in zip(recv_rows, cols)],
                           [int(r) * n_local for r in send_rows], group=group)
    lap("all_to_all_ms")
    arr = backend.merge_partial(recv_keys, recv_var, recv_cnt, recv_rows, cols, col_base, len(names_all), k, rc, names_all)
    arr.local_rows = n_send                                # rows of the local build (what its merge kernel wrote)
    del local, keys, var, cnt                              # after the merge: it orders itself behind the transport
    lap("merge_ms")
    if timings is not None:
        timings["sent_rows"] = int(n_send - send_rows[rank])
        timings["recv_rows"] = n_recv
        timings["sent_bytes"] = int((n_send - send_rows[rank]) * (12 + n_local))
    return arr


_lib_comm = {"ready": False}


def ensure_library_communicator(group=None):
    """libskagpu's own NCCL communicator over the ranks of `group` (csrc/comm.cu): rank 0 draws the id, torch.distributed only
    carries those 128 bytes.  Once per process."""
    if _lib_comm["ready"]:
        return
    box = [api.comm_unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    api.comm_init(box[0], dist.get_world_size(group), dist.get_rank(group))
    _lib_comm["ready"] = True


def build_and_merge_sharded_rows_lib(backend, samples_local, names_all, sample_base, k, rc, group=None, timings=None, qual=None):
    """"rows" strategy with the exchange inside the library (skagpu_exchange_rows): histogram all-gather, owner cut, one
    grouped send/recv of keys + variant rows + counts, join -- no tensors, no host-side plan in Python."""
    import time
    ensure_library_communicator(group)
    t0 = time.perf_counter()
    n_local = samples_local.n_samples() if hasattr(samples_local, "n_samples") else len(samples_local)
    names_local = names_all[sample_base:sample_base + n_local]
    local = backend.build_local(samples_local, names_local, k, rc, qual) if qual is not None else backend.build_local(samples_local, names_local, k, rc)
    if timings is not None:
        torch.cuda.synchronize()
        timings["local_build_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
    arr, stats = api.exchange_rows(local, sample_base, names_all)
    arr.local_rows = local.dims()["n_rows"]
    if timings is not None:
        torch.cuda.synchronize()
        timings["exchange_and_merge_ms"] = (time.perf_counter() - t0) * 1e3
        timings.update(stats)
    return arr


def build_and_merge_sharded(backend, samples_local, names_all, sample_base, k, rc, prefix_bits=None, group=None,
                            timings=None, strategy="rows", qual=None):
    """-> this rank's slab (a MergeSkaArray with len(names_all) columns).  Collective over `group`.
    strategy "rows" (default) exchanges partial rows after a local build, "records" exchanges records between the
    two partition passes (module docstring); `prefix_bits` only concerns "records"."""
    if strategy == "rows_lib":
        return build_and_merge_sharded_rows_lib(backend, samples_local, names_all, sample_base, k, rc, group=group, timings=timings,
                                                qual=qual)
    if strategy == "rows":
        return build_and_merge_sharded_rows(backend, samples_local, names_all, sample_base, k, rc, group=group, timings=timings,
                                            qual=qual)
    if strategy != "records":
        raise ValueError(f"unknown strategy {strategy!r}")
    if qual is not None:
        raise api.SkaError("FASTQ samples (count filter) shard with strategy='rows' only")
    import time
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    key_bits = 2 * (k - 1)
    dev = backend.device
    t_last = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            if dev == "cuda":
                torch.cuda.synchronize()
            now = time.perf_counter()
            timings[name] = (now - t_last[0]) * 1e3
            t_last[0] = now

    if prefix_bits is None:                               # bucket granularity from the GLOBAL record count
        tot = torch.tensor([backend.n_positions(samples_local)], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, group=group)
        prefix_bits = backend.prefix_bits(int(tot.item()), k, len(names_all))
    prefix_bits = min(prefix_bits, key_bits)
    lap("plan_ms")
    rec, off = backend.extract(samples_local, k, rc, sample_base, len(names_all), prefix_bits)
    lap("extract_ms")
    mine = torch.from_numpy(off.astype(np.int64)).to(dev)
    gathered = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    all_off = gathered.cpu().numpy().reshape(world, -1)
    cuts, send_counts, recv_counts, run_offsets = exchange_plan(all_off, world, rank)
    n_recv = int(recv_counts.sum())
    recv_rec = torch.empty(n_recv, dtype=torch.int64, device=dev)
    lap("owners_ms")
    # records are grouped by region, so each owner's share is one contiguous range; one packed u64 per record
    dist.all_to_all_single(recv_rec, rec[:int(all_off[rank, -1])], list(map(int, recv_counts)), list(map(int, send_counts)),
                           group=group)
    lap("all_to_all_ms")
    arr = backend.merge(recv_rec, run_offsets, int(cuts[rank]), int(cuts[rank + 1]), prefix_bits, len(names_all),
                        k, rc, names_all)
    lap("merge_ms")
    if timings is not None:
        timings["sent_records"] = int(send_counts.sum() - send_counts[rank])
        timings["recv_records"] = n_recv
        timings["prefix_bits"] = prefix_bits
    return arr


def align_sharded(arr, filter_type="no-const", mask_ambig=False, ignore_const_gaps=False, min_freq=0.9,
                  filter_ambig_as_missing=False):
    """generic_modes.rs:22-50 on a slab: filters are row-local (merge_ska_array.rs:312-379).
    -> (n_samples, rows_of_this_slab) uint8; column blocks concatenate in rank order."""
    api.apply_filters(arr, min_freq, filter_ambig_as_missing, filter_type, mask_ambig, ignore_const_gaps)
    return arr.transposed()


def gather_alignment(slab_t, names, group=None, dst=0):
    """Concatenate the ranks' column blocks on `dst` and format the FASTA (merge_ska_array.rs:499-517)."""
    world = dist.get_world_size(group)
    parts = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(np.ascontiguousarray(slab_t), parts, dst=dst, group=group)
    if parts is None:
        return None
    full = np.concatenate(parts, axis=1)
    return b"".join(b">" + n.encode() + b"\n" + full[i].tobytes() + b"\n" for i, n in enumerate(names))


def distance_sharded(arr, min_freq=0.0, filt_ambig=True, group=None, as_array=False):
    """generic_modes.rs:136-189 on slabs: -> [(distance, mismatch_prop, matches, mismatches)] for all pairs,
    identical on every rank (`as_array`: the same as a numpy record array, no per-pair Python objects)."""
    n = arr.nsamples()
    dev = "cuda" if torch.cuda.is_available() and dist.get_backend(group) == "nccl" else "cpu"
    if min_freq * n >= 1.0:
        api.apply_filters(arr, min_freq, False, "no-filter", False, False)
    constant = api.apply_filters(arr, 0.0, False, "no-const", False, False)
    if _lib_comm["ready"] and hasattr(arr, "distance_sharded"):
        # the slabs came from the library's own exchange: its communicator also sums the counters (skagpu_distance_sharded)
        out = arr.distance_sharded(constant, filt_ambig)
        return out if as_array else [(float(a), float(b), int(c_), int(d_)) for a, b, c_, d_ in out.tolist()]
    cnt = arr.distance_counters(filt_ambig).astype(np.int64)
    t = torch.from_numpy(np.concatenate([cnt.reshape(-1), np.array([constant], dtype=np.int64)])).to(dev)
    dist.all_reduce(t, group=group)
    t = t.cpu().numpy()
    cnt, constant = t[:-1].reshape(-1, 4), float(t[-1])
    # vectorised: 1024 samples are 523 776 pairs, a Python loop over them costs more than the kernels
    mism, match, dint, nboth = (cnt[:, i].astype(np.float64) for i in range(4))
    d = dint if filt_ambig else nboth - dint / 36.0
    matches = constant + match                                # merge_ska_array.rs:598
    tot = matches + mism
    prop = np.divide(mism, tot, out=np.zeros_like(tot), where=tot != 0.0)
    out = np.empty(len(cnt), dtype=[("distance", "f8"), ("mismatch_prop", "f8"), ("matches", "i8"), ("mismatches", "i8")])
    out["distance"], out["mismatch_prop"], out["matches"], out["mismatches"] = d, prop, matches.astype(np.int64), mism.astype(np.int64)
    return out if as_array else out.tolist()


def format_distances(names, dists):
    lines = ["Sample1\tSample2\tDistance\tMismatches (proportion)\tMatch count\tMismatch count"]
    it = iter(dists)
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            d, p, m, mm = next(it)
            lines.append(f"{names[i]}\t{names[j]}\t{d:.2f}\t{p:.5f}\t{m}\t{mm}")
    return ("\n".join(lines) + "\n").encode()
