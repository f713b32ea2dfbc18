#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on its configs: genomes/s for `ska build` + `ska align`
(k=31, both strands) over synthetic 5 Mbp bacterial genomes, 0.1 % SNPs (tree model).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (configs[1] per GPU)
  python bench.py --impl reference [...]                         # the reference's CPU algorithm (oracle port)
  python bench.py --workload distance|viral|fastq|cli [...]      # the other BASELINE configs (profiles/ lines)

A "step" is one pass of the hot path over one batch: build (K1 extract -> partition -> bucket merge
into the n_kmers x n_samples array) + align (default filter: min_freq 0.9, no-const; transpose for
the FASTA writer).  `value` times K steps with the sequence bytes already resident in HBM (CUDA
events on the library's own stream); `e2e` times the same K steps through the C ABI with pinned
HOST buffers in and out (H2D of every genome and D2H of the alignment inside the timed region).
Inputs (640 MB per 128 genomes) are larger than the 126 MB L2, so no L2 flush is needed between steps.

Every line is self-checking (`parity`): the GPU alignment of a sample of the genomes is compared byte for
byte with the oracle's (the CPU leg that is timed anyway), the full-size array is checked for count / order
invariants on the device, and on several ranks the joined slabs are compared, column block by column block,
with the ranks' local builds through an order-independent checksum (skagpu_array_checksum).
"""
import argparse
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_T0 = time.time()


def log(msg):
    if os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


METRIC = "genomes/sec build+align (k=31)"
UNIT = "genomes/s"
K1_ALGO_BYTES_PER_BASE = 10.0     # read 1 B ASCII + write 9 B (u64 key + u8 base) per window


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full capture (profiles/ncu_summary_r*.json,
    same workload size), or None."""
    for rnd in ("r02b", "r02", "r01"):
        p = os.path.join(ROOT, "profiles", f"ncu_summary_{rnd}.json")
        for name in (kernel if isinstance(kernel, (list, tuple)) else [kernel]):
            try:
                with open(p) as fh:
                    d = json.load(fh)[name]
                unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                return sum(float(d[m][0]) * unit[d[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            except Exception:
                continue
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval=0.2):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None
        self.interval = interval

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(self.interval)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(sm)}


def host_threads_rule(threads, n):
    """merge_ska_dict.rs:381-385: 2^floor(log2(max(1, min(threads, 1 + n/10))))"""
    m = max(1, min(threads, 1 + n // 10))
    return 1 << (m.bit_length() - 1)


def population_text(population, n_total):
    return ("star model, 0.1% SNPs/sample" if population == "star" else
            f"tree model (SURVEY 8d): {n_total} leaves, every leaf 0.1% from the root")


def workload_config(args, world, population):
    """the `config` object both arms print (same keys and values: the driver compares them)"""
    G, L, k = args.genomes, args.length, args.k
    return {"workload": f"ska build + align, {G} x {L} bp synthetic bacterial FASTA per GPU, k={k}, both strands "
                        "(BASELINE configs[1])", "k": k, "rc": True, "genomes_per_gpu": G, "genome_bp": L,
            "population": population_text(population, G * world), "align": "min_freq 0.9, filter no-const",
            "l2": f"inputs ({G * L // 1000000} MB/{G} genomes) larger than L2; no flush"}


def cpu_reference_step(genomes, k, threads, names=None):
    """One build+align step of the reference's algorithm on the host (oracle port) -> the alignment (FASTA bytes)."""
    from oracle import oracle as O
    samples = [O.SampleData([(g.tobytes(), None)]) for g in genomes]
    arr = O.OracleArray.build(samples, names or [f"s{i}" for i in range(len(genomes))], k=k, rc=True, threads=threads)
    return arr.align()


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm for this path (C++ restatement in oracle/, kind "port": the
    Rust reference cannot be built here) on the SAME job as our arm's rank 0 -- all `--genomes` genomes of the same
    population, the thread count the reference's own rule gives for that many samples (merge_ska_dict.rs:381-385).  One
    such pass takes about a minute, so the run is bounded by timing ONE pass whatever --steps/--warmup say (a CPU build
    has no warm-up effect worth a second minute); the line reports that pass as ms_per_step."""
    if rank != 0:
        return
    from ska_rust_b200 import synth
    cores = os.cpu_count() or 1
    n = args.genomes
    threads = host_threads_rule(cores, n)
    population = args.population
    genomes = synth.genomes(population, n, args.length, args.genomes * world, seed=args.seed)
    cpu_reference_step(genomes[:2], args.k, 1)              # loads the library, touches the allocator
    reps = max(1, args.ref_reps)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = cpu_reference_step(genomes, args.k, threads)
    dt = (time.perf_counter() - t0) / reps
    v = n / dt
    sample = (f"all {n} genomes x {args.length} bp of rank 0's share, {reps} timed build+align pass(es) "
              f"(--steps/--warmup are not repeated: one pass is ~1 min), {threads} threads = the reference's rule "
              f"2^floor(log2(min(cores={cores}, 1+n/10))) for n={n}; alignment {len(out)} bytes")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, world, population),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "timed_passes": reps,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU: with several ranks per host the
    host<->device copies of ranks bound to the far socket otherwise cross the inter-socket link.  Best effort."""
    if os.environ.get("BENCH_NO_AFFINITY"):
        return None
    try:
        import torch
        prop = torch.cuda.get_device_properties(local_rank)
        if all(hasattr(prop, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{int(prop.pci_domain_id):04x}:{int(prop.pci_bus_id):02x}:{int(prop.pci_device_id):02x}.0"
        else:
            bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().lower()
            if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
                bdf = bdf[4:]                               # nvidia-smi prints an 8-digit domain, sysfs uses 4
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:
        pass
    return None


# stage of the library's timers -> (kernel, algorithmic bytes per launch).  n = records (= bases), R = rows built by the
# merge this rank ran, C = its columns, Rk = rows kept by the filter, N = columns of the filtered array.  SURVEY 8d.
STAGE_KERNELS = {
    "k1_extract": ("k1_extract_kernel", lambda n, R, C, Rk, N: 10.0 * n),
    "hist1": ("part_hist_kernel", lambda n, R, C, Rk, N: 8.0 * n),
    "scatter1": ("part_scatter2_kernel<1>", lambda n, R, C, Rk, N: 16.0 * n),
    "hist2": ("part_hist_bytes_kernel", lambda n, R, C, Rk, N: 1.0 * n),     # reads the digit bytes the first scatter wrote (SKA_HIST2_BYTES=0: 8 B)
    "scatter2": ("part_scatter2_kernel<2>", lambda n, R, C, Rk, N: 16.0 * n),
    "count": ("bucket_count_kernel", lambda n, R, C, Rk, N: 8.0 * n + 8.0 * R),
    "emit": ("bucket_emit_smem_kernel", lambda n, R, C, Rk, N: 9.0 * n + R * (16.0 + C)),
    "merge": ("bucket_merge_kernel", lambda n, R, C, Rk, N: 9.0 * n + R * (16.0 + C)),
    "transpose": ("transpose_words_kernel", lambda n, R, C, Rk, N: 2.0 * Rk * N),
}
# kernel -> its entries in profiles/ncu_summary_r02.json / _r01.json (the first one found is used)
NCU_NAMES = {"k1_extract_kernel": ["k1_extract<31, 0, 1, 1, 0>", "k1_extract"], "part_scatter2_kernel<1>": ["part_scatter2<1, 8, 0>", "part_scatter<1, 8>"],
             "part_scatter2_kernel<2>": ["part_scatter2<2, 8, 0>", "part_scatter<2, 8>"], "part_hist_kernel": ["part_hist"], "part_hist_bytes_kernel": ["part_hist_bytes"], "bucket_count_kernel": ["bucket_count"],
             "bucket_emit_smem_kernel": ["bucket_emit_smem"], "transpose_words_kernel": ["transpose_words"]}


def run_ours(args, rank, world, local_rank):
    import torch

    import ska_rust_b200 as S
    from ska_rust_b200 import synth

    log("torch imported")
    torch.cuda.set_device(local_rank)
    S.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    G, L, k = args.genomes, args.length, args.k
    # weak scaling: every rank owns G genomes of the same population (sample indices rank*G ...)
    pinned = torch.empty((G, L), dtype=torch.uint8, pin_memory=True)
    population = args.population
    genomes = synth.genomes(population, G, L, G * world, seed=args.seed, first=rank * G, out=pinned.numpy())
    off = np.array([0, L], dtype=np.uint64)
    samples = [S.Sample(g, off) for g in genomes]      # views into the pinned block, no copy
    for s, g in zip(samples, genomes):
        assert s.seq.ctypes.data == g.ctypes.data
    names = [f"s{rank * G + i}" for i in range(G)]
    # the packing FASTA reader's output (skagpu_pack_bases): 2-bit codes + invalid mask in pinned memory, 0.375 B per base
    pin2 = torch.empty((G, (L + 3) // 4), dtype=torch.uint8, pin_memory=True).numpy()
    pinv = torch.empty((G, (L + 7) // 8), dtype=torch.uint8, pin_memory=True).numpy()
    t_pack = time.perf_counter()
    packed = [S.PackedSample.from_sample(s, pin2[i], pinv[i]) for i, s in enumerate(samples)]
    t_pack = time.perf_counter() - t_pack
    log(f"generated {G} x {L} bp genomes ({population}) in pinned memory, packed in {t_pack * 1e3:.0f} ms on one host thread "
        f"({G * L / t_pack / 1e9:.2f} Gbase/s)" + (f" (rank bound to CPUs {numa})" if numa else ""))
    n_bases = G * L
    threshold = int(np.ceil(G * 0.9))                  # generic_modes.rs:121 with cli.rs default min_freq 0.9

    def align_device(arr):
        arr.filter(threshold, False, "no-const", False, False, False)
        arr.transpose_on_device()

    if world > 1:
        import ska_rust_b200.distributed as D
        backend = D.GpuBackend()
        names = [f"s{i}" for i in range(G * world)]          # every slab has all N = G*world columns
        threshold = int(np.ceil(G * world * 0.9))

    shard_t = {}
    local_rows = [None]                                       # rows strategy: rows of this rank's local build
    strategy = os.environ.get("BENCH_STRATEGY", "rows_lib")   # distributed.py: "rows_lib" (exchange inside libskagpu), "rows" (torch.distributed), "records"

    def build_device(batch):
        if world == 1:
            return S.api.build_from_batch(batch, names, k, True)
        return D.build_and_merge_sharded(backend, batch, names, rank * G, k, True, strategy=strategy,
                                         timings=shard_t if os.environ.get("BENCH_SHARD_TIMINGS") else None)

    # ---------------- device-resident arm
    batch = S.SeqBatch(samples if args.ascii else packed)
    log("batch uploaded")
    for _ in range(args.warmup):
        arr = build_device(batch)
        align_device(arr)
        del arr
    log(f"device warm-up done; stages {[(s_, round(ms, 3)) for s_, ms, _ in S.api.last_timings()]}")
    barrier()
    l0 = S.api.launch_count()
    # one rank: back-to-back queries (the timed region of 5 steps is only ~100 ms long); several ranks: the measured cadence
    with ClockSampler(local_rank, 0.02 if world == 1 else 0.2) as clocks:
        S.api.timer_start()
        for _ in range(args.steps):
            arr = build_device(batch)
            rows_built = arr.dims()["n_rows"]
            local_rows[0] = getattr(arr, "local_rows", None)
            align_device(arr)
            stages = S.api.last_timings()
            rows_aligned = arr.dims()["n_rows"]
            del arr
        ms_dev = S.api.timer_stop()
    barrier()
    launches = S.api.launch_count() - l0
    ms_dev = max_over_ranks(ms_dev)
    value = args.steps * G * world / (ms_dev * 1e-3)

    log(f"device arm: {ms_dev / args.steps:.2f} ms/step, rows built {rows_built}, aligned {rows_aligned}")
    if shard_t:
        log("sharded build phases of the last call (ms): " + str({k_: round(v, 1) if isinstance(v, float) else v for k_, v in shard_t.items()}))
    # ---------------- K1 alone (the kernel the north star quotes a target on), CUDA events inside the library
    _rec, k1_ms = batch.k1_extract(k, rc=True, iters=10, want_records=False)

    # ---------------- parity at full size (not timed): device-side invariants + slabs against local builds
    parity = {"checked": True}
    arr = build_device(batch)
    full_sum = arr.checksum()
    parity["full_size_rows"] = arr.dims()["n_rows"]
    parity["rows_with_wrong_count"] = full_sum[2]
    parity["keys_out_of_order"] = full_sum[3]
    ok = full_sum[2] == 0 and full_sum[3] == 0
    if world > 1:
        # MergeSkaDict::merge only moves cells: the cells of rank r's local build == columns [r*G, (r+1)*G) of all slabs
        local = S.api.build_from_batch(batch, names[rank * G:(rank + 1) * G], k, True)
        mine = local.checksum()
        del local
        blocks = np.array([arr.checksum(r * G, (r + 1) * G)[0] for r in range(world)], dtype=np.uint64)
        t = torch.from_numpy(blocks.view(np.int64).copy()).cuda()
        dist.all_reduce(t)                                   # wraps mod 2^64 like the checksum itself
        total = t.cpu().numpy().view(np.uint64)
        flags = torch.tensor([int(int(total[rank]) == mine[0]), int(ok), int(mine[2] == 0 and mine[3] == 0)], dtype=torch.int64, device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        fl = flags.cpu().numpy()
        parity["slabs_vs_local_builds"] = ("cell checksums of every rank's local build equal its column block summed over all slabs"
                                           if fl[0] else "MISMATCH")
        ok = bool(fl.min())
    del arr
    del batch
    if not ok:
        raise SystemExit(f"parity check FAILED: {parity}")

    # ---------------- end-to-end arm: pinned host buffers in, alignment bytes out
    e2e = None
    e2e_serial = None
    if not args.no_e2e:
        out_host = [torch.empty((G * world, max(1, rows_aligned)), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]

        prefetched = [None]

        def e2e_step(i, host_samples, pipelined, last=False):
            if pipelined and isinstance(host_samples[0], S.PackedSample):
                # this step's batch was queued for upload during the previous step (the first one here); the NEXT step's
                # upload is queued now, before this step's kernels, so that it crosses PCIe underneath them
                cur = prefetched[0] or S.SeqBatch(host_samples, asynchronous=True)
                prefetched[0] = None if last else S.SeqBatch(host_samples, asynchronous=True)
                arr = build_device(cur)
                del cur
            elif world == 1:
                arr = S.build_from_samples(host_samples, names, k, True)
            else:
                arr = build_device(host_samples)                 # H2D of this rank's genomes inside the local build, then the exchange
            arr.filter(threshold, False, "no-const", False, False, False)
            r = arr.dims()["n_rows"]
            assert r == rows_aligned
            if pipelined:
                arr.transposed_async(out_host[i % 2])            # this copy overlaps the next step's upload and kernels
            else:
                arr.transposed(out=out_host[i % 2])
            return r

        def e2e_run(host_samples, pipelined, steps):
            n_warm = max(args.warmup, 5) if pipelined else 2     # the pipeline's buffers (two batches, pending exports) reach their steady state
            for i in range(n_warm):
                e2e_step(i, host_samples, pipelined, last=(i + 1 == n_warm))
            S.api.export_wait()
            barrier()
            S.api.timer_start()
            walls = []
            for i in range(steps):
                tw = time.perf_counter()
                e2e_step(i, host_samples, pipelined, last=(i + 1 == steps))    # exactly `steps` uploads and downloads in the timed region
                walls.append(round((time.perf_counter() - tw) * 1e3, 1))
            S.api.export_wait()                                  # every alignment is in host memory before the clock stops
            ms = S.api.timer_stop()
            log(f"e2e ({'packed' if isinstance(host_samples[0], S.PackedSample) else 'ascii'}, {'pipelined' if pipelined else 'serial'}) "
                f"per-step wall ms: {walls}; last build stages {[(s, round(m, 2)) for s, m, _ in S.api.last_timings()]}")
            barrier()
            return max_over_ranks(ms) / steps

        host_in = samples if args.ascii else packed
        ms_e2e = e2e_run(host_in, True, args.steps)
        log(f"e2e arm: {ms_e2e:.2f} ms/step")
        in_bytes = n_bases if args.ascii else sum(len(p.packed2) + len(p.invalid) for p in packed)
        h2d = in_bytes + 8 * (3 * G + 2)                   # sequence bytes + record boundaries + sample offsets
        d2h = G * world * rows_aligned + 4 * G + 24        # this rank's alignment slab + per-sample valid counts + scalars
        e2e = {"value": G * world / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
               "input": ("ASCII bases in pinned memory" if args.ascii else
                         "host-packed bases in pinned memory (2-bit codes + invalid mask, 0.375 B/base: the packing FASTA reader's "
                         "output, skagpu_pack_bases; packing is parse work and not in the timed region)"),
               "pipeline": "double-buffered: the upload of step i+1 (skagpu_seqbatch_upload_packed_async) and the device->host copy "
                           "of step i-1 (skagpu_export_transposed_async) cross PCIe under the kernels of step i; every step's "
                           "upload and download lie inside the timed region"}
        if world == 1 and not args.ascii:
            # round 1's form for comparison: ASCII in, every step strictly serial (upload -> kernels -> download)
            ms_ser = e2e_run(samples, False, min(args.steps, 5))
            e2e_serial = {"value": G / (ms_ser * 1e-3), "unit": UNIT, "ms_per_step": ms_ser, "h2d_bytes_per_step": n_bases + 8 * (2 * G + 1),
                          "d2h_bytes_per_step": d2h, "input": "ASCII bases in pinned memory, no overlap between steps (round 1's e2e)"}

    # ---------------- CPU baseline beside it + oracle parity on the same sample (rank 0)
    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_genomes if world == 1 else min(args.cpu_genomes, 8)
        threads = host_threads_rule(cores, n_cpu)
        sub_names = [f"s{i}" for i in range(n_cpu)]
        t0 = time.perf_counter()
        ref_aln = cpu_reference_step([g for g in genomes[:n_cpu]], k, threads, sub_names)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n_cpu / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"first {n_cpu} of the {G} genomes, one build+align pass, {threads} threads (reference's "
                                  f"own thread rule for n={n_cpu}; host has {cores} cores)"}
        garr = S.build_from_samples(samples[:n_cpu], sub_names, k, True)
        buf = io.BytesIO()
        S.api.align(garr, buf)
        same = buf.getvalue() == ref_aln
        parity["oracle_sample"] = (f"GPU build+align of the same {n_cpu} genomes: {len(ref_aln)} alignment bytes "
                                   + ("identical to the oracle's" if same else "DIFFER from the oracle's"))
        del garr
        if not same:
            raise SystemExit(f"parity check FAILED: {parity}")
    elif rank == 0:
        parity["oracle_sample"] = "skipped (--no-cpu-baseline)"

    if rank == 0:
        peak, peak_src = measured_peak()
        step_ms = ms_dev / args.steps
        st = {s_: ms for s_, ms, _ in stages}
        n_cols = G * world if not strategy.startswith("rows") or world == 1 else G
        rows_local = local_rows[0] if local_rows[0] else rows_built
        full_size = (G == 128 and L == 5_000_000 and world == 1)   # what profiles/ was captured on

        def roof(stage):
            kern, fn = STAGE_KERNELS[stage]
            ms = st.get(stage, 0.0) or 1e-9
            by = fn(float(n_bases), float(rows_local), float(n_cols), float(rows_aligned), float(G * world))
            return {"kernel": kern, "stage": stage, "bound": "hbm", "achieved": by / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": by / (ms * 1e-3) / 1e9 / peak, "traffic": ncu_traffic(NCU_NAMES.get(kern, [])) if full_size else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": by, "launch_ms": ms, "share_of_step": ms / step_ms}

        STAGE_KERNELS["k1_extract"] = ("k1_extract_kernel", lambda n, R, C, Rk, N: (10.0 if args.ascii else 9.375) * n)
        # the dominant kernel = the single-kernel stage with the largest share of the step
        dom = max((s_ for s_ in st if s_ in STAGE_KERNELS), key=lambda s_: st[s_])
        k1 = roof("k1_extract")
        k1_bytes = (K1_ALGO_BYTES_PER_BASE if args.ascii else 9.375) * n_bases     # 0.375 B read per base when host-packed
        k1.update({"achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "launch_ms": k1_ms, "algorithmic_bytes_per_launch": k1_bytes,
                   "note": "the extraction kernel alone, 10 launches timed back to back inside the library (skagpu_k1_extract); "
                           "inside a build the same kernel also fills the first partition histogram (stage k1_extract)"})
        k1["frac"] = k1["achieved"] / peak
        k1["frac_of_nominal_8TBs"] = k1["achieved"] / 8000.0
        # whole step against the roofline: SURVEY 8d's algorithmic bytes of K1 + K2/K3 + K4 (the partition is overhead)
        step_bytes = 10.0 * n_bases + 9.0 * n_bases + rows_local * (16.0 + n_cols) + rows_built * 4.0 + 2.0 * rows_aligned * (G * world) * 2
        cfg = workload_config(args, world, population)
        cfg.update({"parallelism": "1 GPU" if world == 1 else
                    (f"{world} ranks: every rank builds its {G} samples, one NCCL exchange of partial rows by key range, "
                     f"owners join on the key; rows sharded for filter/align ({G * world} columns per slab)"
                     + ("; the exchange runs inside libskagpu (skagpu_exchange_rows)" if strategy == "rows_lib" else "") if strategy.startswith("rows") else
                     f"{world} ranks: samples sharded for extraction, one NCCL all-to-all of records by key range, "
                     f"rows sharded for merge/filter ({G * world} columns per slab)"),
                    "rows_built": rows_built, "rows_aligned": rows_aligned})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": cfg,
            "gpu_launches": launches,
            "roofline": roof(dom),
            "roofline_k1": k1,
            "roofline_step": {"bound": "hbm", "algorithmic_bytes": step_bytes, "achieved": step_bytes / (step_ms * 1e-3) / 1e9,
                              "peak": peak, "unit": "GB/s", "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak},
            "roofline_stages": {s_: round(roof(s_)["frac"], 3) for s_ in st if s_ in STAGE_KERNELS and st[s_] >= 0.05},   # hist1 is a scan only when K1 counted
            "stages_ms": {s_: round(ms, 3) for s_, ms, _ in stages},
            "clocks": clocks.summary(),
            "parity_checked": True, "parity": parity,
        }
        if e2e:
            line["e2e"] = e2e
        if e2e_serial:
            line["e2e_ascii_serial"] = e2e_serial
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    import faulthandler
    faulthandler.enable()
    if os.environ.get("BENCH_DUMP_AFTER"):
        faulthandler.dump_traceback_later(int(os.environ["BENCH_DUMP_AFTER"]), repeat=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="align", choices=["align", "distance", "viral", "fastq", "cli"],
                    help="align = BASELINE configs[1] (the driver's line); the others are configs[2..4] and the file-to-file CLI run")
    ap.add_argument("--genomes", type=int, default=128, help="genomes per GPU (BASELINE configs[1]: 128)")
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--seed", type=int, default=0x5CA20002)
    ap.add_argument("--cpu-genomes", type=int, default=40, help="bounded CPU-baseline sample (genomes) beside our arm: 4 threads by the reference's rule")
    ap.add_argument("--ref-reps", type=int, default=1, help="timed passes of the --impl reference arm over the full job (~1 min each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (kernel A/B runs)")
    ap.add_argument("--ascii", action="store_true", help="feed ASCII bases (1 B/base) instead of host-packed bases (0.375 B/base)")
    ap.add_argument("--fastq-samples", type=int, default=64, help="--workload fastq: samples (BASELINE configs[3]: 64)")
    ap.add_argument("--fastq-pairs", type=int, default=2_000_000, help="--workload fastq: read pairs per sample (configs[3]: 2 M)")
    ap.add_argument("--fastq-parity-samples", type=int, default=4, help="--workload fastq: whole samples checked against the CPU port")
    ap.add_argument("--fastq-ref-samples", type=int, default=4, help="--workload fastq --impl reference: whole samples timed on the host")
    ap.add_argument("--population", default="tree", choices=["star", "tree"],
                    help="SURVEY 8d generator.  tree (default at every N, so that the 1 -> 8 GPU curve is like for like): the "
                         "samples are leaves of a balanced tree over all N x 128 genomes, every leaf 0.1%% from the root; "
                         "star: every sample 0.1%% from the ancestor independently (round 1's 1-GPU population)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload != "align":
        import importlib
        mod = importlib.import_module("profiles.workloads")
        return mod.run(args, rank, world, local_rank)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
