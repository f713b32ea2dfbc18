#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on its config: genomes/s for `ska build` + `ska align`
(k=31, both strands) over synthetic 5 Mbp bacterial genomes (star model, 0.1 % SNPs).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                         # the reference's CPU algorithm (oracle port)

A "step" is one pass of the hot path over one batch: build (K1 extract -> partition -> bucket merge
into the n_kmers x n_samples array) + align (default filter: min_freq 0.9, no-const; transpose for
the FASTA writer).  `value` times K steps with the sequence bytes already resident in HBM (CUDA
events on the library's own stream); `e2e` times the same K steps through the C ABI with pinned
HOST buffers in and out (H2D of every genome and D2H of the alignment inside the timed region).
Inputs (640 MB per 128 genomes) are larger than the 126 MB L2, so no L2 flush is needed between steps.
"""
import argparse
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
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_summary_r01.json,
    same workload: 128 x 5 Mbp), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_summary_r01.json")
    try:
        with open(p) as fh:
            d = json.load(fh)[kernel]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return sum(float(d[m][0]) * unit[d[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
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


def cpu_reference_step(genomes, k, threads):
    """One build+align step of the reference's algorithm on the host (oracle port)."""
    from oracle import oracle as O
    samples = [O.SampleData([(g.tobytes(), None)]) for g in genomes]
    arr = O.OracleArray.build(samples, [f"s{i}" for i in range(len(genomes))], k=k, rc=True, threads=threads)
    out = arr.align()
    return len(out)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm for this path (C++ restatement in oracle/,
    kind "port": the Rust reference cannot be built here), all host threads its own rule allows."""
    if rank != 0:
        return
    from ska_rust_b200 import synth
    cores = os.cpu_count() or 1
    n_warm = args.warmup if args.ref_warmup < 0 else args.ref_warmup
    # bounded sample: at ~2 genomes/s the whole run (warm-up + steps) stays within a few minutes whatever --steps is
    n = max(16, min(args.ref_genomes, 480 // max(1, args.steps + n_warm)))
    threads = host_threads_rule(cores, n)
    population = args.population if args.population != "auto" else ("star" if args.gpus == 1 else "tree")
    genomes = synth.genomes(population, n, args.length, args.genomes * args.gpus, seed=args.seed)
    for _ in range(n_warm):
        cpu_reference_step(genomes, args.k, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(genomes, args.k, threads)
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt
    sample = (f"{n} of {args.genomes} genomes x {args.length} bp per step (build+align), {threads} threads = the reference's "
              f"rule 2^floor(log2(min(cores={cores}, 1+n/10))) for n={n}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"ska build + align, {args.genomes} x {args.length} bp synthetic bacterial FASTA, k={args.k}, CPU",
                   "k": args.k, "rc": True, "genomes_per_step": n, "genome_bp": args.length, "population": population},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
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
    population = args.population if args.population != "auto" else ("star" if world == 1 else "tree")
    genomes = synth.genomes(population, G, L, G * world, seed=args.seed, first=rank * G, out=pinned.numpy())
    off = np.array([0, L], dtype=np.uint64)
    samples = [S.Sample(g, off) for g in genomes]      # views into the pinned block, no copy
    for s, g in zip(samples, genomes):
        assert s.seq.ctypes.data == g.ctypes.data
    names = [f"s{rank * G + i}" for i in range(G)]
    log(f"generated {G} x {L} bp genomes in pinned memory" + (f" (rank bound to CPUs {numa})" if numa else ""))
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
    strategy = os.environ.get("BENCH_STRATEGY", "rows")       # distributed.py: "rows" (default) or "records"

    def build_device(batch):
        if world == 1:
            return S.api.build_from_batch(batch, names, k, True)
        return D.build_and_merge_sharded(backend, batch, names, rank * G, k, True, strategy=strategy,
                                         timings=shard_t if os.environ.get("BENCH_SHARD_TIMINGS") else None)

    # ---------------- device-resident arm
    batch = S.SeqBatch(samples)
    log("batch uploaded")
    for _ in range(args.warmup):
        arr = build_device(batch)
        align_device(arr)
        del arr
    log(f"device warm-up done; stages {S.api.last_timings()}")
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
    # ---------------- K1 alone (the kernel the roofline is quoted on), CUDA events inside the library
    _rec, k1_ms = batch.k1_extract(k, rc=True, iters=10, want_records=False)
    del batch

    # ---------------- end-to-end arm: pinned host buffers in, alignment bytes out
    out_host = torch.empty((G * world, max(1, rows_aligned)), dtype=torch.uint8, pin_memory=True).numpy()

    def e2e_step():
        if world == 1:
            arr = S.build_from_samples(samples, names, k, True)
        else:
            arr = build_device(samples)                      # H2D of this rank's genomes inside the local build, then the exchange
        arr.filter(threshold, False, "no-const", False, False, False)
        r = arr.dims()["n_rows"]
        assert r == rows_aligned
        arr.transposed(out=out_host)
        return r

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    S.api.timer_start()
    walls = []
    for _ in range(args.steps):
        tw = time.perf_counter()
        e2e_step()
        walls.append(round((time.perf_counter() - tw) * 1e3, 1))
    ms_e2e = S.api.timer_stop()
    log(f"e2e per-step wall ms: {walls}; last build stages {[(s, round(ms, 2)) for s, ms, _ in S.api.last_timings()]}")
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    log(f"e2e arm: {ms_e2e / args.steps:.2f} ms/step")
    e2e_value = args.steps * G * world / (ms_e2e * 1e-3)
    h2d = n_bases + 8 * (2 * G + 1)                    # sequence bytes + record starts + sample offsets
    d2h = G * world * rows_aligned + 4 * G + 24        # this rank's alignment slab + per-sample valid counts + scalars

    # ---------------- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_genomes
        threads = host_threads_rule(cores, n_cpu)
        t0 = time.perf_counter()
        cpu_reference_step([g for g in genomes[:n_cpu]], k, threads)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n_cpu / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"first {n_cpu} of the {G} genomes, one build+align pass, {threads} threads (reference's "
                                  f"own thread rule for n={n_cpu}; host has {cores} cores)"}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = K1_ALGO_BYTES_PER_BASE * n_bases / (k1_ms * 1e-3) / 1e9
        step_ms = ms_dev / args.steps
        st = {s_: ms for s_, ms, _ in stages}
        # the merge kernel: bucket_merge_kernel (stage "merge": one pass over the records) or, when the fused pass does not
        # apply, bucket_emit_smem_kernel (stage "emit").  Algorithmic bytes: 9 B per record read +
        # (8 + N + 8) B per row written
        n_cols = G * world if strategy != "rows" or world == 1 else G
        rows_local = local_rows[0] if local_rows[0] else rows_built
        emit_bytes = 9.0 * n_bases + rows_local * (16.0 + n_cols)
        merge_stage = "merge" if "merge" in st else "emit"
        merge_kernel = "bucket_merge_kernel" if merge_stage == "merge" else "bucket_emit_smem_kernel"
        emit_ms = st.get(merge_stage, 0.0) or 1e-9
        emit_achieved = emit_bytes / (emit_ms * 1e-3) / 1e9
        part1_bytes = (16.0 if os.environ.get("SKA_K1_HIST", "0") not in ("", "0") else 24.0) * n_bases
        full_size = (G == 128 and L == 5_000_000 and world == 1 and population == "star")   # what profiles/ was captured on
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"ska build + align, {G} x {L} bp synthetic bacterial FASTA per GPU, k={k}, both strands "
                                   "(BASELINE configs[1])", "k": k, "rc": True, "genomes_per_gpu": G, "genome_bp": L,
                       "population": ("star model, 0.1% SNPs/sample" if population == "star" else
                                      f"tree model (SURVEY 8d, configs[2]): {G * world} leaves, every leaf 0.1% from the root"),
                       "align": "min_freq 0.9, filter no-const",
                       "l2": "inputs (640 MB/128 genomes) larger than L2; no flush",
                       "parallelism": "1 GPU" if world == 1 else
                       (f"{world} ranks: every rank builds its {G} samples, one NCCL exchange of partial rows by key range, "
                        f"owners join on the key; rows sharded for filter/align ({G * world} columns per slab)" if strategy == "rows" else
                        f"{world} ranks: samples sharded for extraction, one NCCL all-to-all of records by key range, "
                        f"rows sharded for merge/filter ({G * world} columns per slab)"),
                       "rows_built": rows_built, "rows_aligned": rows_aligned},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            # the merge kernel (K3); the partition's scatter kernels have a larger share of the step, see stages_ms / roofline_scatter
            "roofline": {"kernel": merge_kernel, "bound": "hbm", "achieved": emit_achieved, "peak": peak, "unit": "GB/s",
                         "frac": emit_achieved / peak, "traffic": ncu_traffic(merge_kernel.replace("_kernel", "")) if full_size else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": emit_bytes, "launch_ms": emit_ms,
                         "share_of_step": emit_ms / step_ms,
                         "note": "instruction issue / shared-memory atomics bound it, not DRAM (profiles/README.md); launch_ms covers "
                                 "the kernel's two launches (bucket size classes), traffic is ncu's for the single-launch form"},
            # the kernel the north star quotes a target for (>= 60% of HBM peak): split-k-mer extraction
            "roofline_k1": {"kernel": "k1_extract_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": ncu_traffic("k1_extract") if full_size else None,
                            "peak_source": peak_src, "algorithmic_bytes_per_launch": K1_ALGO_BYTES_PER_BASE * n_bases,
                            "launch_ms": k1_ms, "share_of_step": st.get("k1_extract", 0.0) / step_ms,
                            "frac_of_nominal_8TBs": achieved / 8000.0},
            # first partition pass (stage "partition1") = part_hist_kernel + scan + part_scatter_kernel<1>: 8 B read by the
            # histogram, 8 B read + 8 B written by the scatter, per record (with SKA_K1_HIST=1 K1 counts the histogram: 16 B)
            "roofline_scatter": {"kernel": "part_hist_kernel + part_scatter_kernel<1>", "bound": "hbm", "unit": "GB/s", "peak": peak,
                                 "achieved": part1_bytes / ((st.get("partition1", 0.0) or 1e-9) * 1e-3) / 1e9,
                                 "frac": part1_bytes / ((st.get("partition1", 0.0) or 1e-9) * 1e-3) / 1e9 / peak,
                                 "algorithmic_bytes_per_launch": part1_bytes, "launch_ms": st.get("partition1", 0.0),
                                 "traffic": None, "share_of_step": st.get("partition1", 0.0) / step_ms},
            "stages_ms": {s: round(ms, 3) for s, ms, _ in stages},
            "clocks": clocks.summary(),
        }
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
    ap.add_argument("--genomes", type=int, default=128, help="genomes per GPU (BASELINE configs[1]: 128)")
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--seed", type=int, default=0x5CA20002)
    ap.add_argument("--cpu-genomes", type=int, default=40, help="bounded CPU-baseline sample (genomes): 4 threads by the reference's rule")
    ap.add_argument("--ref-genomes", type=int, default=80, help="genomes per step of the --impl reference arm: the smallest job for which "
                    "the reference's rule (merge_ska_dict.rs:381-385) uses 8 threads, as it does for the full 128")
    ap.add_argument("--ref-warmup", type=int, default=1, help="warm-up steps of the reference arm (-1: use --warmup)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--population", default="auto", choices=["auto", "star", "tree"],
                    help="SURVEY 8d generator: star (configs[1], the 1-GPU default) or tree (recommended there for the "
                         "1024-genome configs[2]; the default when several GPUs share one population)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
