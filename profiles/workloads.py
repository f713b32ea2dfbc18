"""The other BASELINE.json configs, run through bench.py (`python bench.py --workload <name> [--impl reference]`):

  distance  configs[2]: ska build + distance, 128 x 5 Mbp genomes per GPU (1024 x 5 Mbp on 8 GPUs), k=31, tree model
  viral     configs[4]: ska build + align, 256 x 30 kbp genomes, k=15 single strand
  fastq     configs[3]: ska build from FASTQ, paired 150 bp reads, min-count 5 (size by --fastq-samples / --fastq-pairs)
  cli       file to file: `ska build -o x.skf` + `ska align x.skf` on FASTA files in tmpfs (first byte read to last byte written)

Same JSON line as bench.py's default workload (metric / value / e2e / roofline / cpu_baseline / parity); these lines are
committed under profiles/ -- the driver's own line is the default workload (configs[1]).
"""
import io
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POPC_PEAK = 4.5e12       # popc32 per second on one B200, measured (profiles/microbench/popc_mb.cu: 15.4 lanes/clk/SM at 1.965 GHz)


def _bench():
    import bench
    return bench


def _emit(line, rank):
    if rank == 0:
        print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- distance (configs[2])
def run_distance(args, rank, world, local_rank):
    B = _bench()
    metric, unit = "genomes/sec build+distance (k=31)", "genomes/s"
    G, L, k = args.genomes, args.length, args.k
    from ska_rust_b200 import synth
    cfg = {"workload": f"ska build + distance, {G} x {L} bp synthetic bacterial FASTA per GPU, k={k}, both strands (BASELINE configs[2]: "
                       f"1024 x 5 Mbp on 8 GPUs)", "k": k, "rc": True, "genomes_per_gpu": G, "genome_bp": L,
           "population": B.population_text(args.population, G * world), "distance": "min_freq 0, ambiguous bases ignored (cli.rs:287-298)"}
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        threads = B.host_threads_rule(cores, G)
        genomes = synth.genomes(args.population, G, L, G * world, seed=args.seed)
        samples = [O.SampleData([(g.tobytes(), None)]) for g in genomes]
        names = [f"s{i}" for i in range(G)]
        t0 = time.perf_counter()
        arr = O.OracleArray.build(samples, names, k=k, rc=True, threads=threads)
        out = arr.distance(threads=cores)                       # ska distance: rayon default pool = all cores (lib.rs:714)
        dt = time.perf_counter() - t0
        v = G / dt
        sample = (f"all {G} genomes of rank 0's share, one build (threads={threads}, the reference's rule) + distance (all {cores} cores) pass; "
                  f"{len(out)} bytes of TSV")
        _emit({"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
               "config": cfg, "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
               "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}, rank)
        return

    import torch
    import ska_rust_b200 as S
    torch.cuda.set_device(local_rank)
    S.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        import ska_rust_b200.distributed as D
        backend = D.GpuBackend()

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

    pinned = torch.empty((G, L), dtype=torch.uint8, pin_memory=True)
    genomes = synth.genomes(args.population, G, L, G * world, seed=args.seed, first=rank * G, out=pinned.numpy())
    off = np.array([0, L], dtype=np.uint64)
    samples = [S.Sample(g, off) for g in genomes]
    pin2 = torch.empty((G, (L + 3) // 4), dtype=torch.uint8, pin_memory=True).numpy()
    pinv = torch.empty((G, (L + 7) // 8), dtype=torch.uint8, pin_memory=True).numpy()
    packed = [S.PackedSample.from_sample(s, pin2[i], pinv[i]) for i, s in enumerate(samples)]
    names_all = [f"s{i}" for i in range(G * world)]
    B.log(f"generated and packed {G} x {L} bp genomes ({args.population})")

    info = {}

    def step(src):
        """build + distance; src = resident SeqBatch or host samples.  -> list of (distance, prop, matches, mismatches)"""
        if world == 1:
            arr = (S.api.build_from_batch(src, names_all, k, True) if isinstance(src, S.SeqBatch)
                   else S.api.build_from_packed(src, names_all, k, True))
            info["rows_built"] = arr.dims()["n_rows"]
            constant = S.api.apply_filters(arr, 0.0, False, "no-const", False, False)
            info["rows_compared"] = arr.dims()["n_rows"]
            out = arr.distance(float(constant), True)
            info["stages"] = S.api.last_timings()
            return out
        arr = D.build_and_merge_sharded(backend, src, names_all, rank * G, k, True, strategy=os.environ.get("BENCH_STRATEGY", "rows_lib"))
        info["rows_built"] = arr.dims()["n_rows"]
        out = D.distance_sharded(arr, 0.0, True, as_array=True)
        info["rows_compared"] = arr.dims()["n_rows"]
        info["stages"] = S.api.last_timings()
        return out

    batch = S.SeqBatch(packed)
    for _ in range(args.warmup):
        step(batch)
    barrier()
    l0 = S.api.launch_count()
    with B.ClockSampler(local_rank, 0.05) as clocks:
        S.api.timer_start()
        for _ in range(args.steps):
            dists = step(batch)
        ms_dev = S.api.timer_stop()
    barrier()
    launches = S.api.launch_count() - l0
    ms_dev = max_over_ranks(ms_dev) / args.steps
    stages = {s_: ms for s_, ms, _ in info["stages"]}
    B.log(f"device arm: {ms_dev:.2f} ms/step; rows built {info['rows_built']}, compared {info['rows_compared']}; stages "
          f"{[(s_, round(ms, 3)) for s_, ms, _ in info['stages']]}")
    del batch
    # e2e: packed host buffers in, the distance table out (it is computed on the host from the counters).  Two-deep pipeline as
    # in the default workload: the upload of step i+1 is queued before the kernels of step i; every upload is inside the timed region
    pre = [None]

    def e2e_step(last):
        cur = pre[0] or S.SeqBatch(packed, asynchronous=True)
        pre[0] = None if last else S.SeqBatch(packed, asynchronous=True)
        return step(cur)

    for i in range(3):
        e2e_step(i == 2)
    barrier()
    S.api.timer_start()
    for i in range(args.steps):
        dists = e2e_step(i + 1 == args.steps)
    ms_e2e = max_over_ranks(S.api.timer_stop()) / args.steps
    n_pairs = G * world * (G * world - 1) // 2
    # parity: the oracle's TSV on a sample of rank 0's genomes against the GPU path on the same sample
    parity = {"checked": True}
    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as O
        n_cpu = min(args.cpu_genomes, 24 if world == 1 else 8)
        cores = os.cpu_count() or 1
        threads = B.host_threads_rule(cores, n_cpu)
        sub = [f"s{i}" for i in range(n_cpu)]
        t0 = time.perf_counter()
        oarr = O.OracleArray.build([O.SampleData([(g.tobytes(), None)]) for g in genomes[:n_cpu]], sub, k=k, rc=True, threads=threads)
        ref = oarr.distance(threads=cores)
        dt = time.perf_counter() - t0
        garr = S.api.build_from_packed(packed[:n_cpu], sub, k, True)
        buf = io.BytesIO()
        S.api.distance(garr, buf)
        same = buf.getvalue() == ref
        parity["oracle_sample"] = f"build+distance of the first {n_cpu} genomes: {len(ref)} bytes of TSV " + ("identical" if same else "DIFFER")
        cpu_baseline = {"value": n_cpu / dt, "unit": unit, "cores": cores, "kind": "port",
                        "sample": f"first {n_cpu} genomes, build on {threads} threads (reference's rule) + distance on all {cores} cores"}
        if not same:
            raise SystemExit(f"parity check FAILED: {parity}")
    # invariants at full size: distance <= matches, matches + mismatches <= constant + rows, symmetric population sanity
    if isinstance(dists, np.ndarray):
        d, m, mm = dists["distance"], dists["matches"], dists["mismatches"]
    else:
        d = np.array([x[0] for x in dists]); m = np.array([x[2] for x in dists]); mm = np.array([x[3] for x in dists])
    parity["full_size"] = {"pairs": int(len(dists)), "distance_le_matches": bool(np.all(d <= m)), "max_distance": float(d.max()) if len(d) else 0.0,
                           "mean_distance": float(d.mean()) if len(d) else 0.0}
    if not np.all(d <= m) or len(dists) != n_pairs:
        raise SystemExit(f"parity check FAILED: {parity}")
    if rank == 0:
        words = (info["rows_compared"] + 31) // 32
        popc = 3.0 * n_pairs * words                           # per pair and word: mismatches, matches, distance
        pp_ms = stages.get("pair_popc", 0.0) or 1e-9
        line = {"metric": metric, "value": G * world / (ms_dev * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": dict(cfg, rows_built=info["rows_built"], rows_compared=info["rows_compared"], pairs=n_pairs,
                               parallelism="1 GPU" if world == 1 else f"{world} ranks: rows sharded by key range, per-pair integer counters all-reduced once"),
                "e2e": {"value": G * world / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(sum(len(p.packed2) + len(p.invalid) for p in packed)), "d2h_bytes_per_step": int(n_pairs * 24)},
                "gpu_launches": launches,
                "roofline": {"kernel": "pair_popc_kernel", "bound": "popc", "achieved": popc / (pp_ms * 1e-3) / 1e12, "peak": POPC_PEAK / 1e12,
                             "unit": "Tpopc/s", "frac": popc / (pp_ms * 1e-3) / POPC_PEAK, "traffic": None, "launch_ms": pp_ms,
                             "peak_source": "measured (profiles/microbench/popc_mb.cu on this pool's B200: 15.4 popc lanes/clk/SM)",
                             "algorithmic_popc_per_launch": popc, "share_of_step": pp_ms / ms_dev,
                             "note": "this rank's rows only; pairs on the diagonal tiles are computed twice (i<j kept), not counted here"},
                "stages_ms": {s_: round(ms, 3) for s_, ms in stages.items()},
                "clocks": clocks.summary(), "parity_checked": True, "parity": parity}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        _emit(line, rank)
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- viral (configs[4])
def run_viral(args, rank, world, local_rank):
    """256 x 30 kbp genomes, k = 15, single strand (BASELINE configs[4], SURVEY 8d C5: star model, 28-bit keys so repeats and
    IUPAC merges do occur).  The whole job is small enough for the oracle: parity is checked on the FULL alignment."""
    if rank != 0:
        return
    B = _bench()
    metric, unit = "genomes/sec build+align (k=15, single strand, 30 kbp genomes)", "genomes/s"
    G, L, k = 256, 30_000, 15
    from ska_rust_b200 import synth
    cores = os.cpu_count() or 1
    threads = B.host_threads_rule(cores, G)
    cfg = {"workload": f"ska build + align, {G} x {L} bp synthetic viral genomes, k={k}, single strand, 1 GPU (BASELINE configs[4])",
           "k": k, "rc": False, "genomes": G, "genome_bp": L, "population": "star model, 0.1% SNPs/sample",
           "align": "min_freq 0.9, filter no-const", "l2": "the whole job (7.7 MB of sequence) is L2-resident: launch latency, not bandwidth, bounds it"}
    genomes = synth.star_genomes(G, L, seed=0x5CA20005)
    names = [f"s{i}" for i in range(G)]

    def oracle_pass():
        from oracle import oracle as O
        arr = O.OracleArray.build([O.SampleData([(g.tobytes(), None)]) for g in genomes], names, k=k, rc=False, threads=threads)
        return arr.align()

    if args.impl == "reference":
        oracle_pass()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            out = oracle_pass()
        dt = (time.perf_counter() - t0) / reps
        _emit({"impl": "reference", "metric": metric, "value": G / dt, "unit": unit, "n_gpus": 1, "steps": reps, "warmup": 1, "ms_per_step": dt * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": cfg,
               "cpu_baseline": {"value": G / dt, "unit": unit, "cores": threads, "kind": "port",
                                "sample": f"the whole job, {reps} passes, {threads} threads (reference's rule); alignment {len(out)} bytes"},
               "e2e": {"value": G / dt, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}, rank)
        return
    import torch
    import ska_rust_b200 as S
    torch.cuda.set_device(local_rank)
    S.set_device(local_rank)
    off = np.array([0, L], dtype=np.uint64)
    pinned = torch.empty((G, L), dtype=torch.uint8, pin_memory=True).numpy()
    for i, g in enumerate(genomes):
        pinned[i] = g
    samples = [S.Sample(pinned[i], off) for i in range(G)]
    pin2 = torch.empty((G, (L + 3) // 4), dtype=torch.uint8, pin_memory=True).numpy()
    pinv = torch.empty((G, (L + 7) // 8), dtype=torch.uint8, pin_memory=True).numpy()
    packed = [S.PackedSample.from_sample(s, pin2[i], pinv[i]) for i, s in enumerate(samples)]
    threshold = int(np.ceil(G * 0.9))
    batch = S.SeqBatch(packed)
    info = {}

    def step(src, out=None):
        arr = S.api.build_from_batch(src, names, k, False) if isinstance(src, S.SeqBatch) else S.api.build_from_packed(src, names, k, False)
        info["rows_built"] = arr.dims()["n_rows"]
        arr.filter(threshold, False, "no-const", False, False, False)
        info["rows_aligned"] = arr.dims()["n_rows"]
        if out is None:
            arr.transpose_on_device()
        else:
            arr.transposed(out=out)
        info["stages"] = S.api.last_timings()

    for _ in range(max(args.warmup, 3)):
        step(batch)
    l0 = S.api.launch_count()
    steps = max(args.steps, 20)
    with B.ClockSampler(local_rank, 0.02) as clocks:
        S.api.timer_start()
        for _ in range(steps):
            step(batch)
        ms_dev = S.api.timer_stop() / steps
    launches = (S.api.launch_count() - l0) // steps
    dev_stages = {s_: round(ms, 4) for s_, ms, _ in info["stages"]}
    out_host = torch.empty((G, max(1, info["rows_aligned"])), dtype=torch.uint8, pin_memory=True).numpy()
    for _ in range(3):
        step(packed, out_host)
    S.api.timer_start()
    for _ in range(steps):
        step(packed, out_host)
    ms_e2e = S.api.timer_stop() / steps
    # parity on the whole job
    t0 = time.perf_counter()
    ref = oracle_pass()
    dt = time.perf_counter() - t0
    garr = S.api.build_from_packed(packed, names, k, False)
    buf = io.BytesIO()
    S.api.align(garr, buf)
    same = buf.getvalue() == ref
    parity = {"checked": True, "full_job": f"alignment of all {G} genomes: {len(ref)} bytes " + ("identical to the oracle's" if same else "DIFFER from the oracle's")}
    if not same:
        raise SystemExit(f"parity check FAILED: {parity}")
    stages = dev_stages
    peak, peak_src = B.measured_peak()
    k1_ms = stages.get("k1_extract", 0.0) or 1e-9
    _emit({"metric": metric, "value": G / (ms_dev * 1e-3), "unit": unit, "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": dict(cfg, rows_built=info["rows_built"], rows_aligned=info["rows_aligned"]),
           "e2e": {"value": G / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(sum(len(p.packed2) + len(p.invalid) for p in packed)),
                   "d2h_bytes_per_step": int(G * info["rows_aligned"])},
           "gpu_launches": launches,
           "roofline": {"kernel": "k1_extract_kernel", "bound": "hbm", "achieved": 9.375 * G * L / (k1_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": 9.375 * G * L / (k1_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "launch_ms": k1_ms,
                        "note": "a 7.7 Mbase batch is 49 us of HBM time at peak: every kernel of the step is shorter than its launch + tail; "
                                f"the step is {launches} launches and a few host synchronisations"},
           "stages_ms": stages, "clocks": clocks.summary(), "parity_checked": True, "parity": parity,
           "cpu_baseline": {"value": G / dt, "unit": unit, "cores": threads, "kind": "port", "sample": f"the whole job, one pass, {threads} threads"}}, rank)


# ---------------------------------------------------------------------------------------------- fastq (configs[3])
def _gpu_reads(genome_dev, n_pairs, seed, seq_out, qual_out, read_len=150, frag_len=400, err_rate=0.005, lowq_prob=0.5):
    """synth.paired_reads' model drawn with torch on the device -- 6e8 bases per sample is minutes of numpy --
    and copied into the host arrays seq_out / qual_out (file 1's reads, then file 2's)."""
    import torch
    dev = genome_dev.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    lut = torch.tensor(list(b"ACTG"), dtype=torch.uint8, device=dev)       # code (c >> 1) & 3 -> letter
    ar = torch.arange(read_len, device=dev)
    chunk = 250_000
    for f in range(2):
        for c0 in range(0, n_pairs, chunk):
            n = min(chunk, n_pairs - c0)
            gg = torch.Generator(device=dev)
            gg.manual_seed(seed * 1_000_003 + c0)                          # both files see the same fragment starts
            starts = torch.randint(0, genome_dev.numel() - frag_len + 1, (n,), generator=gg, device=dev)
            idx = starts[:, None] + (ar if f == 0 else (frag_len - 1 - ar))[None, :]
            reads = genome_dev[idx]
            if f == 1:
                reads = comp[reads.long()]
            err = torch.rand(reads.shape, generator=g, device=dev) < err_rate
            codes = (reads >> 1) & 3
            sub = (codes + torch.randint(1, 4, reads.shape, generator=g, device=dev, dtype=torch.uint8)) & 3
            reads = torch.where(err, lut[sub.long()], reads)
            qual = torch.full(reads.shape, ord("I"), dtype=torch.uint8, device=dev)
            qual[err & (torch.rand(reads.shape, generator=g, device=dev) < lowq_prob)] = ord("+")
            lo = (f * n_pairs + c0) * read_len
            torch.from_numpy(seq_out[lo:lo + n * read_len]).copy_(reads.reshape(-1))
            torch.from_numpy(qual_out[lo:lo + n * read_len]).copy_(qual.reshape(-1))


def _oracle_reads_sample(seq, qual, rec_off, n_rec_file1):
    from oracle import oracle as O
    sd = object.__new__(O.SampleData)                                      # the flat arrays as they are (no per-record Python objects)
    sd.seq, sd.qual, sd.rec_off = seq, qual, rec_off
    sd.n_rec, sd.n_rec_file1, sd.is_reads = len(rec_off) - 1, n_rec_file1, True
    return sd


def run_fastq(args, rank, world, local_rank):
    """ska build from paired FASTQ with the default count filter (BASELINE configs[3]: 64 samples x 2 M pairs x 150 bp,
    --min-count 5 --min-qual 20 --qual-filter strict).  The reads are drawn on the device, copied to host arrays, and enter the
    library the way parsed FASTQ files do (skagpu_seqbatch_upload of bases + qualities)."""
    if rank != 0:
        return
    import torch
    B = _bench()
    from ska_rust_b200 import synth
    NS, NP, L, RL, k = args.fastq_samples, args.fastq_pairs, args.length, 150, 31
    metric, unit = "read pairs/sec through ska build (paired FASTQ, min-count 5)", "read pairs/s"
    cores = os.cpu_count() or 1
    cfg = {"workload": f"ska build from FASTQ, {NS} samples x {NP} paired {RL} bp reads of a {L} bp genome (star model), k={k}, both strands, "
                       f"min-count 5 / min-qual 20 / strict (BASELINE configs[3])",
           "k": k, "rc": True, "samples": NS, "pairs_per_sample": NP, "read_len": RL, "genome_bp": L,
           "reads": "fragment 400, 0.5% substitution errors, Q40 with half of the error bases at Q10 (SURVEY 8d C4)",
           "l2": "every sample's working set (1.2 GB of reads, 19 GB of observation records) is far larger than L2"}
    bases = 2 * NP * RL
    genomes = synth.star_genomes(NS, L, seed=0x5CA20004)
    rec_off = np.arange(2 * NP + 1, dtype=np.uint64) * np.uint64(RL)
    names = [f"r{i}" for i in range(NS)]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def draw(i, seq_out, qual_out):
        _gpu_reads(torch.from_numpy(np.ascontiguousarray(genomes[i])).to(dev), NP, 1000 + i, seq_out, qual_out)

    if args.impl == "reference":
        # the reference's own path on the host cores: whole samples, as many as fit a bounded run, one thread per sample
        from oracle import oracle as O
        n_ref = min(NS, args.fastq_ref_samples, cores)
        seq = np.empty((n_ref, bases), dtype=np.uint8)
        qual = np.empty((n_ref, bases), dtype=np.uint8)
        for i in range(n_ref):
            draw(i, seq[i], qual[i])
        torch.cuda.empty_cache()
        sds = [_oracle_reads_sample(seq[i], qual[i], rec_off, NP) for i in range(n_ref)]
        threads = B.host_threads_rule(cores, n_ref)                        # merge_ska_dict.rs:381-385 applied to this many samples
        t0 = time.perf_counter()
        arr = O.OracleArray.build(sds, names[:n_ref], k=k, rc=True, min_count=5, min_qual=20, qual_filter="strict", threads=threads)
        dt = time.perf_counter() - t0
        v = n_ref * NP / dt
        _emit({"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": 1, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": cfg,
               "cpu_baseline": {"value": v, "unit": unit, "cores": threads, "kind": "port",
                                "sample": f"{n_ref} whole samples of the job ({n_ref * NP} pairs), one pass, {threads} thread(s): the reference's rule, 2^floor(log2(1 + n/10)), "
                                          f"gives {n_ref} samples {threads} and the whole job of {NS} samples {B.host_threads_rule(cores, NS)}; "
                                          f"{arr.dims()['n_rows']} split k-mers"},
               "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}, rank)
        return

    import ska_rust_b200 as S
    S.set_device(local_rank)
    q = S.QualOpts(min_count=5, min_qual=20, qual_filter="strict")
    t0 = time.perf_counter()
    try:                                                       # pinned host arrays (77 GB at full size): what a reader would fill
        seq = torch.empty((NS, bases), dtype=torch.uint8, pin_memory=True).numpy()
        qual = torch.empty((NS, bases), dtype=torch.uint8, pin_memory=True).numpy()
        host_kind = "pinned"
    except RuntimeError:
        seq = np.empty((NS, bases), dtype=np.uint8)
        qual = np.empty((NS, bases), dtype=np.uint8)
        host_kind = "pageable (pinning failed)"
    B.log(f"host arrays ({host_kind}) allocated in {time.perf_counter() - t0:.1f} s")
    t0 = time.perf_counter()
    for i in range(NS):
        draw(i, seq[i], qual[i])
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    t_draw = time.perf_counter() - t0
    samples = [S.Sample(seq[i], rec_off, qual[i], n_rec_file1=NP) for i in range(NS)]
    info = {}

    def step(src):
        arr = S.api.build_from_batch(src, names, k, True, q) if isinstance(src, S.SeqBatch) else S.build_from_samples(src, names, k, True, q)
        info["rows"] = arr.dims()["n_rows"]
        info["stages"] = S.api.last_timings()
        return arr

    batch = S.SeqBatch(samples)
    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        step(batch)
    steps = max(1, args.steps)
    l0 = S.api.launch_count()
    with B.ClockSampler(local_rank, 0.05) as clocks:
        S.api.timer_start()
        for _ in range(steps):
            step(batch)
        ms_dev = S.api.timer_stop() / steps
    launches = (S.api.launch_count() - l0) // steps
    stages = {}
    for s_, ms, _ in info["stages"]:
        stages[s_] = round(stages.get(s_, 0.0) + ms, 3)
    del batch
    # end to end: host arrays of bases + qualities in, the array's dimensions out
    e2e = None
    if not args.no_e2e:
        step(samples)
        S.api.timer_start()
        n_e2e = max(1, min(steps, 2))
        for _ in range(n_e2e):
            step(samples)
        ms_e2e = S.api.timer_stop() / n_e2e
        e2e = {"value": NS * NP / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(2 * NS * bases + NS * rec_off.nbytes),
               "d2h_bytes_per_step": 64, "note": f"{host_kind} host arrays of bases + qualities (77 GB at full size), uploaded and built inside the timed region"}
    # parity: the first samples of the job, whole, against the oracle
    n_par = min(NS, args.fastq_parity_samples)
    from oracle import oracle as O
    threads = B.host_threads_rule(cores, n_par)                            # the rule gives a handful of samples one thread
    t0 = time.perf_counter()
    oarr = O.OracleArray.build([_oracle_reads_sample(seq[i], qual[i], rec_off, NP) for i in range(n_par)], names[:n_par], k=k, rc=True,
                               min_count=5, min_qual=20, qual_filter="strict", threads=threads)
    dt_cpu = time.perf_counter() - t0
    garr = S.build_from_samples(samples[:n_par], names[:n_par], k, True, q)
    gk, gv, gc = garr.export()
    ok, ov, oc = oarr.export()
    same = (len(gk) == len(ok) and np.array_equal(np.asarray(gk, dtype=np.uint64), np.asarray(ok, dtype=np.uint64))
            and np.array_equal(gv, ov) and np.array_equal(gc, oc))
    parity = {"checked": True, "subsample": f"the first {n_par} samples, whole ({n_par * NP} pairs): {len(ok)} split k-mers, keys / variants / counts "
                                            + ("identical to the oracle's" if same else "DIFFER from the oracle's")}
    if not same:
        raise SystemExit(f"parity check FAILED: {parity}")
    peak, peak_src = B.measured_peak()
    n_obs = NS * 2 * NP * (RL - k + 1)
    dom = max((s_ for s_ in stages if s_ in ("reads_hash", "obs_group", "obs_sort", "count_filter", "k1_extract", "sample_unique")), key=lambda s_: stages[s_])
    alg = {"reads_hash": (2 * RL / (RL - k + 1) + 16, "2 B per base read (base + quality), 16 B observation record written per window"),
           "obs_group": (80, "16 B observation record read by two histogram passes, read + written by two scatter passes"),
           "obs_sort": (16 + 32 * 0.2, "16 B observation record read by the warp that sorts its bucket; a 32 B entry written per distinct hash (~0.2 per window)"),
           "count_filter": (16 + 16 + 1, "sorted observations read for the run heads, one 16 B distinct entry written per k-mer (upper bound: per window)"),
           "k1_extract": (2 * RL / (RL - k + 1) + 8, "2 B per base read, 8 B record written per accepted window"),
           "sample_unique": (40, "8 B record read + written by two scatter passes, read once by the bucket set")}[dom]
    ach = alg[0] * n_obs / (stages[dom] * 1e-3) / 1e9
    _emit({"metric": metric, "value": NS * NP / (ms_dev * 1e-3), "unit": unit, "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": ms_dev,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": dict(cfg, rows_built=info["rows"], windows=n_obs, draw_seconds=round(t_draw, 1)),
           "e2e": e2e, "gpu_launches": launches,
           "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                        "peak_source": peak_src, "launch_ms": stages[dom], "bytes_per_window": alg[0], "what": alg[1],
                        "note": "stage time summed over the samples of one step"},
           "stages_ms": stages, "clocks": clocks.summary(), "parity_checked": True, "parity": parity,
           "cpu_baseline": {"value": n_par * NP / dt_cpu, "unit": unit, "cores": threads, "kind": "port",
                            "sample": f"the parity subsample: {n_par} whole samples, one pass, {threads} thread(s) by the reference's rule "
                                      f"(the whole job of {NS} samples would get {B.host_threads_rule(cores, NS)})"}}, rank)


# ---------------------------------------------------------------------------------------------- cli (file to file)
def _write_fasta_fast(path, name, seq, width=80):
    """>name + 80-column lines, written with one numpy reshape instead of a Python loop per line"""
    n = len(seq)
    full = n // width * width
    body = np.empty((full // width, width + 1), dtype=np.uint8)
    body[:, :width] = seq[:full].reshape(-1, width)
    body[:, width] = 10
    with open(path, "wb") as fh:
        fh.write(b">" + name.encode() + b"\n")
        fh.write(body.tobytes())
        if n > full:
            fh.write(seq[full:].tobytes() + b"\n")


def _timed(cmd, **kw):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, capture_output=True, env=dict(os.environ, SKA_TIMING="1"), **kw)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise SystemExit(f"{cmd[:3]} failed ({p.returncode}): {p.stderr.decode()[-2000:]}")
    return dt, p.stderr.decode()


def _sha(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def run_cli(args, rank, world, local_rank):
    """`ska build -o x.skf` + `ska align x.skf -o x.aln` on FASTA files in tmpfs: wall-clock from process start (first input
    byte read) to process exit (last output byte written), both commands; the CPU arm is oracle/ska_ref_cli on the same files."""
    if rank != 0:
        return
    B = _bench()
    metric, unit = "genomes/sec ska build + ska align, file to file (CLI wall-clock, k=31)", "genomes/s"
    G, L, k = args.genomes, args.length, args.k
    from ska_rust_b200 import synth
    cores = os.cpu_count() or 1
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="ska_cli_", dir=base)
    cfg = {"workload": f"ska build + ska align on {G} x {L} bp FASTA files (80-column lines) in {'tmpfs' if base else 'a temporary directory'}, k={k}, "
                       "both strands; .skf written and read back between the two commands", "k": k, "rc": True, "genomes": G, "genome_bp": L,
           "population": B.population_text(args.population, G), "align": "min_freq 0.9, filter no-const", "host_cores": cores}
    try:
        genomes = synth.genomes(args.population, G, L, G, seed=args.seed)
        files = []
        for i, g in enumerate(genomes):
            f = os.path.join(tmp, f"s{i}.fa")
            _write_fasta_fast(f, f"s{i}", g)
            files.append(f)
        del genomes
        B.log(f"wrote {G} FASTA files to {tmp}")
        ref_cli = os.path.join(ROOT, "oracle", "ska_ref_cli")
        ours = os.path.join(ROOT, "ska.rust_b200", "ska")
        if args.impl == "reference":
            tb, eb = _timed([ref_cli, "build", "-o", os.path.join(tmp, "ref"), "-k", str(k), "--threads", str(cores)] + files)
            ta, ea = _timed([ref_cli, "align", os.path.join(tmp, "ref.skf"), "-o", os.path.join(tmp, "ref.aln")])
            v = G / (tb + ta)
            threads = B.host_threads_rule(cores, G)
            sample = (f"all {G} files: build {tb:.1f} s ({threads} threads, the reference's rule) + align {ta:.1f} s; phases: "
                      + " | ".join(l.split("] ")[1] for l in (eb + ea).splitlines() if l.startswith("[ska_ref_cli]")))
            _emit({"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                   "ms_per_step": (tb + ta) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                   "config": cfg, "cpu_baseline": {"value": v, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
                   "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}, rank)
            return
        runs = []
        for rep in range(max(1, args.steps)):
            for f in ("x.skf", "x.aln", "y.aln"):
                if os.path.exists(os.path.join(tmp, f)):
                    os.remove(os.path.join(tmp, f))
            tb, eb = _timed([ours, "build", "-o", os.path.join(tmp, "x"), "-k", str(k), "--threads", str(cores)] + files)
            ta, ea = _timed([ours, "align", os.path.join(tmp, "x.skf"), "-o", os.path.join(tmp, "x.aln"), "--threads", str(cores)])
            t1, e1 = _timed([ours, "align", "-o", os.path.join(tmp, "y.aln"), "--threads", str(cores)] + files)
            runs.append((tb, ta, t1))
            B.log(f"run {rep}: build {tb:.3f} s, align {ta:.3f} s, one-command align from FASTA {t1:.3f} s\n" + eb + ea)
        tb, ta, t1 = min(runs, key=lambda r: r[0] + r[1])
        skf_bytes = os.path.getsize(os.path.join(tmp, "x.skf"))
        aln_bytes = os.path.getsize(os.path.join(tmp, "x.aln"))
        parity = {"checked": True}
        same_one = _sha(os.path.join(tmp, "x.aln")) == _sha(os.path.join(tmp, "y.aln"))
        parity["one_command_align_equals_two_commands"] = same_one
        cpu_baseline = None
        if not args.no_cpu_baseline:
            n_cpu = min(G, args.cpu_genomes)
            threads = B.host_threads_rule(cores, n_cpu)
            rb, _ = _timed([ref_cli, "build", "-o", os.path.join(tmp, "ref"), "-k", str(k), "--threads", str(cores)] + files[:n_cpu])
            ra, _ = _timed([ref_cli, "align", os.path.join(tmp, "ref.skf"), "-o", os.path.join(tmp, "ref.aln")])
            cpu_baseline = {"value": n_cpu / (rb + ra), "unit": unit, "cores": threads, "kind": "port",
                            "sample": f"first {n_cpu} files: ska_ref_cli build {rb:.1f} s ({threads} threads) + align {ra:.1f} s"}
            # the same sample through our CLI, and our .skf read back by the CPU port
            _timed([ours, "build", "-o", os.path.join(tmp, "sub"), "-k", str(k), "--threads", str(cores)] + files[:n_cpu])
            _timed([ours, "align", os.path.join(tmp, "sub.skf"), "-o", os.path.join(tmp, "sub.aln")])
            _timed([ref_cli, "align", os.path.join(tmp, "sub.skf"), "-o", os.path.join(tmp, "sub_by_ref.aln")])
            a, b, c = (_sha(os.path.join(tmp, f)) for f in ("ref.aln", "sub.aln", "sub_by_ref.aln"))
            parity["oracle_sample"] = (f"{n_cpu} files: alignment of our CLI " + ("identical to" if a == b else "DIFFERS from") +
                                       " the CPU port's; our .skf read by the CPU port gives " + ("the same alignment" if a == c else "a DIFFERENT alignment"))
            if not (a == b == c):
                raise SystemExit(f"parity check FAILED: {parity}")
        if not same_one:
            raise SystemExit(f"parity check FAILED: {parity}")
        total = tb + ta
        line = {"metric": metric, "value": G / total, "unit": unit, "n_gpus": 1, "steps": len(runs), "warmup": 0, "ms_per_step": total * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": dict(cfg, threads=cores, skf_bytes=skf_bytes, alignment_bytes=aln_bytes),
                "e2e": {"value": G / total, "unit": unit, "ms_per_step": total * 1e3, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                        "build_s": tb, "align_s": ta, "one_command_align_from_fasta_s": t1,
                        "note": "process wall-clock of `ska build` + `ska align` (CUDA context creation of both processes included)"},
                "gpu_launches": None, "parity_checked": True, "parity": parity}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        _emit(line, rank)
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)


def run(args, rank, world, local_rank):
    fn = {"distance": run_distance, "cli": run_cli, "viral": run_viral, "fastq": run_fastq}.get(args.workload)
    if fn is None:
        raise SystemExit(f"workload {args.workload!r} is not built yet")
    return fn(args, rank, world, local_rank)
