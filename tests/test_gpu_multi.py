"""Row-sharded build across 2 GPUs (NCCL) must reproduce the single-GPU array exactly.  Skipped unless the
box has >= 2 GPUs (run with `gpurun --gpus 2`)."""
import io
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, n, length, k, strategy, ret):
    import torch
    import torch.distributed as dist

    import ska_rust_b200 as S
    import ska_rust_b200.distributed as D
    from ska_rust_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    S.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        genomes = synth.star_genomes(n, length, seed=0x5CA20009, snp_rate=0.002)
        # a few N runs and a second contig so sentinels and record boundaries cross the exchange too
        seqs = []
        for i, g in enumerate(genomes):
            g = g.copy()
            g[1000 * (i + 1):1000 * (i + 1) + 37] = ord("N")
            seqs.append([g[:length // 2].tobytes(), g[length // 2:].tobytes()])
        names = [f"s{i}" for i in range(n)]
        per = n // world
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else n
        local = [S.Sample.from_records(seqs[i]) for i in range(lo, hi)]
        backend = D.GpuBackend()
        slab = D.build_and_merge_sharded(backend, local, names, lo, k, True, strategy=strategy)
        keys, var, cnt = slab.export()
        parts = [None] * world if rank == 0 else None
        dist.gather_object((np.asarray(keys), var, cnt), parts, dst=0)
        slab_d = D.build_and_merge_sharded(backend, local, names, lo, k, True, prefix_bits=10, strategy=strategy)
        dists = D.distance_sharded(slab_d, filt_ambig=True)
        aln = D.gather_alignment(D.align_sharded(slab), names)
        if rank == 0:
            full = S.build_from_samples([S.Sample.from_records(s) for s in seqs], names, k, True)
            fk, fv, fc = full.export()
            ok = (np.array_equal(np.concatenate([p[0] for p in parts]), fk) and
                  np.array_equal(np.concatenate([p[1] for p in parts]), fv) and
                  np.array_equal(np.concatenate([p[2] for p in parts]), fc))
            buf = io.BytesIO()
            S.api.align(full, buf)
            full2 = S.build_from_samples([S.Sample.from_records(s) for s in seqs], names, k, True)
            dbuf = io.BytesIO()
            S.api.distance(full2, dbuf)
            ret["array_equal"] = bool(ok)
            ret["rows"] = [int(len(p[0])) for p in parts]
            ret["aln_equal"] = aln == buf.getvalue()
            ret["dist_equal"] = D.format_distances(names, dists) == dbuf.getvalue()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("strategy", ["rows", "records"])
def test_two_gpu_build_matches_single_gpu(strategy):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(2, free_port(), 8, 200_000, 31, strategy, ret), nprocs=2, join=True)
    assert ret["array_equal"], "concatenated slabs differ from the single-GPU array"
    assert min(ret["rows"]) > 0.3 * max(ret["rows"]), f"unbalanced slabs {ret['rows']}"
    assert ret["aln_equal"] and ret["dist_equal"]


@pytest.mark.parametrize("world,k,prefix_bits", [(3, 31, None), (2, 15, 7), (8, 31, 12), (2, 9, 0)])
def test_shard_path_emulated_on_one_gpu(world, k, prefix_bits):
    """The device steps of the sharded build (first pass per 'rank', exchange plan, second pass over the received
    runs + merge) driven sequentially on ONE GPU, with the all-to-all replaced by tensor copies: the slabs must
    concatenate to the single-GPU array.  Covers the shard kernels where only one GPU is available."""
    import torch

    import ska_rust_b200 as S
    import ska_rust_b200.distributed as D
    from ska_rust_b200 import synth
    S.set_device(0)
    n, length = 9, 60_000
    genomes = synth.tree_genomes(n, length, n, seed=0x5CA20011, snp_rate=0.004)
    seqs = []
    for i, g in enumerate(genomes):
        g = g.copy()
        g[500 * (i + 1):500 * (i + 1) + 20] = ord("N")
        seqs.append([g[:length // 3].tobytes(), g[length // 3:].tobytes()])
    names = [f"s{i}" for i in range(n)]
    backend = D.GpuBackend()
    per = (n + world - 1) // world
    bounds = [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]
    if prefix_bits is None:
        prefix_bits = backend.prefix_bits(sum(len(a) + len(b) for a, b in seqs), k, n)
    recs, offs = [], []
    for lo, hi in bounds:
        if hi == lo:                                          # a rank without samples sends nothing
            recs.append(torch.empty(0, dtype=torch.int64, device="cuda"))
            offs.append(None)
            continue
        rec, off = backend.extract([S.Sample.from_records(seqs[i]) for i in range(lo, hi)], k, True, lo, n, prefix_bits)
        recs.append(rec.clone())                              # the library's buffer is reused by the next extract
        offs.append(off)
    width = next(len(o) for o in offs if o is not None)
    all_off = np.stack([o if o is not None else np.zeros(width, dtype=np.uint64) for o in offs]).astype(np.int64)
    parts = []
    for g in range(world):
        cuts, send, recv, runs = D.exchange_plan(all_off, world, g)
        lo_r, hi_r = int(cuts[g]), int(cuts[g + 1])
        buf = torch.cat([recs[r][int(all_off[r, lo_r]):int(all_off[r, hi_r])] for r in range(world)])
        assert buf.numel() == int(recv.sum())
        slab = backend.merge(buf, runs, lo_r, hi_r, prefix_bits, n, k, True, names)
        keys, var, cnt = slab.export()
        parts.append((np.asarray(keys), var, cnt))
    full = S.build_from_samples([S.Sample.from_records(s) for s in seqs], names, k, True)
    fk, fv, fc = full.export()
    assert np.array_equal(np.concatenate([p[0] for p in parts]), fk)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), fv)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), fc)


@pytest.mark.parametrize("general_union", [0, 1])
@pytest.mark.parametrize("world,k,n,rc", [(3, 31, 9, True), (2, 15, 8, False), (8, 31, 17, True), (2, 9, 5, True), (4, 21, 16, True)])
def test_rows_path_emulated_on_one_gpu(world, k, n, rc, general_union, monkeypatch):
    """"rows" strategy on ONE GPU: a local array per 'rank', the exchange plan, the all-to-all replaced by tensor
    copies, the owners' join on the key (rowmerge.cu).  Uneven column blocks (n % world != 0), word-aligned and
    unaligned column counts, and k = 9 where every rank holds almost every key.  SKA_UNION_GENERAL=1 forces the
    fallback of the key union (general partition + merge instead of cutting the sorted runs)."""
    import torch
    monkeypatch.setenv("SKA_UNION_GENERAL", str(general_union))

    import ska_rust_b200 as S
    import ska_rust_b200.distributed as D
    from ska_rust_b200 import synth
    S.set_device(0)
    length = 40_000
    genomes = synth.tree_genomes(n, length, n, seed=0x5CA20021, snp_rate=0.004)
    seqs = []
    for i, g in enumerate(genomes):
        g = g.copy()
        g[300 * (i + 1):300 * (i + 1) + 25] = ord("N")
        seqs.append([g[:length // 3].tobytes(), g[length // 3:].tobytes(), b"ACGT" * 12])
    names = [f"s{i}" for i in range(n)]
    backend = D.GpuBackend()
    per = (n + world - 1) // world
    bounds = [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]
    bounds = [(lo, hi) for lo, hi in bounds if hi > lo]
    world = len(bounds)
    region_bits = min(backend.region_bits_max, 2 * (k - 1))
    local, cums, tensors = [], [], []
    for lo, hi in bounds:
        arr = backend.build_local([S.Sample.from_records(seqs[i]) for i in range(lo, hi)], names[lo:hi], k, rc)
        local.append(arr)
        cums.append(backend.key_cumulative(arr, region_bits))
        tensors.append(backend.array_tensors(arr))
    all_cum = np.stack(cums)
    cols = np.array([hi - lo for lo, hi in bounds], dtype=np.int64)
    col_base = np.array([lo for lo, _ in bounds], dtype=np.int64)
    parts = []
    for g in range(world):
        cuts, send, recv = D.rows_exchange_plan(all_cum, world, g)
        b = all_cum[:, cuts]
        keys = torch.cat([tensors[r][0][int(b[r, g]):int(b[r, g + 1])] for r in range(world)])
        cnt = torch.cat([tensors[r][2][int(b[r, g]):int(b[r, g + 1])] for r in range(world)])
        var = torch.cat([tensors[r][1][int(b[r, g]) * int(cols[r]):int(b[r, g + 1]) * int(cols[r])] for r in range(world)])
        assert keys.numel() == int(recv.sum())
        slab = backend.merge_partial(keys, var, cnt, recv, cols, col_base, n, k, rc, names)
        sk, sv, sc = slab.export()
        parts.append((np.asarray(sk), sv, sc))
    full = S.build_from_samples([S.Sample.from_records(s) for s in seqs], names, k, rc)
    fk, fv, fc = full.export()
    assert np.array_equal(np.concatenate([p[0] for p in parts]), fk)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), fv)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), fc)
    assert sum(len(p[0]) > 0 for p in parts) >= min(world, 2)
