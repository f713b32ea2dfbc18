"""world_size=2 (and 3) tests of the row-sharded orchestration (ska.rust_b200/distributed.py) on CPU with
the gloo backend.  The three device steps are replaced by an oracle-backed stand-in (test-only), so what
is tested here is the host logic: owner planning, the all-to-all exchange plan, slab assembly, the
distance all-reduce -- against the single-process oracle result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O

LETTERS = np.frombuffer(b"-ACMTWYHGRSVKDBN", dtype=np.uint8)


def revcomp_key(key, n_bases):
    out = 0
    for i in range(n_bases):
        out = (out << 2) | (((key >> (2 * i)) & 3) ^ 2)
    return out


class OracleSlab:
    """adapter giving an OracleArray the method names the orchestration uses"""

    def __init__(self, arr):
        self.arr = arr

    def nsamples(self):
        return self.arr.dims()["n_samples"]

    def names(self):
        return self.arr.names

    def filter(self, *a):
        return self.arr.filter(*a)

    def transposed(self):
        _, var, _ = self.arr.export()
        return np.ascontiguousarray(var.T)

    def export(self):
        return self.arr.export()

    def distance_counters(self, filt_ambig):
        _, var, _ = self.arr.export()
        n = var.shape[1]
        sets = {ord(c): m for c, m in zip("ACTGRYSWKMBDHV", [1, 2, 4, 8, 9, 6, 10, 5, 12, 3, 14, 13, 7, 11])}
        sets[ord("U")] = 4
        out = []
        for i in range(n):
            for j in range(i + 1, n):
                mism = match = dint = both = 0
                for a, b in zip(var[:, i], var[:, j]):
                    if a == 45 or b == 45:
                        mism += a != b
                    elif filt_ambig:
                        ua = chr(a).lower() in "acgtu"
                        ub = chr(b).lower() in "acgtu"
                        if ua and ub:
                            match += 1
                            dint += a != b
                    else:
                        both += 1
                        ma, mb = sets.get(int(a), 0), sets.get(int(b), 0)
                        inter = bin(ma & mb).count("1")
                        if inter:
                            match += 1
                            dint += 36 * inter // (bin(ma).count("1") * bin(mb).count("1"))
                out.append((mism, match, dint, both))
        return np.array(out, dtype=np.uint64).reshape(-1, 4)


class OracleBackend:
    device = "cpu"

    # stand-in record: key << 12 | sample << 4 | one-hot middle  (keys here are <= 28 bits)
    def extract(self, samples, k, rc, sample_base, n_samples_total, prefix_bits):
        recs = []
        for i, records in enumerate(samples):
            for r in records:
                w = O.windows(r, k, rc=rc)
                for key, mid in zip(w["key"], w["mid"]):
                    recs.append((key << 12) | ((sample_base + i) << 4) | (1 << int(mid)))
        rec = np.array(recs, dtype=np.uint64)
        # first pass only: grouped by REGION (the top region_bits of the key), any order inside a region
        rec = rec[np.random.default_rng(sample_base).permutation(len(rec))]
        region_bits = self.region_bits(prefix_bits)
        shift = 12 + 2 * (k - 1) - region_bits
        region = (rec >> np.uint64(shift)).astype(np.int64)
        order = np.argsort(region, kind="stable")
        rec, region = rec[order], region[order]
        off = np.searchsorted(region, np.arange((1 << region_bits) + 1)).astype(np.uint64)
        return torch.from_numpy(rec.view(np.int64)), off

    @staticmethod
    def region_bits(prefix_bits):
        return (prefix_bits + 1) // 2                     # like the library: the first of two passes

    def n_positions(self, samples):
        return sum(len(r) for recs in samples for r in recs)

    def prefix_bits(self, n_records_total, k, n_samples_total):
        return min(6, 2 * (k - 1))

    # ---- "rows" strategy stand-ins
    region_bits_max = 5

    def build_local(self, samples, names_local, k, rc):
        arr = O.OracleArray.build([O.SampleData([(r, None) for r in recs]) for recs in samples], names_local, k=k, rc=rc)
        keys, var, cnt = arr.export()
        assert list(keys) == sorted(keys)
        return OracleSlab(arr)

    def key_cumulative(self, slab, region_bits):
        keys = np.array(slab.export()[0], dtype=np.uint64)
        k = slab.arr.dims()["k"]
        cuts = np.arange((1 << region_bits) + 1, dtype=np.uint64) << np.uint64(2 * (k - 1) - region_bits)
        cum = np.searchsorted(keys, cuts, side="left").astype(np.int64)
        cum[-1] = len(keys)
        return cum

    def array_tensors(self, slab):
        keys, var, cnt = slab.export()
        return (torch.from_numpy(np.array(keys, dtype=np.uint64).view(np.int64)), torch.from_numpy(np.ascontiguousarray(var).reshape(-1)),
                torch.from_numpy(np.asarray(cnt).astype(np.int32)))

    def merge_partial(self, keys, var, cnt, run_rows, run_cols, run_col_base, n_samples_total, k, rc, names):
        keys = keys.numpy().view(np.uint64)
        var, cnt = var.numpy(), cnt.numpy()
        assert len(keys) == int(np.sum(run_rows)) == len(cnt) and len(var) == int(np.sum(run_rows * run_cols))
        uni = sorted(set(int(x) for x in keys))
        row = {key: i for i, key in enumerate(uni)}
        out = np.full((len(uni), n_samples_total), ord("-"), dtype=np.uint8)
        oc = np.zeros(len(uni), dtype=np.uint64)
        r0 = b0 = 0
        for rows, cols, base in zip(map(int, run_rows), map(int, run_cols), map(int, run_col_base)):
            rk = keys[r0:r0 + rows]
            assert np.all(rk[1:] > rk[:-1]), "every run must be ascending in key"
            block = var[b0:b0 + rows * cols].reshape(rows, cols)
            for i in range(rows):
                out[row[int(rk[i])], base:base + cols] = block[i]
                oc[row[int(rk[i])]] += int(cnt[r0 + i])
            r0 += rows
            b0 += rows * cols
        return OracleSlab(O.OracleArray.from_arrays(k, rc, 64, uni, out, oc, names))

    def merge(self, rec, run_offsets, region_lo, region_hi, prefix_bits, n_samples_total, k, rc, names):
        rec = rec.numpy().view(np.uint64)
        # the exchange plan must describe the receive buffer exactly: every run grouped by region, regions in range
        shift = 12 + 2 * (k - 1) - self.region_bits(prefix_bits)
        nr = region_hi - region_lo
        assert run_offsets.shape[1] == nr + 1
        covered = 0
        for r in range(run_offsets.shape[0]):
            assert int(run_offsets[r, 0]) == covered, "runs must tile the receive buffer in rank order"
            for j in range(nr):
                seg = rec[int(run_offsets[r, j]):int(run_offsets[r, j + 1])]
                assert np.all((seg >> np.uint64(shift)) == np.uint64(region_lo + j))
            covered = int(run_offsets[r, -1])
        assert covered == len(rec)
        keys = sorted(set(int(r) >> 12 for r in rec))
        row = {key: i for i, key in enumerate(keys)}
        mask = np.zeros((len(keys), n_samples_total), dtype=np.uint8)
        for r in rec:
            mask[row[int(r) >> 12], (int(r) >> 4) & 0xFF] |= int(r) & 0xF
        var = LETTERS[mask]
        for i, key in enumerate(keys):
            if rc and revcomp_key(key, k - 1) == key:
                for s in range(n_samples_total):
                    m = int(mask[i, s])
                    at, cg = m & 5, m & 10
                    var[i, s] = ord("-") if not m else ord("N") if (at and cg) else ord("W") if at else ord("S")
        cnt = (var != ord("-")).sum(axis=1).astype(np.uint64)
        return OracleSlab(O.OracleArray.from_arrays(k, rc, 64, keys, var, cnt, names))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_population(n, seed):
    rng = np.random.default_rng(seed)
    anc = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=1500)
    pops = []
    for _ in range(n):
        g = anc.copy()
        pos = rng.integers(0, len(g), size=12)
        g[pos] = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=12)
        recs = [g[:700].tobytes(), g[700:].tobytes()]
        if rng.random() < 0.5:
            recs.append(g[100:400].tobytes())            # repeated region -> IUPAC merges
        pops.append(recs)
    return pops


def worker(rank, world, port, k, rc, strategy, ret):
    import ska_rust_b200.distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 7
        pops = make_population(n, 123)
        names = [f"s{i}" for i in range(n)]
        per = (n + world - 1) // world
        lo, hi = rank * per, min(n, (rank + 1) * per)
        backend = OracleBackend()
        slab = D.build_and_merge_sharded(backend, pops[lo:hi], names, lo, k, rc, strategy=strategy)
        keys, var, cnt = slab.export()
        parts = [None] * world if rank == 0 else None
        dist.gather_object((keys, var, cnt), parts, dst=0)
        # distance on a second identical build (filters mutate the slab)
        slab_d = D.build_and_merge_sharded(backend, pops[lo:hi], names, lo, k, rc, prefix_bits=6, strategy=strategy)
        dists = D.distance_sharded(slab_d, filt_ambig=True)
        slab_a = D.build_and_merge_sharded(backend, pops[lo:hi], names, lo, k, rc, prefix_bits=6, strategy=strategy)
        dists_amb = D.distance_sharded(slab_a, filt_ambig=False)
        aln = D.gather_alignment(D.align_sharded(slab, min_freq=0.5), names)
        if rank == 0:
            ret["parts"] = parts
            ret["dists"] = dists
            ret["dists_amb"] = dists_amb
            ret["aln"] = aln
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("strategy", ["rows", "records"])
@pytest.mark.parametrize("world,k,rc", [(2, 9, True), (3, 15, True), (2, 11, False)])
def test_sharded_matches_single_process(world, k, rc, strategy):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(world, free_port(), k, rc, strategy, ret), nprocs=world, join=True)
    n = 7
    pops = make_population(n, 123)
    names = [f"s{i}" for i in range(n)]
    samples = [O.SampleData([(r, None) for r in recs]) for recs in pops]
    full = O.OracleArray.build(samples, names, k=k, rc=rc)
    fk, fv, fc = full.export()
    parts = ret["parts"]
    keys = [key for p in parts for key in p[0]]
    assert keys == fk, "slabs must be ascending key ranges in rank order"
    assert np.array_equal(np.concatenate([p[1] for p in parts]), fv)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), fc)
    assert all(len(p[0]) > 0 for p in parts), "every rank should own some rows"
    # distances: identical integers; alignment: identical bytes
    want = O.OracleArray.build(samples, names, k=k, rc=rc)
    constant = want.apply_filters(0.0, False, "no-const", False, False)
    assert ret["dists"] == want.distance_raw(float(constant), True)
    want2 = O.OracleArray.build(samples, names, k=k, rc=rc)
    constant = want2.apply_filters(0.0, False, "no-const", False, False)
    for a, b in zip(ret["dists_amb"], want2.distance_raw(float(constant), False)):
        assert abs(a[0] - b[0]) < 1e-9 and a[1:] == b[1:]
    assert ret["aln"] == O.OracleArray.build(samples, names, k=k, rc=rc).align(min_freq=0.5)


def test_rows_exchange_plan_is_consistent_across_ranks():
    """rows strategy: what rank g expects from rank r == what rank r sends to rank g; all rows are sent exactly once"""
    import ska_rust_b200.distributed as D
    rng = np.random.default_rng(2)
    for world in (1, 2, 5, 8):
        counts = rng.integers(0, 50, size=(world, 64))
        counts[:, 10:20] = 0
        all_cum = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(counts, axis=1)], axis=1)
        plans = [D.rows_exchange_plan(all_cum, world, r) for r in range(world)]
        for g in range(world):
            cuts, send, recv = plans[g]
            assert np.array_equal(cuts, plans[0][0]) and cuts[0] == 0 and cuts[-1] == 64
            assert send.sum() == all_cum[g, -1] and np.all(send >= 0)
            assert [int(plans[r][1][g]) for r in range(world)] == [int(x) for x in recv]
            assert recv.sum() == counts[:, cuts[g]:cuts[g + 1]].sum()


def test_plan_owners_balances_and_is_monotone():
    import ska_rust_b200.distributed as D
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 1000, size=256)
    counts[:32] *= 8                                     # canonical k-mers crowd the low prefixes
    for world in (1, 2, 3, 8):
        cuts = D.plan_owners(counts, world)
        assert cuts[0] == 0 and cuts[-1] == 256 and np.all(np.diff(cuts) >= 0)
        loads = [counts[cuts[g]:cuts[g + 1]].sum() for g in range(world)]
        assert max(loads) - min(loads) <= 2 * counts.max()
    assert list(D.plan_owners(np.zeros(16, dtype=np.int64), 4)) == [0, 0, 0, 0, 16]


def test_exchange_plan_is_consistent_across_ranks():
    """what rank g expects from rank r == what rank r sends to rank g; run offsets tile the receive buffer"""
    import ska_rust_b200.distributed as D
    rng = np.random.default_rng(1)
    for world in (1, 2, 5, 8):
        counts = rng.integers(0, 50, size=(world, 64))
        counts[:, 40:48] = 0                              # empty regions
        all_off = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(counts, axis=1)], axis=1)
        plans = [D.exchange_plan(all_off, world, r) for r in range(world)]
        for g in range(world):
            cuts, send, recv, runs = plans[g]
            assert np.array_equal(cuts, plans[0][0])
            assert send.sum() == all_off[g, -1]
            assert [int(plans[r][1][g]) for r in range(world)] == [int(x) for x in recv]
            lo, hi = int(cuts[g]), int(cuts[g + 1])
            assert runs.shape == (world, hi - lo + 1)
            assert int(runs[0, 0]) == 0 and int(runs[-1, -1]) == recv.sum()
            assert np.array_equal(np.diff(runs.astype(np.int64), axis=1), counts[:, lo:hi])
            if world > 1:
                assert np.array_equal(runs[1:, 0], runs[:-1, -1])
