"""Single-GPU stand-in for one rank of an 8-GPU build: 1024 samples x 625 kbp = the 6.4e8 records and the
1024-column rows a rank merges at world 8 (no exchange).  Sweeps the bucket prefix bits.
usage: python profiles/wide_rows.py [n_samples] [length] [populations] [P list]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ska_rust_b200 as S
from ska_rust_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 625_000
pops = (sys.argv[3] if len(sys.argv) > 3 else "tree,star").split(",")
plist = (sys.argv[4] if len(sys.argv) > 4 else "auto,16,18,20").split(",")
S.set_device(0)
names = [f"s{i}" for i in range(n)]
off = np.array([0, L], dtype=np.uint64)
for pop in pops:
    t0 = time.time()
    gs = synth.genomes(pop, n, L, n)
    batch = S.SeqBatch([S.Sample(g, off) for g in gs])
    print(f"# {pop}: generated in {time.time() - t0:.1f}s", flush=True)
    for P in plist:
        if P == "auto":
            os.environ.pop("SKA_PREFIX_BITS", None)
        else:
            os.environ["SKA_PREFIX_BITS"] = P
        best = None
        for it in range(3):
            S.api.timer_start()
            arr = S.api.build_from_batch(batch, names, 31, True)
            rows = arr.dims()["n_rows"]
            st = S.api.last_timings()
            arr.filter(int(np.ceil(n * 0.9)), False, "no-const", False, False, False)
            arr.transpose_on_device()
            ms = S.api.timer_stop()
            kept = arr.dims()["n_rows"]
            del arr
            if best is None or ms < best[0]:
                best = (ms, st)
        print(f"{pop} N={n} L={L} P={P}: {best[0]:.2f} ms  rows {rows} kept {kept}  " +
              " ".join(f"{s}={m:.2f}" for s, m, _ in best[1]), flush=True)
    del batch
