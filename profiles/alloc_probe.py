import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ska_rust_b200 as S
from ska_rust_b200 import synth
G = 128; L = 5_000_000
genomes = synth.star_genomes(G, L, seed=0x5CA20002)
off = np.array([0, L], dtype=np.uint64)
samples = [S.Sample(g, off) for g in genomes]
names = [f"s{i}" for i in range(G)]
batch = S.SeqBatch(samples)
for it in range(8):
    t0 = time.perf_counter()
    arr = S.api.build_from_batch(batch, names, 31, True)
    arr.filter(116, False, "no-const", False, False, False)
    arr.transpose_on_device()
    st = S.api.last_timings()
    del arr
    print(f"iter {it}: wall {(time.perf_counter()-t0)*1e3:.1f} ms", [(s, round(ms, 1)) for s, ms, _ in st], flush=True)
