"""Throughput probe of the FASTQ path (K1q/K2q + merge) at a reduced BASELINE configs[3] shape:
S samples x P read pairs x 150 bp from 1 Mbp star-model genomes (coverage = 300 P / 1e6), --min-count 5, strict quality.
    python profiles/fastq_probe.py [samples] [pairs]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ska_rust_b200 as S  # noqa: E402
from ska_rust_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
S.set_device(0)
genomes = synth.star_genomes(n, 1_000_000, seed=0x5CA20004)
samples = []
t0 = time.time()
for i, g in enumerate(genomes):
    (s1, q1, o1), (s2, q2, o2) = synth.paired_reads(g, pairs, err_rate=0.005, seed=1000 + i)
    samples.append(S.Sample(np.concatenate([s1, s2]), np.concatenate([o1, o2[1:] + o1[-1]]), np.concatenate([q1, q2]), n_rec_file1=pairs))
bases = sum(len(s.seq) for s in samples)
print(f"generated {n} samples x {pairs} pairs ({bases / 1e6:.0f} Mbase) in {time.time() - t0:.1f} s", flush=True)
names = [f"r{i}" for i in range(n)]
q = S.QualOpts(min_count=5, min_qual=20, qual_filter="strict")
batch = S.SeqBatch(samples)
for it in range(3):
    S.api.timer_start()
    arr = S.api.build_from_batch(batch, names, 31, True, q)
    ms = S.api.timer_stop()
    rows = arr.dims()["n_rows"]
    print(f"run {it}: {ms:.1f} ms = {bases / ms / 1e6:.2f} Gbase/s, {2 * pairs * n / ms / 1e3:.2f} M reads/s; rows {rows}; stages "
          + " ".join(f"{s}={m:.2f}" for s, m, _ in S.api.last_timings()), flush=True)
    del arr
