"""Small driver for ncu captures: one build + align of G genomes x L bp through the C ABI.
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/prof python profiles/profile_driver.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ska_rust_b200 as S  # noqa: E402
from ska_rust_b200 import synth  # noqa: E402

G = int(os.environ.get("PROF_GENOMES", "32"))
L = int(os.environ.get("PROF_LENGTH", "5000000"))
POP = os.environ.get("PROF_POPULATION", "tree")               # bench.py's default population
genomes = synth.genomes(POP, G, L, G, seed=0x5CA20002)
off = np.array([0, L], dtype=np.uint64)
samples = [S.Sample(g, off) for g in genomes]
names = [f"s{i}" for i in range(G)]
if os.environ.get("PROF_ASCII"):
    batch = S.SeqBatch(samples)
else:                                                         # host-packed input, as bench.py feeds it
    batch = S.SeqBatch([S.PackedSample.from_sample(s) for s in samples])
arr = S.api.build_from_batch(batch, names, 31, True)
print(S.api.last_timings())
arr.filter(int(np.ceil(G * 0.9)), False, "no-const", False, False, False)
arr.transpose_on_device()
print(arr.dims())
