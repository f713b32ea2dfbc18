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
genomes = synth.star_genomes(G, L, seed=0x5CA20002)
off = np.array([0, L], dtype=np.uint64)
samples = [S.Sample(g, off) for g in genomes]
names = [f"s{i}" for i in range(G)]
batch = S.SeqBatch(samples)
arr = S.api.build_from_batch(batch, names, 31, True)
print(S.api.last_timings())
arr.filter(int(np.ceil(G * 0.9)), False, "no-const", False, False, False)
arr.transpose_on_device()
print(arr.dims())
