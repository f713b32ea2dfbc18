"""One full-size read set (2 M pairs x 150 bp of a 5 Mbp genome) through skagpu_build -- the command behind the FASTQ launch
lists / ncu captures under profiles/ (ncu ... python profiles/fastq_one.py [n_pairs] [builds] [min_count] [genome_bp])."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ska_rust_b200 as S
from profiles.workloads import _gpu_reads
from ska_rust_b200 import synth

NP = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
builds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
min_count = int(sys.argv[3]) if len(sys.argv) > 3 else 5
L = int(sys.argv[4]) if len(sys.argv) > 4 else 5_000_000
g = synth.star_genomes(1, L, seed=0x5CA20004)[0]
seq = np.empty(2 * NP * 150, dtype=np.uint8)
qual = np.empty_like(seq)
_gpu_reads(torch.from_numpy(np.ascontiguousarray(g)).cuda(), NP, 1000, seq, qual)
torch.cuda.synchronize()
torch.cuda.empty_cache()
off = np.arange(2 * NP + 1, dtype=np.uint64) * np.uint64(150)
batch = S.SeqBatch([S.Sample(seq, off, qual, n_rec_file1=NP)])
q = S.QualOpts(min_count=min_count, min_qual=20, qual_filter="strict")
for i in range(builds):
    S.api.timer_start()
    a = S.api.build_from_batch(batch, ["r0"], 31, True, q)
    ms = S.api.timer_stop()
    print(f"build {i}: {ms:.2f} ms, rows {a.dims()['n_rows']}", {n: round(t, 3) for n, t, _ in S.api.last_timings()}, flush=True)
