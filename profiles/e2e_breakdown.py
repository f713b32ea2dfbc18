"""Where does the end-to-end (host buffers in, alignment out) time go?  Prints per-stage device timings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ska_rust_b200 as S
from ska_rust_b200 import synth
G = int(os.environ.get("PROF_GENOMES", "128")); L = 5_000_000
pinned = torch.empty((G, L), dtype=torch.uint8, pin_memory=True)
genomes = synth.star_genomes(G, L, seed=0x5CA20002, out=pinned.numpy())
off = np.array([0, L], dtype=np.uint64)
samples = [S.Sample(g, off) for g in genomes]
names = [f"s{i}" for i in range(G)]
thr = int(np.ceil(G * 0.9))
out_host = None
for it in range(3):
    t0 = time.perf_counter(); S.api.timer_start()
    arr = S.build_from_samples(samples, names, 31, True)
    t_build = S.api.timer_stop(); w_build = time.perf_counter() - t0
    st = S.api.last_timings()
    t0 = time.perf_counter(); S.api.timer_start()
    arr.filter(thr, False, "no-const", False, False, False)
    t_filter = S.api.timer_stop(); w_filter = time.perf_counter() - t0
    r = arr.dims()["n_rows"]
    if out_host is None:
        out_host = torch.empty((G, r), dtype=torch.uint8, pin_memory=True).numpy()
    t0 = time.perf_counter(); S.api.timer_start()
    arr.transposed(out=out_host)
    t_tr = S.api.timer_stop(); w_tr = time.perf_counter() - t0
    t0 = time.perf_counter(); del arr; w_free = time.perf_counter() - t0
    print(f"iter {it}: build dev {t_build:.1f} ms wall {w_build*1e3:.1f} | filter dev {t_filter:.1f} wall {w_filter*1e3:.1f} | "
          f"transpose+d2h dev {t_tr:.1f} wall {w_tr*1e3:.1f} | free wall {w_free*1e3:.1f}")
    print("   stages:", [(s, round(ms, 2)) for s, ms, _ in st])
# raw copy speeds
x = torch.empty(640_000_000, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); x.copy_(pinned.view(-1), non_blocking=True); torch.cuda.synchronize()
    print("torch H2D 640MB pinned: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
    t0 = time.perf_counter(); pinned.view(-1).copy_(x, non_blocking=True); torch.cuda.synchronize()
    print("torch D2H 640MB pinned: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
