import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import ska_rust_b200 as S
from oracle import oracle as O
gi = lambda n: os.path.join("tests/golden/ref/in", n)
k, rc = 41, True
files = [gi("test_1.fa"), gi("test_2.fa")]
samples = [S.Sample.from_files(f) for f in files]
names = ["a", "b"]
o = O.OracleArray.build_from_files(files, k=k, rc=rc, names=names)
ok, ov, oc = o.export()
for label, fn in (("samples", lambda: S.build_from_samples(samples, names, k, rc)), ("batch", lambda: S.api.build_from_batch(S.SeqBatch(samples), names, k, rc))):
    g = fn()
    gk, gv, gc = g.export()
    gk = [int(x) for x in gk]
    print(label, "rows", len(gk), "expected", len(ok), "common", len(set(gk) & set(ok)))
    extra = sorted(set(gk) - set(ok))[:5]
    missing = sorted(set(ok) - set(gk))[:5]
    print("  extra  ", [hex(x) for x in extra])
    print("  missing", [hex(x) for x in missing])
    print("  stages", S.api.last_timings())
b = S.SeqBatch(samples)
rec, _ = b.k1_extract(k, rc=rc)
rec = rec.reshape(-1, 2)
keys = sorted({(int(lo) | (int(hi) << 64)) >> 4 for lo, hi in rec if not (lo == 2**64 - 1 and hi == 2**64 - 1)})
print("k1 distinct keys", len(keys), "== oracle", keys == sorted(ok))
