import os, sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
if os.environ.get("PROBE_TORCH"): import torch
from oracle import p3_oracle as O
O.build(native=True)
f = 1
m = O.random_matrix(f, 1 << 20, 100, seed=1)
O.coset_lde_batch(f, m, 1, O.generator(f))
ts = []
for _ in range(3):
    t = time.time(); O.coset_lde_batch(f, m, 1, O.generator(f)); ts.append(time.time() - t)
print(os.environ.get("OMP_PROC_BIND"), os.environ.get("OMP_NUM_THREADS"), bool(os.environ.get("PROBE_TORCH")), ["%.2f" % x for x in ts])
