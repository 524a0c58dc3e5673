"""Race hunt: repeat small LDEs against the oracle and report where mismatches fall."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from oracle import p3_oracle as O
from plonky3_b200.field import BabyBear, KoalaBear
from plonky3_b200.gpu import default_gpu
gpu = default_gpu(0)
shapes = [(KoalaBear, 21, 24, 1), (BabyBear, 21, 40, 1), (BabyBear, 12, 100, 1), (KoalaBear, 13, 52, 1), (BabyBear, 12, 100, 2), (KoalaBear, 14, 128, 1), (BabyBear, 16, 40, 2)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
if len(sys.argv) > 2: shapes = shapes[:int(sys.argv[2])]
for f, log_h, w, ab in shapes:
    m = O.random_matrix(f.id, 1 << log_h, w, seed=7 * log_h + w)
    want = O.coset_lde_batch(f.id, m, ab, f.generator, bitrev_out=True)
    x = torch.from_numpy(m.astype(np.int32)).cuda()
    bad = 0
    for r in range(reps):
        y = gpu.coset_lde_batch(f.id, x, ab, f.generator)
        got = y.cpu().numpy().astype(np.uint32)
        if not np.array_equal(got, want):
            bad += 1
            d = np.argwhere(got != want)
            rows, cols = np.unique(d[:, 0]), np.unique(d[:, 1])
            print(f"  {f.name} 2^{log_h}x{w} +{ab} rep {r}: {len(d)} wrong elements, rows {rows[:8]}..({len(rows)}), cols {cols[:16]}..({len(cols)})")
    print(f"{f.name} 2^{log_h}x{w} added_bits {ab}: {bad}/{reps} bad")
