"""Runs the BASELINE configs[1] LDE a few times (target for ncu)."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from plonky3_b200.field import KoalaBear as KB
from plonky3_b200.gpu import default_gpu
gpu = default_gpu(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = int(sys.argv[2]) if len(sys.argv) > 2 else 100
x = torch.randint(0, KB.P, (1 << 20, w), device="cuda", dtype=torch.int32)
for _ in range(n):
    y = gpu.coset_lde_batch(KB.id, x, 1, KB.generator)
torch.cuda.synchronize()
print("done", gpu.launches)
