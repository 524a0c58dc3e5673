import sys, pathlib, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from plonky3_b200 import _lib
from plonky3_b200.field import BabyBear as BB, KoalaBear as KB
from plonky3_b200.gpu import default_gpu
from plonky3_b200.poseidon2 import default_poseidon2
gpu = default_gpu(0)
for f in (KB, BB):
    for w in (16, 24):
        default_poseidon2(f, w).upload(gpu)
betas = np.random.default_rng(2).integers(0, BB.P, size=(8, 4), dtype=np.uint32)
v0 = torch.randint(0, BB.P, (1 << 21, 4), device="cuda", dtype=torch.int32)
for hk, name in ((_lib.HASH_KECCAK, "keccak"), (_lib.HASH_POSEIDON2_W16, "p2w16")):
    for it in range(3):
        v = v0.clone(); torch.cuda.synchronize(); t = time.time()
        gpu.fri_commit_phase(BB.id, hk, v, 1, 0, 3, 3, betas)
        torch.cuda.synchronize(); print(name, "fri_commit_phase wall ms", (time.time() - t) * 1e3)
