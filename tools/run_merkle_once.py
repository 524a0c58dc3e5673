import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from plonky3_b200 import _lib
from plonky3_b200.field import KoalaBear as KB, BabyBear as BB
from plonky3_b200.gpu import default_gpu
from plonky3_b200.poseidon2 import default_poseidon2
gpu = default_gpu(0)
for f in (KB, BB):
    for w in (16, 24):
        default_poseidon2(f, w).upload(gpu)
x = torch.randint(0, KB.P, (1 << 20, 100), device="cuda", dtype=torch.int32)
gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W16, [x])
x2 = torch.randint(0, KB.P, (1 << 19, 328), device="cuda", dtype=torch.int32)
gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W24, [x2])
gpu.merkle_commit(BB.id, _lib.HASH_KECCAK, [x])
gpu.merkle_commit(BB.id, _lib.HASH_POSEIDON2_W16, [x])
torch.cuda.synchronize()
