"""A/B of two builds of libp3gpu (P3GPU_LIB): Merkle commits at the config-3 and config-5 leaf shapes.  Usage: P3GPU_LIB=... python tools/hash_ab.py"""
import pathlib, sys, time
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
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
x = torch.randint(0, KB.P, (1 << 22, 100), device="cuda", dtype=torch.int32)
print("KB w16 2^22x100 ms", t(lambda: gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W16, [x])))
xb = torch.randint(0, BB.P, (1 << 22, 100), device="cuda", dtype=torch.int32)
print("BB w16 2^22x100 ms", t(lambda: gpu.merkle_commit(BB.id, _lib.HASH_POSEIDON2_W16, [xb])))
del x, xb
x2 = torch.randint(0, KB.P, (1 << 21, 1312), device="cuda", dtype=torch.int32)
print("KB w24 2^21x1312 ms", t(lambda: gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W24, [x2]), 3))
print("lib", _lib.LIB_PATH)
