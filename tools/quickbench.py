"""Developer micro-benchmarks (CUDA events on torch's current stream).  Not the driver's bench (see bench.py)."""
import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from plonky3_b200 import _lib
from plonky3_b200.field import KoalaBear as KB, BabyBear as BB
from plonky3_b200.gpu import default_gpu
from plonky3_b200.poseidon2 import default_poseidon2

gpu = default_gpu(0)
for f in (KB, BB):
    for w in (16, 24):
        default_poseidon2(f, w).upload(gpu)


def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def rnd(f, h, w):
    return torch.randint(0, f.P, (h, w), device="cuda", dtype=torch.int32)

which = sys.argv[1:] or ["lde", "merkle16", "merkle24", "keccak", "fold"]
if "lde" in which:
    x = rnd(KB, 1 << 20, 100)
    t, avg = timeit(lambda: gpu.coset_lde_batch(KB.id, x, 1, KB.generator))
    print(f"LDE KB 2^20x100 blowup2: best {t:.3f} ms avg {avg:.3f} ms  -> {209.7152/t:.1f} Gelem/s out, alg {1258.2912/t:.1f} GB/s")
    x2 = rnd(KB, 1 << 20, 128)
    t, avg = timeit(lambda: gpu.coset_lde_batch(KB.id, x2, 1, KB.generator))
    print(f"LDE KB 2^20x128 blowup2: best {t:.3f} ms avg {avg:.3f}")
    t, avg = timeit(lambda: gpu.dft_batch(KB.id, _lib.DFT, x2))
    print(f"DFT KB 2^20x128 nat->nat: best {t:.3f} ms")
if "merkle16" in which:
    x = rnd(KB, 1 << 22, 100)
    t, avg = timeit(lambda: gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W16, [x]), n=3, warm=1)
    print(f"Merkle KB 2^22x100 P2-16: best {t:.3f} ms -> {4.194304/t*1e3:.1f} Mleaf/s, {58.720255/t*1e3:.0f} Mperm/s")
if "merkle24" in which:
    x = rnd(KB, 1 << 21, 328)
    t, avg = timeit(lambda: gpu.merkle_commit(KB.id, _lib.HASH_POSEIDON2_W24, [x]), n=3, warm=1)
    print(f"Merkle KB 2^21x328 P2-24: best {t:.3f} ms -> {(2**21*21+2**21)/t/1e3:.0f} Mperm/s")
if "keccak" in which:
    x = rnd(BB, 1 << 22, 100)
    t, avg = timeit(lambda: gpu.merkle_commit(BB.id, _lib.HASH_KECCAK, [x]), n=3, warm=1)
    print(f"Merkle BB 2^22x100 Keccak: best {t:.3f} ms -> {(2**22*3+2**22)/t/1e3:.0f} Mperm/s")
if "fold" in which:
    v = rnd(KB, 1 << 21, 4)
    beta = np.array([5, 6, 7, 8], dtype=np.uint32)
    t, avg = timeit(lambda: gpu.fri_fold(KB.id, v, 3, beta))
    print(f"fold KB 2^21 arity 8: best {t:.3f} ms -> {(2**21*16*1.125)/t/1e6:.1f} GB/s")
print("launches", gpu.launches)
