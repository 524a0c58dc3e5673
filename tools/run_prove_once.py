"""One uni-stark prove of the config-5 statement at 2^L rows (target for ncu: trace generation, quotient, leaf hashing kernels)."""
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from plonky3_b200.dft import Radix2DitParallel
from plonky3_b200.field import KoalaBear as KB
from plonky3_b200.fri import FriParameters, TwoAdicFriPcs
from plonky3_b200.gpu import default_gpu
from plonky3_b200.merkle_tree import MerkleTreeMmcs
from plonky3_b200.poseidon2 import default_poseidon2
from plonky3_b200.uni_stark import RoundConstants, StarkConfig, VectorizedPoseidon2Air, prove

L = int(sys.argv[1]) if len(sys.argv) > 1 else 18
gpu = default_gpu(0)
mm = MerkleTreeMmcs.poseidon2(default_poseidon2(KB, 16), default_poseidon2(KB, 24), 3, gpu)
cfg = StarkConfig(TwoAdicFriPcs(Radix2DitParallel(KB, gpu), mm, FriParameters.new_benchmark_high_arity(mm)), default_poseidon2(KB, 24), 16)
rs = np.random.default_rng(7)
air = VectorizedPoseidon2Air(KB, RoundConstants(rs.integers(0, KB.P, (4, 16), dtype=np.uint32), rs.integers(0, KB.P, 20, dtype=np.uint32),
                                                rs.integers(0, KB.P, (4, 16), dtype=np.uint32)), gpu)
inputs = torch.randint(0, KB.P, (8 << L, 16), device="cuda", dtype=torch.int32)
trace = air.generate_trace_rows(inputs)
p = prove(cfg, air, trace)
torch.cuda.synchronize()
print("done", p.timings_ms)
