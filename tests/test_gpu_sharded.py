"""The row-sharded multi-GPU commit (p3gpu_coset_lde_batch_sharded_dev / p3gpu_commit_sharded_dev: peer-memory stores from
the LDE's last pass, flag barrier, peer all-gather of the cap slices) exercised on ONE device: `world` ranks are simulated
inside this process, one libp3gpu context (= one stream) per rank, all row blocks and control blocks on cuda:0.  The
kernels, the addressing and the barrier protocol are exactly what runs across GPUs (there the pointers are CUDA-IPC
mappings; tests/test_gpu_multi.py covers that on a 2-GPU box)."""
import threading

import numpy as np
import pytest
import torch

from oracle import p3_oracle as O

from plonky3_b200 import _lib
from plonky3_b200.distributed import PeerGroup, column_block
from plonky3_b200.field import BabyBear, KoalaBear
from plonky3_b200.gpu import Gpu
from plonky3_b200.poseidon2 import default_poseidon2

pytestmark = pytest.mark.gpu


def host(t):
    return t.cpu().numpy().view(np.uint32)


def _gpus(n):
    out = []
    for _ in range(n):
        g = Gpu(0)
        for f in (BabyBear, KoalaBear):
            for w in (16, 24):
                default_poseidon2(f, w).upload(g)
        out.append(g)
    return out


@pytest.mark.parametrize("f,log_h,w,world", [(KoalaBear, 12, 100, 2), (BabyBear, 13, 72, 4), (KoalaBear, 14, 328, 8)])
def test_sharded_lde_scatters_row_blocks(f, log_h, w, world):
    """Every rank's column-block LDE lands in the right rows/columns of every rank's row block."""
    gpus = _gpus(world)
    H = 2 << log_h
    groups = PeerGroup.simulate(gpus, H // world, w)
    full = O.random_matrix(f.id, 1 << log_h, w, seed=11)
    exp = O.coset_lde_batch(f.id, full, 1, f.generator, bitrev_out=True)
    for q, grp in enumerate(groups):
        c0, c1 = column_block(w, world, q, align=8)
        local = torch.from_numpy(np.ascontiguousarray(full[:, c0:c1]).view(np.int32)).cuda()
        grp.lde_sharded(f, local, 1, f.generator, c0)
        torch.cuda.synchronize()
    got = np.concatenate([host(grp.rows_tensor()) for grp in groups], axis=0)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("kind,f,log_h,w,world,cap_height", [
    ("p2w24", KoalaBear, 12, 100, 2, 3), ("p2w24", KoalaBear, 12, 100, 2, 0), ("p2w16", BabyBear, 13, 40, 4, 1),
    ("keccak", BabyBear, 12, 60, 4, 3), ("p2w24", KoalaBear, 13, 164, 8, 3), ("p2w24", KoalaBear, 13, 164, 8, 5)])
def test_sharded_commit_equals_single_commit(kind, f, log_h, w, world, cap_height):
    """cap of the sharded commit (every rank) == cap of TwoAdicFriPcs::commit on the whole trace (oracle), and every rank's
    sub-tree == the corresponding slice of the oracle's tree."""
    hash_kind = {"p2w16": _lib.HASH_POSEIDON2_W16, "p2w24": _lib.HASH_POSEIDON2_W24, "keccak": _lib.HASH_KECCAK}[kind]
    ohs = O.keccak_hasher() if kind == "keccak" else O.poseidon2_hasher(O.default_perm(f.id, 24 if kind == "p2w24" else 16), O.default_perm(f.id, 16))
    gpus = _gpus(world)
    H = 2 << log_h
    groups = PeerGroup.simulate(gpus, H // world, w)
    full = O.random_matrix(f.id, 1 << log_h, w, seed=5)
    olayers = O.merkle_tree(ohs, [O.coset_lde_batch(f.id, full, 1, f.generator, bitrev_out=True)])
    exp_cap = O.merkle_cap(olayers, cap_height)
    locals_ = []
    for q in range(world):
        c0, c1 = column_block(w, world, q, align=8)
        locals_.append((c0, torch.from_numpy(np.ascontiguousarray(full[:, c0:c1]).view(np.int32)).cuda()))
    torch.cuda.synchronize()
    results, errors = [None] * world, []

    streams = [torch.cuda.Stream() for _ in range(world)]

    def run(q):
        try:
            with torch.cuda.stream(streams[q]):              # one stream per simulated rank (the barrier kernel spins)
                for _ in range(2):                           # twice: the epoch counter and the row blocks are reused
                    results[q] = groups[q].commit(f, hash_kind, locals_[q][1], locals_[q][0], 1, cap_height, phases=True)
        except Exception as e:                               # noqa: BLE001 — surfaced below
            errors.append((q, repr(e)))

    ths = [threading.Thread(target=run, args=(q,)) for q in range(world)]
    for t in ths: t.start()
    for t in ths: t.join()
    assert not errors, errors
    rows = H // world
    for q in range(world):
        cap, layers, ph = results[q]
        assert np.array_equal(cap, exp_cap), f"rank {q}"
        assert len(ph) == 4 and all(p >= 0 for p in ph)
        for k, lay in enumerate(layers):                     # my sub-tree = slice q of the global tree's lower layers
            n = max(rows >> k, 1)
            assert np.array_equal(host(lay)[:n], olayers[k][q * n:(q + 1) * n]), (q, k)
