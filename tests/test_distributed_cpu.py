"""world_size-2 gloo tests (CPU) of the multi-GPU sharding/collective logic in plonky3_b200.distributed.
The compute backend is the CPU oracle here (tests may use it); on the GPU box the same functions run with GpuBackend
over NCCL (tests/test_gpu_multi.py, bench.py --gpus N)."""
import os
import sys
import pathlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

F = 1  # KoalaBear
LOG_H, W, LOG_BLOWUP, CAP_H = 6, 11, 1, 1


class OracleBackend:
    def __init__(self):
        from oracle import p3_oracle as O
        self.O = O
        self.hs = O.poseidon2_hasher(O.default_perm(F, 24), O.default_perm(F, 16))

    def _np(self, t): return t.numpy().view(np.uint32)
    def _t(self, a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int32))
    def lde(self, ev): return self._t(self.O.coset_lde_batch(F, self._np(ev), LOG_BLOWUP, self.O.generator(F), True))
    def commit_rows(self, mats): return [self._t(l) for l in self.O.merkle_tree(self.hs, [self._np(m) for m in mats])]

    def tree_from_digests(self, digests):
        lay = self._np(digests); out = [self._t(lay)]
        while lay.shape[0] > 1:
            lay = np.array([self.O.compress(self.hs, lay[2 * i], lay[2 * i + 1]) for i in range(lay.shape[0] // 2)])
            out.append(self._t(lay))
        return out


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:  # surface the failure instead of hanging the parent on the queue
        q.put((rank, False, False, repr(e)))


def _worker_body(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import p3_oracle as O
    from plonky3_b200.distributed import column_block, commit_bit_exact, commit_column_blocks
    be = OracleBackend()
    full = O.random_matrix(F, 1 << LOG_H, W, seed=3)                  # every rank derives the same trace, keeps its block
    widths = [column_block(W, world, g)[1] - column_block(W, world, g)[0] for g in range(world)]
    c0, c1 = column_block(W, world, rank)
    local = be._t(full[:, c0:c1])
    cap, recv, layers = commit_bit_exact(be, local, widths, CAP_H)
    roots, _, _ = commit_column_blocks(be, local)
    # reference: the whole trace committed in one piece (TwoAdicFriPcs::commit semantics)
    lde_full = O.coset_lde_batch(F, full, LOG_BLOWUP, O.generator(F), True)
    ol = O.merkle_tree(be.hs, [lde_full])
    ok_cap = np.array_equal(cap.numpy().view(np.uint32), O.merkle_cap(ol, CAP_H))
    rows = lde_full.shape[0] // world
    ok_rows = np.array_equal(np.concatenate([r.numpy().view(np.uint32) for r in recv], axis=1), lde_full[rank * rows:(rank + 1) * rows])
    exp_roots = [O.merkle_tree(be.hs, [O.coset_lde_batch(F, full[:, column_block(W, world, g)[0]:column_block(W, world, g)[1]], LOG_BLOWUP,
                                                       O.generator(F), True)])[-1][0] for g in range(world)]
    ok_blocks = np.array_equal(roots.numpy().view(np.uint32), np.array(exp_roots))
    q.put((rank, bool(ok_cap), bool(ok_rows), bool(ok_blocks)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_commit_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    for rank, ok_cap, ok_rows, ok_blocks in res:
        assert ok_cap, f"rank {rank}: bit-exact cap mismatch"
        assert ok_rows, f"rank {rank}: all-to-all row blocks mismatch"
        assert ok_blocks is True, f"rank {rank}: per-block roots mismatch / worker error: {ok_blocks}"


def test_column_block_partition():
    from plonky3_b200.distributed import column_block
    for w in (1, 7, 8, 100, 1312):
        for g in (1, 2, 4, 8):
            blocks = [column_block(w, g, r) for r in range(g)]
            assert blocks[0][0] == 0 and blocks[-1][1] == w
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(g - 1))
            assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1


def test_column_starts_and_chunk_bounds():
    """Host-side layout arithmetic of the peer-memory commit (no GPU call): column blocks tile the width, chunk bounds tile a
    block in multiples of 8 columns, and hashing the chunk matrices in column order is hashing the dense row
    (merkle_tree.rs:312-316 — the rule the chunk-major row blocks of p3gpu_commit_sharded_dev rely on)."""
    import ctypes as C
    from oracle import p3_oracle as O
    from plonky3_b200 import _lib
    from plonky3_b200.distributed import column_block, column_starts
    L = _lib.load()
    for width, world, align in [(100, 8, 8), (1312, 2, 8), (40, 8, 8), (11, 2, 1), (8, 4, 8)]:
        st = column_starts(width, world, align)
        assert st[0] == 0 and st[-1] == width and all(a <= b for a, b in zip(st, st[1:]))
        assert [column_block(width, world, r, align) for r in range(world)] == list(zip(st[:-1], st[1:]))
    for w_local in [0, 1, 7, 8, 64, 65, 100, 164, 656, 1312]:
        buf = (C.c_size_t * (w_local // 8 + 3))()
        n = L.p3gpu_shard_chunk_bounds(w_local, buf, len(buf))
        b = [int(buf[i]) for i in range(n)]
        assert n == (2 if 0 < w_local < 96 else len(b)) and b[0] == 0 and b[-1] == w_local and all(x < y for x, y in zip(b, b[1:]))   # w_local == 0: [0], no chunk
        assert all(x % 8 == 0 for x in b[:-1])
    be = OracleBackend()
    m = O.random_matrix(F, 64, 164, seed=9)
    buf = (C.c_size_t * 32)()
    n = L.p3gpu_shard_chunk_bounds(164, buf, 32)
    pieces = [np.ascontiguousarray(m[:, buf[i]:buf[i + 1]]) for i in range(n - 1) if buf[i + 1] > buf[i]]
    dense, chunked = O.merkle_tree(be.hs, [m]), O.merkle_tree(be.hs, pieces)
    assert all(np.array_equal(a, b) for a, b in zip(dense, chunked))
