"""Multi-GPU commit of one trace across the ranks of a torch.distributed group (one process per GPU, NCCL over NVLink).

SURVEY.md section 8(e): the NTT/LDE shards by COLUMN (every column is an independent polynomial, dft/src/traits.rs:22-24),
the Merkle tree shards by ROW RANGE (a leaf digest is a sequential sponge over the whole row, merkle_tree.rs:309-317, and
rows [k*H/G, (k+1)*H/G) of the bit-reversed LDE form a complete sub-tree).  Two modes:

  commit_column_blocks   BASELINE's "independent NTT+Merkle per shard": every rank commits its own column block;
                         ONE all-gather of the G roots.  Each root equals the reference's commitment to that column block
                         alone (G commitments, not the reference's single-trace commitment).
  commit_bit_exact       the reference's single commitment: column-sharded LDE -> ONE all-to-all that re-shards column
                         blocks into row blocks -> local leaf hashing + sub-tree -> ONE all-gather of sub-tree roots ->
                         the top log2(G) levels are compressed redundantly on every rank.  Bit-identical to
                         TwoAdicFriPcs::commit on the whole trace.  With G = 2^cap_height (8 GPUs, cap_height 3 as in
                         examples/src/proofs.rs:150) the gathered sub-tree roots ARE the Merkle cap.

The compute backend is injected (`backend.lde`, `backend.commit_rows`, `backend.tree_from_digests`), so the sharding and
collective logic is testable with gloo on CPU; `GpuBackend` binds it to libp3gpu.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def column_starts(width: int, world: int, align: int = 1):
    """The world + 1 offsets of the `column_block` split."""
    return [column_block(width, world, r, align)[0] for r in range(world)] + [width]


def column_block(width: int, world: int, rank: int, align: int = 1):
    """Contiguous column block of `rank` in units of `align` columns: the first (width/align) % world ranks get one extra
    unit, the last rank also takes the width % align remainder.  align = 8 keeps every 32-byte segment the LDE's last pass
    stores into a row block sector-aligned."""
    units, rem = divmod(width, align)
    base, extra = divmod(units, world)
    start = (rank * base + min(rank, extra)) * align
    stop = start + (base + (1 if rank < extra else 0)) * align
    if rank == world - 1:
        stop += rem
    return start, stop


class GpuBackend:
    """libp3gpu-backed compute (device-resident CUDA int32 tensors)."""

    def __init__(self, gpu, field, hash_kind, log_blowup):
        self.gpu, self.field, self.hash_kind, self.log_blowup = gpu, field, hash_kind, log_blowup

    def lde(self, evals):                       # (h, w_local) -> (h << log_blowup, w_local), bit-reversed rows
        return self.gpu.coset_lde_batch(self.field.id, evals, self.log_blowup, self.field.generator, bitrev_rows=True)

    def commit_rows(self, mats):                # list of same-height matrices (one per source rank) -> digest layers
        return self.gpu.merkle_commit(self.field.id, self.hash_kind, mats)

    def tree_from_digests(self, digests):       # (n, 8) -> layers above
        return self.gpu.merkle_from_digests(self.field.id, self.hash_kind, digests)


def _all_gather(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    world = dist.get_world_size(group)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous(), group=group)
    return out


def _all_to_all(recv: List[torch.Tensor], send: List[torch.Tensor], group=None):
    """NCCL: one all_to_all.  Backends without it (gloo, used by the CPU tests): pairwise isend/irecv."""
    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv, send, group=group)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    recv[rank].copy_(send[rank])
    reqs = []
    for k in range(world):
        if k != rank:
            reqs.append(dist.isend(send[k], dst=dist.get_global_rank(group, k) if group is not None else k, group=group))
            reqs.append(dist.irecv(recv[k], src=dist.get_global_rank(group, k) if group is not None else k, group=group))
    for r in reqs:
        r.wait()


def commit_column_blocks(backend, evals_local: torch.Tensor, group=None):
    """Independent LDE + Merkle per column block; one all-gather of roots.  Returns (roots (G, 8), lde_local, layers_local)."""
    lde = backend.lde(evals_local)
    layers = backend.commit_rows([lde])
    roots = torch.stack(_all_gather(layers[-1][0].contiguous(), group))
    return roots, lde, layers


def commit_bit_exact(backend, evals_local: torch.Tensor, widths: List[int], cap_height: int, group=None):
    """Bit-exact single commitment of the column-sharded trace.  `widths[g]` = columns held by rank g.
    Returns (cap (2^cap_height-ish, 8), lde_rows: list of per-source-rank row-block matrices, local sub-tree layers)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert world & (world - 1) == 0, "row sharding needs a power-of-two number of ranks"
    lde = backend.lde(evals_local)                                   # (H, w_rank), bit-reversed rows
    H = lde.shape[0]
    assert H % world == 0
    rows = H // world
    # all-to-all: rank g sends rows [k*rows, (k+1)*rows) of its column block to rank k
    send = [lde[k * rows:(k + 1) * rows].contiguous() for k in range(world)]
    recv = [torch.empty((rows, widths[g]), dtype=lde.dtype, device=lde.device) for g in range(world)]
    _all_to_all(recv, send, group)
    # the G received pieces, in rank order, are exactly the column blocks of my rows: hashing their row-wise concatenation
    # is hashing the full-width row (MerkleTree::new with several matrices of one height, merkle_tree.rs:312-316)
    layers = backend.commit_rows(recv)
    sub_root = layers[-1][0].contiguous()
    roots = torch.stack(_all_gather(sub_root, group))                # (G, 8): level log2(G) of the global tree
    top = backend.tree_from_digests(roots)                           # [roots, ..., global root]
    log_g = world.bit_length() - 1
    if cap_height <= log_g:
        cap = top[log_g - cap_height][: 1 << cap_height]
    else:                                                            # cap below the sub-tree roots: gather my slice of it
        h_local = cap_height - log_g
        eff = min(h_local, len(layers) - 1)
        piece = layers[len(layers) - 1 - eff][: 1 << eff].contiguous()
        cap = torch.cat(_all_gather(piece, group))
    return cap, recv, layers


# ------------------------------------------------------------------------------------------------------------------
# Peer-memory mode: no collective library on the data path (csrc/peer.cu, include/p3gpu.h "multi-GPU")
# ------------------------------------------------------------------------------------------------------------------
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


class RawBuffer:
    """Device memory from p3gpu_malloc (plain cudaMalloc: exportable through CUDA IPC, unlike a slice of torch's caching
    allocator).  `tensor(shape)` views it as a CUDA int32 tensor without copying."""

    def __init__(self, gpu, nbytes: int):
        self.gpu, self.nbytes = gpu, int(nbytes)
        p = C.c_void_p()
        check(gpu.L.p3gpu_malloc(gpu.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def tensor(self, shape):
        n = int(np.prod(shape))
        assert n * 4 <= self.nbytes
        holder = type("_CudaArray", (), {})()
        holder.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<i4", "data": (self.ptr, False), "version": 3}
        holder._keepalive = self
        return torch.as_tensor(holder, device=f"cuda:{self.gpu.device}")

    def free(self):
        if self.ptr:
            check(self.gpu.L.p3gpu_free(self.gpu.h, C.c_void_p(self.ptr)))
            self.ptr = None


class PeerGroup:
    """One rank's view of the group: its control block and row block plus every peer's, mapped through CUDA IPC.

    rows_per_rank x w_total is the shape of the row block each rank hashes (rows_per_rank = LDE height / world).
    Bootstrap = one all_gather_object of two 64-byte IPC handles per rank over torch.distributed (host side, once);
    afterwards no collective library call is made on the data path."""

    def __init__(self, gpu, rows_per_rank: int, w_total: int, group=None, timeout_s: float = 20.0, _sim=None):
        self.gpu, self.rows_per_rank, self.w_total = gpu, int(rows_per_rank), int(w_total)
        self.epoch = C.c_uint32(0)
        self._imported = []
        if _sim is not None:                       # single-process simulation (tests): all blocks live on this device
            self.world, self.rank, self.ctrl, self.rows, peers = _sim
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
            self.ctrl = RawBuffer(gpu, _lib.PEER_CTRL_BYTES)
            self.rows = RawBuffer(gpu, self.rows_per_rank * self.w_total * 4)
            check(gpu.L.p3gpu_ctx_use_own_stream(gpu.h))
            check(gpu.L.p3gpu_memset_dev(gpu.h, C.c_void_p(self.ctrl.ptr), 0, _lib.PEER_CTRL_BYTES))
            gpu.sync()
            peers = [(self.ctrl.ptr, self.rows.ptr)] * self.world
            if self.world > 1:
                hc, hr = (C.c_uint8 * 64)(), (C.c_uint8 * 64)()
                check(gpu.L.p3gpu_ipc_export(gpu.h, C.c_void_p(self.ctrl.ptr), hc))
                check(gpu.L.p3gpu_ipc_export(gpu.h, C.c_void_p(self.rows.ptr), hr))
                handles = [None] * self.world
                dist.all_gather_object(handles, (bytes(hc), bytes(hr)), group=group)
                peers = []
                for q, (bc, br) in enumerate(handles):
                    if q == self.rank:
                        peers.append((self.ctrl.ptr, self.rows.ptr))
                        continue
                    pc, pr = C.c_void_p(), C.c_void_p()
                    check(gpu.L.p3gpu_ipc_import(gpu.h, (C.c_uint8 * 64).from_buffer_copy(bc), C.byref(pc)))
                    check(gpu.L.p3gpu_ipc_import(gpu.h, (C.c_uint8 * 64).from_buffer_copy(br), C.byref(pr)))
                    self._imported += [pc.value, pr.value]
                    peers.append((pc.value, pr.value))
                dist.barrier(group=group)          # every control block is zeroed and mapped before the first device barrier
        self.struct = _lib.PeerGroupStruct()
        self.struct.world, self.struct.rank, self.struct.timeout_s = self.world, self.rank, timeout_s
        for q, (pc, pr) in enumerate(peers):
            self.struct.ctrl[q], self.struct.rows[q] = pc, pr

    @classmethod
    def simulate(cls, gpus, rows_per_rank: int, w_total: int, timeout_s: float = 10.0):
        """`len(gpus)` ranks inside ONE process on ONE device (one libp3gpu context = one stream per rank): lets a 1-GPU box
        exercise the sharded kernels, the barrier and the all-gather bit for bit.  Ranks must be driven from separate host
        threads, each under its own torch.cuda.stream (the barrier kernel spins until every rank has arrived: two ranks on
        one stream would dead-lock until the watchdog fires)."""
        world = len(gpus)
        ctrl = [RawBuffer(g, _lib.PEER_CTRL_BYTES) for g in gpus]
        rows = [RawBuffer(g, rows_per_rank * w_total * 4) for g in gpus]
        for g, c in zip(gpus, ctrl):
            check(g.L.p3gpu_ctx_use_own_stream(g.h))
            check(g.L.p3gpu_memset_dev(g.h, C.c_void_p(c.ptr), 0, _lib.PEER_CTRL_BYTES))
            g.sync()
        peers = [(c.ptr, r.ptr) for c, r in zip(ctrl, rows)]
        return [cls(g, rows_per_rank, w_total, timeout_s=timeout_s, _sim=(world, q, ctrl[q], rows[q], peers)) for q, g in enumerate(gpus)]

    def rows_tensor(self):
        return self.rows.tensor((self.rows_per_rank, self.w_total))

    def barrier(self):
        self.gpu._use_torch_stream()
        self.epoch.value += 1
        check(self.gpu.L.p3gpu_peer_barrier_dev(self.gpu.h, C.byref(self.struct), self.epoch.value))

    def lde_sharded(self, field, evals_local, added_bits: int, shift: int, col_off: int):
        """p3gpu_coset_lde_batch_sharded_dev on torch's current stream (complete on all ranks after the next barrier)."""
        m = self.gpu._dev(evals_local); self.gpu._use_torch_stream()
        check(self.gpu.L.p3gpu_coset_lde_batch_sharded_dev(self.gpu.h, field.id, C.byref(self.struct), m.data_ptr(), m.shape[0], m.shape[1],
                                                          added_bits, shift, self.w_total, col_off))

    def commit(self, field, hash_kind: int, evals_local, col_starts, log_blowup: int, cap_height: int, phases: bool = False):
        """p3gpu_commit_sharded_dev.  `col_starts`: world + 1 column offsets (rank g holds [col_starts[g], col_starts[g+1]));
        an int is taken as my own offset of an even `column_block` split, for callers that do not know the others' blocks.
        Returns (cap (n, 8) uint32 array — identical on every rank, my sub-tree's digest layers as CUDA tensors,
        [lde_ms, barrier_ms, hash_ms, cap_exchange_ms] or None).  With world > 1 my row block is left chunk-major
        (`row_block_dense` reassembles it)."""
        gpu = self.gpu
        m = gpu._dev(evals_local)                                    # (h, w_local); w_local may be 0 (more ranks than column units)
        h, w_local = int(m.shape[0]), int(m.shape[1])
        assert (h << log_blowup) == self.rows_per_rank * self.world
        if isinstance(col_starts, int):
            raise TypeError("PeerGroup.commit needs the world + 1 column offsets of all ranks (column_starts(...))")
        starts = [int(x) for x in col_starts]
        assert len(starts) == self.world + 1 and starts[-1] == self.w_total and starts[self.rank + 1] - starts[self.rank] == w_local
        self.col_starts = starts
        cs = (C.c_size_t * (self.world + 1))(*starts)
        gpu._use_torch_stream()
        tot = gpu.merkle_total_digests(self.rows_per_rank)
        layers = gpu._empty((tot, 8))
        lens = (C.c_size_t * 65)(); nl = C.c_size_t()
        cap = np.zeros((max(1 << cap_height, self.world), 8), dtype=np.uint32); cap_len = C.c_size_t()
        ph = (C.c_float * 4)() if phases else None
        check(gpu.L.p3gpu_commit_sharded_dev(gpu.h, field.id, hash_kind, C.byref(self.struct), C.byref(self.epoch), m.data_ptr(), h, cs,
                                             log_blowup, cap_height, layers.data_ptr(), lens, C.byref(nl),
                                             cap.ctypes.data, C.byref(cap_len), ph))
        out, off = [], 0
        for k in range(nl.value):
            out.append(layers[off:off + lens[k]]); off += lens[k]
        return cap[: cap_len.value].copy(), out, ([float(x) for x in ph] if phases else None)

    def chunk_bounds(self, w_local: int):
        b = (C.c_size_t * (w_local // 8 + 3))()
        n = self.gpu.L.p3gpu_shard_chunk_bounds(w_local, b, len(b))
        return [int(b[i]) for i in range(n)]

    def row_block_dense(self, col_starts=None):
        """My row block as a dense (rows_per_rank, w_total) tensor (a copy), whatever layout the last commit left it in."""
        starts = col_starts if col_starts is not None else self.col_starts
        R = self.rows_per_rank
        if self.world == 1:
            return self.rows_tensor().clone()
        flat = self.rows_tensor().reshape(-1)
        out = torch.empty((R, self.w_total), dtype=flat.dtype, device=flat.device)
        for g in range(self.world):
            cb = self.chunk_bounds(starts[g + 1] - starts[g])
            for a, b in zip(cb[:-1], cb[1:]):
                if b > a:
                    c0 = starts[g] + a
                    out[:, c0:c0 + (b - a)] = flat[R * c0: R * (c0 + b - a)].reshape(R, b - a)
        return out

    def close(self):
        for p in self._imported:
            self.gpu.L.p3gpu_ipc_close(self.gpu.h, C.c_void_p(p))
        self._imported = []
