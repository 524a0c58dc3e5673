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


def column_block(width: int, world: int, rank: int):
    """Contiguous column block of `rank`: the first width % world ranks get one extra column."""
    base, extra = divmod(width, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


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
