"""GPU implementation of the reference's TwoAdicSubgroupDft surface (dft/src/traits.rs:28-291) with the memory
conventions of Radix2DitParallel (dft/src/radix_2_dit_parallel.rs:144-246): batch results are returned as a
BitReversedMatrixView whose INNER matrix holds the rows in bit-reversed order, so that
`.bit_reverse_rows().to_row_major_matrix()` in TwoAdicFriPcs::commit (fri/src/two_adic_pcs.rs:315-318) is free.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .field import Field
from .gpu import Gpu, default_gpu, _is_torch


def _log2_strict(n: int) -> int:
    """p3_util::log2_strict_usize: panics on non powers of two."""
    if n <= 0 or n & (n - 1):
        raise ValueError(f"Not a power of two: {n}")
    return n.bit_length() - 1


def reverse_matrix_index_bits(mat):
    """matrix/src/util.rs:36-57 (host/device convenience; the GPU kernels never run this as a pass)."""
    h = mat.shape[0]
    bits = _log2_strict(h)
    idx = np.arange(h, dtype=np.int64)
    rev = np.zeros(h, dtype=np.int64)
    for b in range(bits):
        rev |= ((idx >> b) & 1) << (bits - 1 - b)
    if _is_torch(mat):
        import torch
        return mat[torch.from_numpy(rev).to(mat.device)].contiguous()
    return np.ascontiguousarray(mat[rev])


class BitReversedMatrixView:
    """matrix/src/bitrev.rs:82-101: logical row r is inner row bitrev(r)."""

    def __init__(self, inner):
        self.inner = inner

    def height(self): return self.inner.shape[0]
    def width(self): return self.inner.shape[1]

    def bit_reverse_rows(self):
        """Undo the view: returns the inner matrix (zero data movement, bitrev.rs:85-93)."""
        return self.inner

    def to_row_major_matrix(self):
        """Materialise the logical (natural) order."""
        return reverse_matrix_index_bits(self.inner)


class Radix2DitParallel:
    """Drop-in for p3_dft::Radix2DitParallel<F> running on the B200."""

    def __init__(self, field: Field, gpu: Gpu | None = None):
        self.field = field
        self.gpu = gpu or default_gpu()

    # traits.rs:62
    def dft_batch(self, mat):
        _log2_strict(mat.shape[0])
        return self.gpu.dft_batch(self.field.id, _lib.DFT, mat)

    # traits.rs:84
    def coset_dft_batch(self, mat, shift: int):
        _log2_strict(mat.shape[0])
        return self.gpu.dft_batch(self.field.id, _lib.COSET_DFT, mat, shift)

    # traits.rs:112
    def idft_batch(self, mat):
        _log2_strict(mat.shape[0])
        return self.gpu.dft_batch(self.field.id, _lib.IDFT, mat)

    # traits.rs:145
    def coset_idft_batch(self, mat, shift: int):
        _log2_strict(mat.shape[0])
        return self.gpu.dft_batch(self.field.id, _lib.COSET_IDFT, mat, shift)

    # traits.rs:199
    def lde_batch(self, mat, added_bits: int) -> BitReversedMatrixView:
        return self.coset_lde_batch(mat, added_bits, self.field.ONE)

    # traits.rs:227 / radix_2_dit_parallel.rs:181
    def coset_lde_batch(self, mat, added_bits: int, shift: int) -> BitReversedMatrixView:
        _log2_strict(mat.shape[0])
        return BitReversedMatrixView(self.gpu.coset_lde_batch(self.field.id, mat, added_bits, shift, bitrev_rows=True))

    # single-vector conveniences (traits.rs:46-50, 72-76, 100-102, ...)
    def dft(self, vec): return self.dft_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1)).ravel()
    def idft(self, vec): return self.idft_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1)).ravel()
    def coset_dft(self, vec, shift): return self.coset_dft_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1), shift).ravel()
    def coset_idft(self, vec, shift): return self.coset_idft_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1), shift).ravel()
    def lde(self, vec, added_bits): return self.lde_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1), added_bits).to_row_major_matrix().ravel()
    def coset_lde(self, vec, added_bits, shift):
        return self.coset_lde_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1), added_bits, shift).to_row_major_matrix().ravel()

    # *_algebra_* (traits.rs:279-291): EF4 columns are flattened to 4 base columns
    def dft_algebra_batch(self, mat_ef): return self.dft_batch(mat_ef.reshape(mat_ef.shape[0], -1)).reshape(mat_ef.shape)
    def idft_algebra_batch(self, mat_ef): return self.idft_batch(mat_ef.reshape(mat_ef.shape[0], -1)).reshape(mat_ef.shape)
