//! `TwoAdicSubgroupDft` on the GPU: drop-in for `p3_dft::Radix2DitParallel` (`dft/src/radix_2_dit_parallel.rs:144-246`).
use core::marker::PhantomData;
use std::sync::Arc;

use p3_dft::TwoAdicSubgroupDft;
use p3_matrix::Matrix;
use p3_matrix::bitrev::{BitReversedMatrixView, BitReversibleMatrix};
use p3_matrix::dense::RowMajorMatrix;
use p3_matrix::util::reverse_matrix_index_bits;
use p3_util::log2_strict_usize;

use crate::ffi::*;
use crate::{GpuCtx, GpuField};

#[derive(Clone)]
pub struct GpuDft<F> {
    pub(crate) ctx: Arc<GpuCtx>,
    _f: PhantomData<F>,
}

impl<F> GpuDft<F> {
    pub fn new(ctx: Arc<GpuCtx>) -> Self {
        Self { ctx, _f: PhantomData }
    }
}
impl<F> Default for GpuDft<F> {
    fn default() -> Self {
        Self::new(GpuCtx::new(0))
    }
}

impl<F: GpuField> GpuDft<F> {
    fn transform(&self, kind: i32, mut mat: RowMajorMatrix<F>, shift: F) -> RowMajorMatrix<F> {
        let (h, w) = (mat.height(), mat.width());
        log2_strict_usize(h); // keep the reference's panic on non power-of-two heights
        check(unsafe { p3gpu_dft_batch(self.ctx.raw(), F::GPU_ID, kind, mat.values.as_mut_ptr().cast(), h, w, shift.monty_word()) });
        mat
    }
}

impl<F: GpuField> TwoAdicSubgroupDft<F> for GpuDft<F> {
    // same associated type as Radix2DitParallel (:146): the INNER matrix holds the rows in bit-reversed order
    type Evaluations = BitReversedMatrixView<RowMajorMatrix<F>>;

    fn dft_batch(&self, mat: RowMajorMatrix<F>) -> Self::Evaluations {
        // p3gpu_dft_batch returns natural order; wrap it so that the logical order is natural (radix_2_dit_parallel.rs:165)
        let mut out = self.transform(P3GPU_DFT, mat, F::ONE);
        reverse_matrix_index_bits(&mut out);
        out.bit_reverse_rows()
    }

    fn coset_dft_batch(&self, mat: RowMajorMatrix<F>, shift: F) -> Self::Evaluations {
        let mut out = self.transform(P3GPU_COSET_DFT, mat, shift);
        reverse_matrix_index_bits(&mut out);
        out.bit_reverse_rows()
    }

    fn idft_batch(&self, mat: RowMajorMatrix<F>) -> RowMajorMatrix<F> {
        self.transform(P3GPU_IDFT, mat, F::ONE)
    }

    fn coset_idft_batch(&self, mat: RowMajorMatrix<F>, shift: F) -> RowMajorMatrix<F> {
        self.transform(P3GPU_COSET_IDFT, mat, shift)
    }

    fn coset_lde_batch(&self, mat: RowMajorMatrix<F>, added_bits: usize, shift: F) -> Self::Evaluations {
        let (h, w) = (mat.height(), mat.width());
        log2_strict_usize(h);
        let mut out = F::zero_vec((h << added_bits) * w);
        // bitrev_rows = 1: the buffer Radix2DitParallel leaves in memory (radix_2_dit_parallel.rs:245) — commit()'s
        // `.bit_reverse_rows().to_row_major_matrix()` (fri/src/two_adic_pcs.rs:315-318) then costs nothing
        check(unsafe {
            p3gpu_coset_lde_batch(self.ctx.raw(), F::GPU_ID, mat.values.as_ptr().cast(), h, w, added_bits as u32, shift.monty_word(),
                                  out.as_mut_ptr().cast(), 1)
        });
        RowMajorMatrix::new(out, w).bit_reverse_rows_view()
    }
}

/// helper: view an already bit-reversed buffer through the bit-reversal (zero data movement, matrix/src/bitrev.rs:82-101)
trait BitRevView<F> {
    fn bit_reverse_rows_view(self) -> BitReversedMatrixView<RowMajorMatrix<F>>;
}
impl<F: Clone + Send + Sync> BitRevView<F> for RowMajorMatrix<F> {
    fn bit_reverse_rows_view(self) -> BitReversedMatrixView<RowMajorMatrix<F>> {
        p3_matrix::bitrev::BitReversalPerm::new_view(self)
    }
}
