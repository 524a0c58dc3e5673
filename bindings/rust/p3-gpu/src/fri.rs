//! `FriFoldingStrategy::fold_matrix` on the GPU (`fri/src/two_adic_pcs.rs:134-213`); `fold_row` (verifier side) is the reference's.
use std::sync::Arc;

use p3_field::{BasedVectorSpace, ExtensionField};
use p3_fri::{FriFoldingStrategy, TwoAdicFriFolding};
use p3_matrix::Matrix;

use crate::ffi::*;
use crate::{GpuCtx, GpuField};

pub struct GpuFriFolding<F, InputProof, InputError> {
    pub(crate) ctx: Arc<GpuCtx>,
    pub(crate) reference: TwoAdicFriFolding<InputProof, InputError>,
    _f: core::marker::PhantomData<F>,
}

impl<F, EF, InputProof, InputError> FriFoldingStrategy<F, EF> for GpuFriFolding<F, InputProof, InputError>
where
    F: GpuField,
    EF: ExtensionField<F> + BasedVectorSpace<F>,
    TwoAdicFriFolding<InputProof, InputError>: FriFoldingStrategy<F, EF>,
{
    type InputProof = <TwoAdicFriFolding<InputProof, InputError> as FriFoldingStrategy<F, EF>>::InputProof;
    type InputError = <TwoAdicFriFolding<InputProof, InputError> as FriFoldingStrategy<F, EF>>::InputError;

    fn extra_query_index_bits(&self) -> usize {
        0
    }

    fn fold_row(&self, index: usize, log_height: usize, log_arity: usize, beta: EF, evals: impl Iterator<Item = EF>) -> EF {
        self.reference.fold_row(index, log_height, log_arity, beta, evals)
    }

    fn fold_matrix<M: Matrix<EF>>(&self, beta: EF, log_arity: usize, m: M) -> Vec<EF> {
        assert_eq!(EF::DIMENSION, 4, "the GPU fold is built for the quartic extension");
        // BinomialExtensionField<F, 4> is [F; 4] repr(transparent): a Vec<EF> of length n is an n x 4 base matrix
        let m = m.to_row_major_matrix();
        let rows = m.height();
        let mut out = EF::zero_vec(rows);
        check(unsafe {
            p3gpu_fri_fold(self.ctx.raw(), F::GPU_ID, m.values.as_ptr().cast(), rows, log_arity as u32,
                           beta.as_basis_coefficients_slice().as_ptr().cast(), out.as_mut_ptr().cast())
        });
        out
    }
}
