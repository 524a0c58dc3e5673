//! Plonky3 prover hot path on NVIDIA B200 through `libp3gpu` (`include/p3gpu.h`).
//!
//! | reference trait / type                                   | here                  | C entry points                          |
//! |----------------------------------------------------------|-----------------------|-----------------------------------------|
//! | `TwoAdicSubgroupDft` (`dft/src/traits.rs:28`)            | [`dft::GpuDft`]       | `p3gpu_dft_batch`, `p3gpu_coset_lde_batch` |
//! | `Mmcs` (`commit/src/mmcs.rs:19`), `MerkleTreeMmcs`       | [`mmcs::GpuMerkleMmcs`] | `p3gpu_merkle_commit`                 |
//! | `FriFoldingStrategy` (`fri/src/config.rs:147`)           | [`fri::GpuFriFolding`] | `p3gpu_fri_fold`                       |
//! | `Pcs` (`commit/src/pcs/univariate.rs:21`), `TwoAdicFriPcs` | [`pcs::GpuFriPcs`]  | `p3gpu_pcs_commit`, `p3gpu_*_dev`       |
//!
//! Field elements cross the boundary as the `u32` Montgomery representation that `MontyField31` already stores
//! (`monty-31/src/monty_31.rs:34-44`, `#[repr(transparent)]`), so matrices are passed by pointer without conversion.
#![allow(clippy::missing_safety_doc)]

pub mod dft;
pub mod ffi;
pub mod fri;
pub mod mmcs;
pub mod pcs;

use std::sync::Arc;

/// One `p3gpu_ctx` (one device, one stream, its scratch buffers and caches).  A context is re-entrant (its entry points serialise
/// on an internal mutex), so sharing an `Arc<GpuCtx>` between rayon workers is sound; clone a fresh context per worker to keep
/// several calls in flight.
pub struct GpuCtx(pub(crate) *mut ffi::P3GpuCtx);
unsafe impl Send for GpuCtx {}
unsafe impl Sync for GpuCtx {}

impl GpuCtx {
    pub fn new(device: i32) -> Arc<Self> {
        let mut raw = core::ptr::null_mut();
        ffi::check(unsafe { ffi::p3gpu_ctx_create(device, &mut raw) });
        Arc::new(Self(raw))
    }
    pub(crate) fn raw(&self) -> *mut ffi::P3GpuCtx {
        self.0
    }
}
impl Drop for GpuCtx {
    fn drop(&mut self) {
        unsafe { ffi::p3gpu_ctx_destroy(self.0) }
    }
}

/// The two fields the backend supports, with the identifier the C ABI uses.
pub trait GpuField: p3_field::TwoAdicField + p3_field::PrimeField32 {
    const GPU_ID: i32;
    /// The stored Montgomery word (what `MontyField31::value` holds).
    fn monty_word(self) -> u32;
}
impl GpuField for p3_baby_bear::BabyBear {
    const GPU_ID: i32 = ffi::P3GPU_BABY_BEAR;
    fn monty_word(self) -> u32 {
        // MontyField31 is repr(transparent) over its Montgomery u32
        unsafe { core::mem::transmute::<Self, u32>(self) }
    }
}
impl GpuField for p3_koala_bear::KoalaBear {
    const GPU_ID: i32 = ffi::P3GPU_KOALA_BEAR;
    fn monty_word(self) -> u32 {
        unsafe { core::mem::transmute::<Self, u32>(self) }
    }
}
