/*
 * p3gpu.h — C ABI of libp3gpu: the B200 (sm_100a) backend for Plonky3's prover hot path
 * (batched NTT / coset LDE  ->  Merkle-tree commitment  ->  FRI fold/commit loop).
 *
 * This is the drop-in boundary: the entry points are what a Rust FFI shim implementing the reference's
 * trait surfaces would bind (the shim is shown in INTEGRATION.md):
 *     TwoAdicSubgroupDft   dft/src/traits.rs:28-291                  -> p3gpu_dft_batch*, p3gpu_coset_lde_batch*
 *     Mmcs::commit         commit/src/mmcs.rs:42, merkle-tree/src/mmcs/batch.rs:42-64 -> p3gpu_merkle_commit*
 *     FriFoldingStrategy   fri/src/config.rs:147-169, two_adic_pcs.rs:134-213          -> p3gpu_fri_fold*
 *     Pcs::commit / commit_phase  fri/src/two_adic_pcs.rs:300-324, fri/src/prover.rs:192-286
 *                                                                   -> p3gpu_pcs_commit*, p3gpu_fri_commit_phase*
 *
 * Conventions
 *   - Field elements are uint32_t in MONTGOMERY form, bit-identical to MontyField31.value
 *     (monty-31/src/monty_31.rs:34-44), so RowMajorMatrix<F>.values.as_ptr() can be passed unchanged.
 *   - Matrices are row-major: element (r, c) at m[r * width + c] (matrix/src/dense.rs:23-33).
 *   - A digest is 8 x uint32_t (Poseidon2: [F; 8]; Keccak: [u64; 4] little-endian).
 *   - Every function returns 0 on success and a negative P3GPU_E* code otherwise; p3gpu_last_error()
 *     describes the most recent failure of the calling thread.  The reference's prover-side trait methods
 *     panic on shape violations (dft: log2_strict_usize; mmcs/batch.rs:50-54); the shim turns non-zero into panic!.
 *   - "_dev" variants take DEVICE pointers and run asynchronously on the context's stream; the plain variants
 *     take HOST pointers and include the host<->device copies (they synchronise before returning).
 *   - There is no CPU fallback: without a CUDA device p3gpu_ctx_create fails with P3GPU_ECUDA.
 *   - Threads and devices: a context belongs to one device and serialises its work on one stream; every entry point makes
 *     that device current for the calling host thread, so calls may come from any thread (the reference's DFT / MMCS
 *     objects are Clone + Sync).  One context must not be used by two threads at the same time; use one context per
 *     thread (each has its own stream, scratch buffers and twiddle cache) to keep several calls in flight.
 */
#ifndef P3GPU_H
#define P3GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct p3gpu_ctx p3gpu_ctx;

enum { P3GPU_BABY_BEAR = 0, P3GPU_KOALA_BEAR = 1 };

enum {
    P3GPU_OK = 0,
    P3GPU_EINVAL = -1,       /* bad shape (non power-of-two height, height above the field's two-adicity, ...) */
    P3GPU_EUNSUPPORTED = -2, /* unsupported field / width / hash */
    P3GPU_ECUDA = -3,        /* CUDA runtime error (incl. no device) */
    P3GPU_ENOMEM = -4,
    P3GPU_ESTATE = -5        /* missing configuration (e.g. Poseidon2 constants not set) */
};

/* which transform p3gpu_dft_batch computes (dft/src/traits.rs) */
enum {
    P3GPU_DFT = 0,        /* dft_batch            traits.rs:62    */
    P3GPU_IDFT = 1,       /* idft_batch           traits.rs:112   */
    P3GPU_COSET_DFT = 2,  /* coset_dft_batch      traits.rs:84    */
    P3GPU_COSET_IDFT = 3  /* coset_idft_batch     traits.rs:145   */
};

/* hash configurations of MerkleTreeMmcs (examples/src/types.rs:19-53, merkle-tree/benches/merkle_tree.rs:38) */
enum {
    P3GPU_HASH_POSEIDON2_W16 = 0, /* leaf PaddingFreeSponge<Perm16,16,8,8>,  node TruncatedPermutation<Perm16,2,8,16> */
    P3GPU_HASH_POSEIDON2_W24 = 1, /* leaf PaddingFreeSponge<Perm24,24,16,8>, node TruncatedPermutation<Perm16,2,8,16> */
    P3GPU_HASH_KECCAK = 2         /* leaf SerializingHasher<PaddingFreeSponge<KeccakF,25,17,4>>, node CompressionFunctionFromHasher<_,2,4> */
};

/* ---- context ---------------------------------------------------------------------------------- */
int32_t p3gpu_ctx_create(int device, p3gpu_ctx **out);
void p3gpu_ctx_destroy(p3gpu_ctx *ctx);
/* run subsequent calls on this cudaStream_t (e.g. torch's current stream); NULL = the legacy default stream.
 * A fresh context uses a private non-blocking stream; p3gpu_ctx_use_own_stream switches back to it. */
int32_t p3gpu_ctx_set_stream(p3gpu_ctx *ctx, void *cuda_stream);
int32_t p3gpu_ctx_use_own_stream(p3gpu_ctx *ctx);
int32_t p3gpu_ctx_sync(p3gpu_ctx *ctx);
const char *p3gpu_last_error(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t p3gpu_launch_count(const p3gpu_ctx *ctx);

/* device memory helpers for hosts that do not bring their own allocator */
int32_t p3gpu_malloc(p3gpu_ctx *ctx, size_t bytes, void **dptr);
int32_t p3gpu_free(p3gpu_ctx *ctx, void *dptr);
int32_t p3gpu_memcpy_h2d(p3gpu_ctx *ctx, void *dst, const void *src, size_t bytes);
int32_t p3gpu_memcpy_d2h(p3gpu_ctx *ctx, void *dst, const void *src, size_t bytes);
/* page-lock / unlock a host buffer so the host-pointer entry points copy at full PCIe rate */
int32_t p3gpu_host_register(void *ptr, size_t bytes);
int32_t p3gpu_host_unregister(void *ptr);

/* ---- TwoAdicSubgroupDft ----------------------------------------------------------------------- */
/* In-place capable (d_out may equal d_in).  kind: P3GPU_DFT..P3GPU_COSET_IDFT; shift (Montgomery) is used by the
 * coset kinds.  Result rows are in natural order (what `.to_row_major_matrix()` of the reference's result yields).
 * h must be a power of two <= 2^TWO_ADICITY; w >= 1. */
int32_t p3gpu_dft_batch_dev(p3gpu_ctx *ctx, int field, int kind, const uint32_t *d_in, uint32_t *d_out,
                            size_t h, size_t w, uint32_t shift);
int32_t p3gpu_dft_batch(p3gpu_ctx *ctx, int field, int kind, uint32_t *h_inout, size_t h, size_t w, uint32_t shift);

/* coset_lde_batch (traits.rs:227-234; Radix2DitParallel: radix_2_dit_parallel.rs:181-246).
 * in:  h x w evaluations over H (natural order).   out: (h << added_bits) x w evaluations over shift*K.
 * bitrev_rows != 0: memory row m holds the evaluation at shift * w_K^bitrev(m) — exactly the buffer
 *   Radix2DitParallel returns inside its BitReversedMatrixView and TwoAdicFriPcs::commit commits
 *   (fri/src/two_adic_pcs.rs:313-318).  bitrev_rows == 0: natural row order.
 * d_out must not alias d_in. */
int32_t p3gpu_coset_lde_batch_dev(p3gpu_ctx *ctx, int field, const uint32_t *d_in, size_t h, size_t w,
                                  unsigned added_bits, uint32_t shift, uint32_t *d_out, int bitrev_rows);
int32_t p3gpu_coset_lde_batch(p3gpu_ctx *ctx, int field, const uint32_t *h_in, size_t h, size_t w,
                              unsigned added_bits, uint32_t shift, uint32_t *h_out, int bitrev_rows);

/* ---- Poseidon2 / hashing ---------------------------------------------------------------------- */
/* Poseidon2::new (poseidon2/src/lib.rs:50-87): round constants cross the boundary in Montgomery form.
 * width 16 or 24; rc_initial / rc_terminal: 4 x width; rc_internal: rounds_p scalars. */
int32_t p3gpu_poseidon2_set_constants(p3gpu_ctx *ctx, int field, int width, const uint32_t *rc_initial,
                                      const uint32_t *rc_terminal, const uint32_t *rc_internal, int rounds_p);
/* Permutation::permute_mut on n independent states (n x width, device memory) — used by KAT tests/benches. */
int32_t p3gpu_poseidon2_permute_dev(p3gpu_ctx *ctx, int field, int width, uint32_t *d_states, size_t n);
/* Keccak-f[1600] on n independent states (n x 25 u64, device memory). */
int32_t p3gpu_keccak_f_dev(p3gpu_ctx *ctx, uint64_t *d_states, size_t n);

/* ---- Mmcs::commit ----------------------------------------------------------------------------- */
/* total digests in all layers of a tree whose tallest matrix has max_height rows (layers padded as the
 * reference pads them, merkle_tree.rs:473-481) */
size_t p3gpu_merkle_total_digests(size_t max_height);
/* MerkleTree::new with arity 2 over n_mats matrices (merkle_tree.rs:95-178; mixed heights allowed if they sit
 * on the reference's height ladder, mmcs/geometry.rs:83-124).  d_layers receives every digest layer
 * back to back (layer 0 = leaf digests); layer_lens[k] its length in digests; *n_layers the layer count
 * (layer_lens must have room for 65 entries).  The cap of height c is the first 2^c digests of layer n_layers-1-c. */
int32_t p3gpu_merkle_commit_dev(p3gpu_ctx *ctx, int field, int hash, size_t n_mats, const uint32_t *const *d_mats,
                                const size_t *heights, const size_t *widths, uint32_t *d_layers,
                                size_t *layer_lens, size_t *n_layers);
int32_t p3gpu_merkle_commit(p3gpu_ctx *ctx, int field, int hash, size_t n_mats, const uint32_t *const *h_mats,
                            const size_t *heights, const size_t *widths, uint32_t *h_layers,
                            size_t *layer_lens, size_t *n_layers);

/* Digest layers ABOVE an existing layer of n digests (d_digests, device): d_layers receives the (padded) copy of the input
 * layer followed by every layer up to the root, p3gpu_merkle_total_digests(n) digests in all.  Used to finish a tree whose
 * sub-tree roots were produced elsewhere (multi-GPU row sharding). */
int32_t p3gpu_merkle_from_digests_dev(p3gpu_ctx *ctx, int field, int hash, const uint32_t *d_digests, size_t n,
                                      uint32_t *d_layers, size_t *layer_lens, size_t *n_layers);

/* ---- FRI -------------------------------------------------------------------------------------- */
/* TwoAdicFriFolding::fold_matrix (two_adic_pcs.rs:134-213): rows x 2^log_arity EF4 values in bit-reversed
 * evaluation order -> rows EF4 values.  beta: 4 Montgomery words. */
int32_t p3gpu_fri_fold_dev(p3gpu_ctx *ctx, int field, const uint32_t *d_in, size_t rows, unsigned log_arity,
                           const uint32_t beta[4], uint32_t *d_out);
int32_t p3gpu_fri_fold(p3gpu_ctx *ctx, int field, const uint32_t *h_in, size_t rows, unsigned log_arity,
                       const uint32_t beta[4], uint32_t *h_out);

/* acc[i] += s * x[i] over EF4 (n elements, device memory): the roll-in of a shorter FRI input, folded += beta^arity * input
 * (fri/src/prover.rs:258-265). */
int32_t p3gpu_ef_axpy_dev(p3gpu_ctx *ctx, int field, uint32_t *d_acc, const uint32_t *d_x, size_t n, const uint32_t s[4]);

/* commit_phase (fri/src/prover.rs:192-286) for ONE input vector with caller-supplied betas (the Fiat-Shamir
 * transcript stays on the host; with commit_proof_of_work_bits = 0 a round's beta depends only on that round's cap,
 * so a host driving the transcript calls p3gpu_merkle_commit_dev / p3gpu_fri_fold_dev per round instead).
 * d_vec: len EF4 values (bit-reversed), consumed.  Rounds use compute_log_arity_for_round (fri/src/config.rs:180-207).
 * caps: per round 2^min(cap_height, layers-1) digests written back to back into h_caps (host), cap_lens[k] digests;
 * h_final: the folded vector of length 2^(log_blowup+log_final_poly_len) EF4 (before the final-poly iDFT). */
int32_t p3gpu_fri_commit_phase_dev(p3gpu_ctx *ctx, int field, int hash, uint32_t *d_vec, size_t len,
                                   unsigned log_blowup, unsigned log_final_poly_len, unsigned max_log_arity,
                                   unsigned cap_height, const uint32_t *betas /* rounds x 4 */, size_t n_betas,
                                   uint32_t *h_caps, size_t *cap_lens, unsigned *log_arities, size_t *n_rounds,
                                   uint32_t *h_final);

/* ---- Pcs::open, pre-FRI part (fri/src/two_adic_pcs.rs:413-662) ---------------- */
/* compute_inverse_denominators (:743-780): d_inv_denoms[i] = 1/(z - x_i) for x_i = GENERATOR * w^bitrev(i), i < 2^log_height
 * (EF4, bit-reversed coset order, so a prefix serves every smaller height).  If d_adjusted != NULL it receives
 * 1/(z - x_i) - 1/z (compute_adjusted_weights) and zinv = 1/z must be supplied. */
int32_t p3gpu_open_inv_denoms_dev(p3gpu_ctx *ctx, int field, unsigned log_height, const uint32_t z[4], const uint32_t *zinv,
                                  uint32_t *d_inv_denoms, uint32_t *d_adjusted);
/* Matrix::columnwise_dot_product: d_out[j] = scale * sum_i mat[i][j] * vec[i]  (vec: h EF4 values, out: w EF4 values; scale may be
 * NULL).  With vec = adjusted weights and scale = z (z^N - g^N) / (N g^N) this is interpolate_coset_with_precomputation
 * (matrix/src/interpolation.rs:161-193) on the first h rows of a committed bit-reversed LDE. */
int32_t p3gpu_columnwise_dot_dev(p3gpu_ctx *ctx, int field, const uint32_t *d_mat, size_t h, size_t w, const uint32_t *d_vec_ef,
                                 const uint32_t *scale, uint32_t *d_out);
/* rowwise_packed_dot_product with the powers of alpha (:622-626): d_out[i] = sum_j alpha^j * mat[i][j]  (h EF4 values). */
int32_t p3gpu_rowwise_dot_dev(p3gpu_ctx *ctx, int field, const uint32_t *d_mat, size_t h, size_t w, const uint32_t alpha[4],
                              uint32_t *d_out);
/* reduced-opening accumulation (:640-657): d_ro[i] += coeff * (yred - d_r[i]) * d_inv_denoms[i], i < h. */
int32_t p3gpu_open_reduce_dev(p3gpu_ctx *ctx, int field, uint32_t *d_ro, const uint32_t *d_r, const uint32_t *d_inv_denoms, size_t h,
                              const uint32_t coeff[4], const uint32_t yred[4]);

/* ---- Pcs::commit ------------------------------------------------------------------------------ */
/* TwoAdicFriPcs::commit for one matrix whose domain is the subgroup H (shift = GENERATOR / 1):
 * LDE onto GENERATOR*K with K = |H| << log_blowup, bit-reversed rows, then MerkleTreeMmcs::commit.
 * d_lde ((h<<log_blowup) x w) and d_layers stay resident for get_evaluations_on_domain / open. */
int32_t p3gpu_pcs_commit_dev(p3gpu_ctx *ctx, int field, int hash, const uint32_t *d_evals, size_t h, size_t w,
                             unsigned log_blowup, uint32_t *d_lde, uint32_t *d_layers, size_t *layer_lens,
                             size_t *n_layers);

#ifdef __cplusplus
}
#endif
#endif /* P3GPU_H */
