// p3gpu.hpp — header-only C++ host mirror of the reference's trait surfaces over the C ABI in p3gpu.h.
//
// The reference's host code is Rust (no toolchain in this image); this is the C++ equivalent of the shim in
// same names, argument meaning and error behaviour (prover-side shape errors throw, where the
// reference panics).  Matrices are caller-owned row-major uint32_t buffers in Montgomery form.
//   p3gpu::Radix2DitParallel  ~ TwoAdicSubgroupDft          dft/src/traits.rs:28-291, radix_2_dit_parallel.rs:144-246
//   p3gpu::MerkleTreeMmcs     ~ Mmcs::commit                merkle-tree/src/mmcs/batch.rs:42-64
//   p3gpu::TwoAdicFriFolding  ~ FriFoldingStrategy          fri/src/two_adic_pcs.rs:134-213
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "p3gpu.h"

namespace p3gpu {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int32_t rc) {
    if (rc != P3GPU_OK) throw Error("p3gpu error " + std::to_string(rc) + ": " + p3gpu_last_error());
}

class Context {
  public:
    explicit Context(int device = 0) { check(p3gpu_ctx_create(device, &ctx_)); }
    ~Context() { p3gpu_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    p3gpu_ctx *raw() const { return ctx_; }
  private:
    p3gpu_ctx *ctx_ = nullptr;
};

struct RowMajorMatrix {                 // matrix/src/dense.rs:23-36
    std::vector<uint32_t> values;
    size_t width = 0;
    size_t height() const { return width ? values.size() / width : 0; }
};

// Result of the batch transforms: inner matrix holds the rows in bit-reversed order (matrix/src/bitrev.rs:82-101)
struct BitReversedMatrixView {
    RowMajorMatrix inner;
    RowMajorMatrix bit_reverse_rows() && { return std::move(inner); }   // zero data movement, as in the reference
};

class Radix2DitParallel {
  public:
    Radix2DitParallel(Context &c, int field) : c_(c), field_(field) {}
    RowMajorMatrix dft_batch(RowMajorMatrix m) const { return run(std::move(m), P3GPU_DFT, 0); }
    RowMajorMatrix idft_batch(RowMajorMatrix m) const { return run(std::move(m), P3GPU_IDFT, 0); }
    RowMajorMatrix coset_dft_batch(RowMajorMatrix m, uint32_t shift) const { return run(std::move(m), P3GPU_COSET_DFT, shift); }
    RowMajorMatrix coset_idft_batch(RowMajorMatrix m, uint32_t shift) const { return run(std::move(m), P3GPU_COSET_IDFT, shift); }
    BitReversedMatrixView coset_lde_batch(const RowMajorMatrix &m, unsigned added_bits, uint32_t shift) const {
        BitReversedMatrixView out;
        out.inner.width = m.width;
        out.inner.values.resize(m.values.size() << added_bits);
        check(p3gpu_coset_lde_batch(c_.raw(), field_, m.values.data(), m.height(), m.width, added_bits, shift,
                                    out.inner.values.data(), /*bitrev_rows=*/1));
        return out;
    }
  private:
    RowMajorMatrix run(RowMajorMatrix m, int kind, uint32_t shift) const {
        check(p3gpu_dft_batch(c_.raw(), field_, kind, m.values.data(), m.height(), m.width, shift));
        return m;
    }
    Context &c_;
    int field_;
};

struct MerkleTree {                     // merkle-tree/src/merkle_tree.rs:33-69 (arity schedule is all 2)
    std::vector<std::vector<uint32_t>> digest_layers;   // layer k: len_k * 8 words
    std::vector<uint32_t> cap(size_t cap_height) const {
        if (cap_height >= digest_layers.size()) throw Error("cap_height exceeds tree depth");
        const auto &l = digest_layers[digest_layers.size() - 1 - cap_height];
        size_t n = std::min<size_t>((size_t)1 << cap_height, l.size() / 8);
        return std::vector<uint32_t>(l.begin(), l.begin() + n * 8);
    }
};

class MerkleTreeMmcs {
  public:
    MerkleTreeMmcs(Context &c, int field, int hash, size_t cap_height) : c_(c), field_(field), hash_(hash), cap_height_(cap_height) {}
    // Mmcs::commit: returns (cap, tree)
    std::pair<std::vector<uint32_t>, MerkleTree> commit(const std::vector<const RowMajorMatrix *> &inputs) const {
        if (inputs.empty()) throw Error("No matrices given?");
        std::vector<const uint32_t *> ptrs; std::vector<size_t> hs, ws; size_t max_h = 0;
        for (auto *m : inputs) { ptrs.push_back(m->values.data()); hs.push_back(m->height()); ws.push_back(m->width); max_h = std::max(max_h, m->height()); }
        std::vector<uint32_t> flat(p3gpu_merkle_total_digests(max_h) * 8);
        size_t lens[65], n = 0;
        check(p3gpu_merkle_commit(c_.raw(), field_, hash_, ptrs.size(), ptrs.data(), hs.data(), ws.data(), flat.data(), lens, &n));
        MerkleTree t; size_t off = 0;
        for (size_t k = 0; k < n; k++) { t.digest_layers.emplace_back(flat.begin() + off * 8, flat.begin() + (off + lens[k]) * 8); off += lens[k]; }
        return {t.cap(std::min(cap_height_, n - 1)), std::move(t)};
    }
  private:
    Context &c_; int field_, hash_; size_t cap_height_;
};

class TwoAdicFriFolding {
  public:
    TwoAdicFriFolding(Context &c, int field) : c_(c), field_(field) {}
    // fold_matrix: m = rows x (arity * 4) words (EF4 values, bit-reversed evaluation order) -> rows x 4 words
    std::vector<uint32_t> fold_matrix(const uint32_t beta[4], unsigned log_arity, const std::vector<uint32_t> &m) const {
        const size_t rows = (m.size() / 4) >> log_arity;
        std::vector<uint32_t> out(rows * 4);
        check(p3gpu_fri_fold(c_.raw(), field_, m.data(), rows, log_arity, beta, out.data()));
        return out;
    }
  private:
    Context &c_; int field_;
};

}  // namespace p3gpu
