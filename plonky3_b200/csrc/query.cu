// Query phase gathers: Mmcs::open_batch for MANY indices at once on device-resident prover data
// (merkle-tree/src/mmcs/batch.rs:75-121 per index; fri/src/prover.rs:308-417 answer_queries / open_inputs call it for every FRI
// query).  The reference chases pointers on the host; with the matrices and digest layers in HBM the openings of all queries are
// two small gather kernels and one device-to-host copy each, instead of thousands of tiny copies.
#include "common.h"

namespace p3 {

// out[q][c] = mat[idx[q] >> shift][c]
__global__ void gather_rows_kernel(const u32 *mat, size_t w, const u32 *idx, unsigned shift, size_t n, u32 *out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * w) return;
    const size_t q = t / w, c = t - q * w;
    out[t] = __ldg(mat + (size_t)(idx[q] >> shift) * w + c);
}

struct PathArgs { size_t off[64]; int path_len; };   // off[l] = digest offset of layer l inside the layer buffer
// out[q][l] = layer_l[((idx[q] >> shift) >> l) ^ 1], l < path_len   (the sibling on the way up; batch.rs:103-118)
__global__ void merkle_paths_kernel(const u32 *layers, const u32 *idx, unsigned shift, size_t n, u32 *out, const PathArgs a) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * a.path_len * 2) return;
    const size_t half = t & 1, ql = t >> 1, q = ql / a.path_len, l = ql - q * a.path_len;
    const size_t node = ((size_t)(idx[q] >> shift) >> l) ^ 1;
    reinterpret_cast<uint4 *>(out)[t] = __ldg(reinterpret_cast<const uint4 *>(layers + (a.off[l] + node) * 8) + half);
}

static int32_t stage_indices(p3gpu_ctx *ctx, const u32 *h_idx, size_t n, u32 **d_idx) {
    void *p = nullptr;
    P3_TRY(ctx_leaf_table(ctx, n * 4 + 64, &p));
    P3_CUDA(cudaMemcpyAsync(p, h_idx, n * 4, cudaMemcpyHostToDevice, ctx->stream));   // pageable source: staged before return
    *d_idx = (u32 *)p;
    return P3GPU_OK;
}

int32_t query_gather_rows(p3gpu_ctx *ctx, const u32 *d_mat, size_t h, size_t w, const u32 *h_idx, size_t n, unsigned shift, u32 *d_out) {
    if (n == 0 || w == 0) return P3GPU_OK;
    for (size_t q = 0; q < n; q++) P3_CHECK((size_t)(h_idx[q] >> shift) < h, P3GPU_EINVAL, "index %u out of bounds for height %zu", h_idx[q] >> shift, h);
    u32 *d_idx;
    P3_TRY(stage_indices(ctx, h_idx, n, &d_idx));
    gather_rows_kernel<<<(unsigned)((n * w + 255) / 256), 256, 0, ctx->stream>>>(d_mat, w, d_idx, shift, n, d_out);
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}

int32_t query_merkle_paths(p3gpu_ctx *ctx, const u32 *d_layers, const size_t *layer_lens, size_t n_layers, size_t path_len, const u32 *h_idx,
                           size_t n, unsigned shift, u32 *d_out) {
    P3_CHECK(n_layers >= 1 && n_layers <= 64 && path_len < n_layers, P3GPU_EINVAL, "bad layer count / path length");
    if (n == 0 || path_len == 0) return P3GPU_OK;
    PathArgs a;
    size_t off = 0;
    for (size_t l = 0; l < n_layers; l++) { a.off[l] = off; off += layer_lens[l]; }
    a.path_len = (int)path_len;
    for (size_t q = 0; q < n; q++)
        for (size_t l = 0; l < path_len; l++)
            P3_CHECK((((size_t)(h_idx[q] >> shift) >> l) ^ 1) < layer_lens[l], P3GPU_EINVAL, "index %u out of bounds at layer %zu", h_idx[q] >> shift, l);
    u32 *d_idx;
    P3_TRY(stage_indices(ctx, h_idx, n, &d_idx));
    const size_t threads = n * path_len * 2;
    merkle_paths_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(d_layers, d_idx, shift, n, d_out, a);
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}

}  // namespace p3
