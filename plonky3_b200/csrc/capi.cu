// extern "C" surface of libp3gpu (include/p3gpu.h): context, memory plumbing, host-pointer wrappers and the
// PCS-level drivers (TwoAdicFriPcs::commit, fri commit phase) built from the NTT / hash / fold kernels.
#include <cstring>

#include "common.h"

namespace p3 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static int32_t grow(void **buf, size_t *cap, size_t bytes, cudaStream_t s) {
    if (*cap >= bytes) return P3GPU_OK;
    if (*buf) { P3_CUDA(cudaStreamSynchronize(s)); P3_CUDA(cudaFree(*buf)); *buf = nullptr; *cap = 0; }
    cudaError_t e = cudaMalloc(buf, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return P3GPU_ENOMEM; }
    *cap = bytes;
    return P3GPU_OK;
}
int32_t ctx_scratch(p3gpu_ctx *ctx, size_t bytes, void **out) {
    P3_TRY(grow(&ctx->scratch, &ctx->scratch_bytes, bytes, ctx->stream));
    *out = ctx->scratch;
    return P3GPU_OK;
}
int32_t ctx_scratch2(p3gpu_ctx *ctx, size_t bytes, void **out) {
    P3_TRY(grow(&ctx->scratch2, &ctx->scratch2_bytes, bytes, ctx->stream));
    *out = ctx->scratch2;
    return P3GPU_OK;
}

int32_t ctx_pool(p3gpu_ctx *ctx, int slot, size_t bytes, void **out) {
    P3_TRY(grow(&ctx->pool[slot], &ctx->pool_bytes[slot], bytes ? bytes : 1, ctx->stream));
    *out = ctx->pool[slot];
    return P3GPU_OK;
}

struct DevBuf {  // RAII device allocation for the host-pointer wrappers
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int32_t alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); p = nullptr; return P3GPU_ENOMEM; }
        return P3GPU_OK;
    }
};

}  // namespace p3

using namespace p3;

extern "C" {

const char *p3gpu_last_error(void) { return g_err; }

int32_t p3gpu_ctx_create(int device, p3gpu_ctx **out) {
    P3_CHECK(out != nullptr, P3GPU_EINVAL, "ctx_create: null output pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s): libp3gpu has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return P3GPU_ECUDA;
    }
    P3_CHECK(device >= 0 && device < n, P3GPU_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    P3_CUDA(cudaSetDevice(device));
    p3gpu_ctx *ctx = new p3gpu_ctx();
    ctx->device = device;
    memset(ctx->p2_host, 0, sizeof ctx->p2_host);
    cudaDeviceProp prop;
    P3_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    P3_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return P3GPU_OK;
}

void p3gpu_ctx_destroy(p3gpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->twiddles) cudaFree(kv.second);
    for (int f = 0; f < 2; f++) if (ctx->fold_table[f]) cudaFree(ctx->fold_table[f]);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch2) cudaFree(ctx->scratch2);
    for (int i = 0; i < 4; i++) if (ctx->pool[i]) cudaFree(ctx->pool[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int32_t p3gpu_ctx_set_stream(p3gpu_ctx *ctx, void *cuda_stream) {
    P3_CHECK(ctx, P3GPU_EINVAL, "null context");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    ctx->stream = (cudaStream_t)cuda_stream;  // NULL is the legacy default stream (what torch uses by default)
    return P3GPU_OK;
}
int32_t p3gpu_ctx_use_own_stream(p3gpu_ctx *ctx) {
    P3_CHECK(ctx, P3GPU_EINVAL, "null context");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    ctx->stream = ctx->own_stream;
    return P3GPU_OK;
}
int32_t p3gpu_ctx_sync(p3gpu_ctx *ctx) {
    P3_CHECK(ctx, P3GPU_EINVAL, "null context");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CUDA(cudaStreamSynchronize(ctx->stream));
    return P3GPU_OK;
}
uint64_t p3gpu_launch_count(const p3gpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int32_t p3gpu_malloc(p3gpu_ctx *ctx, size_t bytes, void **dptr) {
    P3_CHECK(ctx && dptr, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return P3GPU_ENOMEM; }
    return P3GPU_OK;
}
int32_t p3gpu_free(p3gpu_ctx *ctx, void *dptr) {
    P3_CHECK(ctx, P3GPU_EINVAL, "null context");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    if (dptr) { P3_CUDA(cudaStreamSynchronize(ctx->stream)); P3_CUDA(cudaFree(dptr)); }
    return P3GPU_OK;
}
int32_t p3gpu_memcpy_h2d(p3gpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    P3_CHECK(ctx, P3GPU_EINVAL, "null context");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return P3GPU_OK;
}
int32_t p3gpu_memcpy_d2h(p3gpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    P3_CHECK(ctx, P3GPU_EINVAL, "null context");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    P3_CUDA(cudaStreamSynchronize(ctx->stream));
    return P3GPU_OK;
}
int32_t p3gpu_host_register(void *ptr, size_t bytes) {
    P3_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return P3GPU_OK;
}
int32_t p3gpu_host_unregister(void *ptr) {
    P3_CUDA(cudaHostUnregister(ptr));
    return P3GPU_OK;
}

// ---- TwoAdicSubgroupDft ------------------------------------------------------------------------
int32_t p3gpu_dft_batch_dev(p3gpu_ctx *ctx, int field, int kind, const uint32_t *d_in, uint32_t *d_out, size_t h, size_t w,
                            uint32_t shift) {
    P3_CHECK(ctx && d_in && d_out, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return ntt_dft_batch(ctx, field, kind, d_in, d_out, h, w, shift);
}
int32_t p3gpu_dft_batch(p3gpu_ctx *ctx, int field, int kind, uint32_t *h_inout, size_t h, size_t w, uint32_t shift) {
    P3_CHECK(ctx && h_inout, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    void *buf = nullptr;
    P3_TRY(ctx_pool(ctx, 0, h * w * 4, &buf));
    P3_CUDA(cudaMemcpyAsync(buf, h_inout, h * w * 4, cudaMemcpyHostToDevice, ctx->stream));
    P3_TRY(ntt_dft_batch(ctx, field, kind, (const u32 *)buf, (u32 *)buf, h, w, shift));
    P3_CUDA(cudaMemcpyAsync(h_inout, buf, h * w * 4, cudaMemcpyDeviceToHost, ctx->stream));
    P3_CUDA(cudaStreamSynchronize(ctx->stream));
    return P3GPU_OK;
}
int32_t p3gpu_coset_lde_batch_dev(p3gpu_ctx *ctx, int field, const uint32_t *d_in, size_t h, size_t w, unsigned added_bits,
                                  uint32_t shift, uint32_t *d_out, int bitrev_rows) {
    P3_CHECK(ctx && d_in && d_out, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return ntt_coset_lde(ctx, field, d_in, h, w, added_bits, shift, d_out, bitrev_rows);
}
int32_t p3gpu_coset_lde_batch(p3gpu_ctx *ctx, int field, const uint32_t *h_in, size_t h, size_t w, unsigned added_bits,
                              uint32_t shift, uint32_t *h_out, int bitrev_rows) {
    P3_CHECK(ctx && h_in && h_out, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CHECK(added_bits <= 8, P3GPU_EINVAL, "added_bits %u too large", added_bits);
    void *in = nullptr, *out = nullptr;
    const size_t nin = h * w * 4, nout = nin << added_bits;
    P3_TRY(ctx_pool(ctx, 0, nin, &in));
    P3_TRY(ctx_pool(ctx, 1, nout, &out));
    P3_CUDA(cudaMemcpyAsync(in, h_in, nin, cudaMemcpyHostToDevice, ctx->stream));
    P3_TRY(ntt_coset_lde(ctx, field, (const u32 *)in, h, w, added_bits, shift, (u32 *)out, bitrev_rows));
    P3_CUDA(cudaMemcpyAsync(h_out, out, nout, cudaMemcpyDeviceToHost, ctx->stream));
    P3_CUDA(cudaStreamSynchronize(ctx->stream));
    return P3GPU_OK;
}

// ---- hashing -----------------------------------------------------------------------------------
int32_t p3gpu_poseidon2_set_constants(p3gpu_ctx *ctx, int field, int width, const uint32_t *rc_initial, const uint32_t *rc_terminal,
                                      const uint32_t *rc_internal, int rounds_p) {
    P3_CHECK(ctx && rc_initial && rc_terminal && rc_internal, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    P3_CHECK(width == 16 || width == 24, P3GPU_EUNSUPPORTED, "Poseidon2 width %d unsupported (16 or 24)", width);
    P3_CHECK(rounds_p >= 1 && rounds_p <= 32, P3GPU_EINVAL, "rounds_p %d out of range", rounds_p);
    const uint32_t p = field == BABY_BEAR ? Fp<BABY_BEAR>::P : Fp<KOALA_BEAR>::P;
    Poseidon2Consts &k = ctx->p2_host[field][width == 24];
    memset(&k, 0, sizeof k);
    for (int i = 0; i < 4 * width; i++) {
        P3_CHECK(rc_initial[i] < p && rc_terminal[i] < p, P3GPU_EINVAL, "round constant not in canonical Montgomery range");
        k.rc_ext[i] = rc_initial[i]; k.rc_ext[4 * width + i] = rc_terminal[i];
    }
    for (int i = 0; i < rounds_p; i++) {
        P3_CHECK(rc_internal[i] < p, P3GPU_EINVAL, "round constant not in canonical Montgomery range");
        k.rc_int[i] = rc_internal[i];
    }
    k.rounds_p = rounds_p; k.width = width; k.set = 1;
    return P3GPU_OK;
}
int32_t p3gpu_poseidon2_permute_dev(p3gpu_ctx *ctx, int field, int width, uint32_t *d_states, size_t n) {
    P3_CHECK(ctx && d_states, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return hash_poseidon2_permute(ctx, field, width, d_states, n);
}
int32_t p3gpu_keccak_f_dev(p3gpu_ctx *ctx, uint64_t *d_states, size_t n) {
    P3_CHECK(ctx && d_states, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return hash_keccak_f(ctx, d_states, n);
}

size_t p3gpu_merkle_total_digests(size_t max_height) {
    auto pad = [](size_t raw) { return raw <= 1 ? raw : (raw + 1) / 2 * 2; };
    size_t len = pad(max_height), tot = len;
    while (len > 1) { len = pad(len / 2); tot += len; }
    return tot;
}
int32_t p3gpu_merkle_commit_dev(p3gpu_ctx *ctx, int field, int hash, size_t n_mats, const uint32_t *const *d_mats,
                                const size_t *heights, const size_t *widths, uint32_t *d_layers, size_t *layer_lens,
                                size_t *n_layers) {
    P3_CHECK(ctx && d_mats && heights && widths && d_layers && layer_lens && n_layers, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return hash_merkle_commit(ctx, field, hash, n_mats, d_mats, heights, widths, d_layers, layer_lens, n_layers);
}
int32_t p3gpu_merkle_commit(p3gpu_ctx *ctx, int field, int hash, size_t n_mats, const uint32_t *const *h_mats,
                            const size_t *heights, const size_t *widths, uint32_t *h_layers, size_t *layer_lens,
                            size_t *n_layers) {
    P3_CHECK(ctx && h_mats && heights && widths && h_layers && layer_lens && n_layers, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CHECK(n_mats >= 1 && n_mats <= 1024, P3GPU_EINVAL, "No matrices given?");
    std::vector<DevBuf> bufs(n_mats);
    std::vector<const u32 *> ptrs(n_mats);
    size_t max_h = 0;
    for (size_t i = 0; i < n_mats; i++) {
        const size_t bytes = heights[i] * widths[i] * 4;
        P3_TRY(bufs[i].alloc(bytes));
        P3_CUDA(cudaMemcpyAsync(bufs[i].p, h_mats[i], bytes, cudaMemcpyHostToDevice, ctx->stream));
        ptrs[i] = (const u32 *)bufs[i].p;
        if (heights[i] > max_h) max_h = heights[i];
    }
    DevBuf layers;
    const size_t tot = p3gpu_merkle_total_digests(max_h);
    P3_TRY(layers.alloc(tot * 32));
    P3_TRY(hash_merkle_commit(ctx, field, hash, n_mats, ptrs.data(), heights, widths, (u32 *)layers.p, layer_lens, n_layers));
    P3_CUDA(cudaMemcpyAsync(h_layers, layers.p, tot * 32, cudaMemcpyDeviceToHost, ctx->stream));
    P3_CUDA(cudaStreamSynchronize(ctx->stream));
    return P3GPU_OK;
}

int32_t p3gpu_merkle_from_digests_dev(p3gpu_ctx *ctx, int field, int hash, const uint32_t *d_digests, size_t n, uint32_t *d_layers,
                                      size_t *layer_lens, size_t *n_layers) {
    P3_CHECK(ctx && d_digests && d_layers && layer_lens && n_layers, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return hash_merkle_from_digests(ctx, field, hash, d_digests, n, d_layers, layer_lens, n_layers);
}

// ---- FRI ---------------------------------------------------------------------------------------
int32_t p3gpu_fri_fold_dev(p3gpu_ctx *ctx, int field, const uint32_t *d_in, size_t rows, unsigned log_arity, const uint32_t beta[4],
                           uint32_t *d_out) {
    P3_CHECK(ctx && d_in && d_out && beta, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return fri_fold(ctx, field, d_in, rows, log_arity, beta, d_out);
}
int32_t p3gpu_fri_fold(p3gpu_ctx *ctx, int field, const uint32_t *h_in, size_t rows, unsigned log_arity, const uint32_t beta[4],
                       uint32_t *h_out) {
    P3_CHECK(ctx && h_in && h_out && beta, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CHECK(log_arity >= 1 && log_arity <= 4, P3GPU_EINVAL, "log_arity %u out of range 1..4", log_arity);
    void *in = nullptr, *out = nullptr;
    const size_t nin = (rows << log_arity) * 16;
    P3_TRY(ctx_pool(ctx, 0, nin, &in));
    P3_TRY(ctx_pool(ctx, 1, rows * 16, &out));
    P3_CUDA(cudaMemcpyAsync(in, h_in, nin, cudaMemcpyHostToDevice, ctx->stream));
    P3_TRY(fri_fold(ctx, field, (const u32 *)in, rows, log_arity, beta, (u32 *)out));
    P3_CUDA(cudaMemcpyAsync(h_out, out, rows * 16, cudaMemcpyDeviceToHost, ctx->stream));
    P3_CUDA(cudaStreamSynchronize(ctx->stream));
    return P3GPU_OK;
}

int32_t p3gpu_ef_axpy_dev(p3gpu_ctx *ctx, int field, uint32_t *d_acc, const uint32_t *d_x, size_t n, const uint32_t s[4]) {
    P3_CHECK(ctx && d_acc && d_x && s, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return fri_ef_axpy(ctx, field, d_acc, d_x, n, s);
}

// fri/src/config.rs:180-207 with a single input vector (next_input_log_height = None)
static unsigned log_arity_for_round(unsigned log_cur, unsigned log_final, unsigned max_log_arity) {
    const unsigned m = log_cur - log_final;
    return m < max_log_arity ? m : max_log_arity;
}

int32_t p3gpu_fri_commit_phase_dev(p3gpu_ctx *ctx, int field, int hash, uint32_t *d_vec, size_t len, unsigned log_blowup,
                                   unsigned log_final_poly_len, unsigned max_log_arity, unsigned cap_height, const uint32_t *betas,
                                   size_t n_betas, uint32_t *h_caps, size_t *cap_lens, unsigned *log_arities, size_t *n_rounds,
                                   uint32_t *h_final) {
    P3_CHECK(ctx && d_vec && betas && h_caps && cap_lens && log_arities && n_rounds && h_final, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CHECK(is_pow2(len), P3GPU_EINVAL, "commit phase: length %zu is not a power of two", len);
    P3_CHECK(max_log_arity >= 1 && max_log_arity <= 4, P3GPU_EINVAL, "max_log_arity must be in 1..4 to guarantee folding progress");
    const unsigned log_final = log_blowup + log_final_poly_len;
    // digest layers of the largest round + ping-pong buffer for the folded vector
    const unsigned la0 = log2_floor(len) > log_final ? log_arity_for_round(log2_floor(len), log_final, max_log_arity) : 1;
    void *layers = nullptr, *pong = nullptr;
    P3_TRY(ctx_pool(ctx, 2, p3gpu_merkle_total_digests(len >> la0) * 32, &layers));
    P3_TRY(ctx_pool(ctx, 3, (len >> la0) * 16 + 16, &pong));
    u32 *cur = d_vec, *other = (u32 *)pong;
    size_t cur_len = len, round = 0, cap_off = 0;
    while (cur_len > ((size_t)1 << log_final)) {
        P3_CHECK(round < n_betas, P3GPU_EINVAL, "commit phase: %zu betas supplied, more rounds needed", n_betas);
        const unsigned la = log_arity_for_round(log2_floor(cur_len), log_final, max_log_arity);
        const size_t rows = cur_len >> la, width = ((size_t)4) << la;  // ExtensionMmcs: EF4 -> 4 base columns
        const u32 *mats[1] = {cur};
        size_t lens[65], nl = 0;
        P3_TRY(hash_merkle_commit(ctx, field, hash, 1, mats, &rows, &width, (u32 *)layers, lens, &nl));
        // cap(min(cap_height, layers-1)): mmcs/batch.rs:56-62, merkle_tree.rs:198-217
        const size_t eff = cap_height < nl - 1 ? cap_height : nl - 1;
        size_t off = 0;
        for (size_t k = 0; k + 1 + eff < nl; k++) off += lens[k];
        const size_t cl = std::min((size_t)1 << eff, lens[nl - 1 - eff]);
        P3_CUDA(cudaMemcpyAsync(h_caps + cap_off * 8, (u32 *)layers + off * 8, cl * 32, cudaMemcpyDeviceToHost, ctx->stream));
        cap_lens[round] = cl; cap_off += cl; log_arities[round] = la;
        P3_TRY(fri_fold(ctx, field, cur, rows, la, betas + 4 * round, other));
        std::swap(cur, other);
        cur_len = rows; round++;
    }
    P3_CUDA(cudaMemcpyAsync(h_final, cur, cur_len * 16, cudaMemcpyDeviceToHost, ctx->stream));
    P3_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_rounds = round;
    return P3GPU_OK;
}

// ---- Pcs::open (pre-FRI part) ------------------------------------------------------------------
int32_t p3gpu_open_inv_denoms_dev(p3gpu_ctx *ctx, int field, unsigned log_height, const uint32_t z[4], const uint32_t *zinv,
                                  uint32_t *d_inv_denoms, uint32_t *d_adjusted) {
    P3_CHECK(ctx && z && d_inv_denoms, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return open_inv_denoms(ctx, field, log_height, z, zinv, d_inv_denoms, d_adjusted);
}
int32_t p3gpu_columnwise_dot_dev(p3gpu_ctx *ctx, int field, const uint32_t *d_mat, size_t h, size_t w, const uint32_t *d_vec_ef,
                                 const uint32_t *scale, uint32_t *d_out) {
    P3_CHECK(ctx && d_mat && d_vec_ef && d_out, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return open_columnwise_dot(ctx, field, d_mat, h, w, d_vec_ef, d_out, scale);
}
int32_t p3gpu_rowwise_dot_dev(p3gpu_ctx *ctx, int field, const uint32_t *d_mat, size_t h, size_t w, const uint32_t alpha[4], uint32_t *d_out) {
    P3_CHECK(ctx && d_mat && alpha && d_out, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return open_rowwise_dot(ctx, field, d_mat, h, w, alpha, d_out);
}
int32_t p3gpu_open_reduce_dev(p3gpu_ctx *ctx, int field, uint32_t *d_ro, const uint32_t *d_r, const uint32_t *d_inv_denoms, size_t h,
                              const uint32_t coeff[4], const uint32_t yred[4]) {
    P3_CHECK(ctx && d_ro && d_r && d_inv_denoms && coeff && yred, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    return open_reduce(ctx, field, d_ro, d_r, d_inv_denoms, h, coeff, yred);
}

// ---- Pcs::commit -------------------------------------------------------------------------------
int32_t p3gpu_pcs_commit_dev(p3gpu_ctx *ctx, int field, int hash, const uint32_t *d_evals, size_t h, size_t w, unsigned log_blowup,
                             uint32_t *d_lde, uint32_t *d_layers, size_t *layer_lens, size_t *n_layers) {
    P3_CHECK(ctx && d_evals && d_lde && d_layers && layer_lens && n_layers, P3GPU_EINVAL, "null argument");
    P3_CUDA(cudaSetDevice(ctx->device));   // CUDA's current device is per host thread: callers may come from any thread
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    // shift = GENERATOR / domain.shift() with domain.shift() = 1 (two_adic_pcs.rs:312)
    const u32 shift = field == BABY_BEAR ? to_monty<BABY_BEAR>(Fp<BABY_BEAR>::GEN) : to_monty<KOALA_BEAR>(Fp<KOALA_BEAR>::GEN);
    P3_TRY(ntt_coset_lde(ctx, field, d_evals, h, w, log_blowup, shift, d_lde, 1));
    const u32 *mats[1] = {d_lde};
    const size_t lh = h << log_blowup;
    return hash_merkle_commit(ctx, field, hash, 1, mats, &lh, &w, d_layers, layer_lens, n_layers);
}

}  // extern "C"
