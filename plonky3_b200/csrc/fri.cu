// FRI folding on sm_100a: TwoAdicFriFolding::fold_matrix (fri/src/two_adic_pcs.rs:134-213).
//
// Input: a vector of EF4 = F[X]/(X^4 - W) evaluations in bit-reversed order viewed as `rows x arity`; output: `rows`
// folded values.  An arity-2^k fold is k chained arity-2 folds with beta, beta^2, beta^4, ...:
//     out = (lo + hi)/2 + (lo - hi) * beta * t_j ,   t_j = (1/2) * g_inv^bitrev(j)
// The reference rebuilds the table t per step (t_j <- 2*t_2j^2, :188-192).  Here ONE bit-reversed table per field serves
// every step and every round, because bit-reversed tables nest:  T[j] = (1/2) * prod_{bit b of j} g_(b+2)^-1  depends
// only on j (g_k = two-adic generator of order 2^k), and step s / level j uses T[j] for j < current height.
// One thread folds one output row entirely in registers (arity <= 8: 8 EF4 = 32 words), so the vector is read once
// and written once: the kernel is HBM-bound (16*arity bytes in, 16 bytes out per row).
#include "common.h"

namespace p3 {

struct FoldTableArgs { u32 ginv[32]; u32 half; };  // ginv[k] = two_adic_generator(k)^-1, Montgomery

template <int F> __global__ void gen_fold_table(u32 *T, size_t len, const FoldTableArgs a) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    u32 t = a.half;
    for (int b = 0; (j >> b) != 0; b++)
        if ((j >> b) & 1) t = mont_mul<F>(t, a.ginv[b + 2]);
    T[j] = t;
}

template <int F> __device__ __forceinline__ Ef4<F> ef_square(const Ef4<F> &a) { return ef_mul<F>(a, a); }

template <int F, int LOG_ARITY>
__global__ void __launch_bounds__(128) fri_fold_kernel(const u32 *in, u32 *out, size_t rows, const u32 *T, const Ef4<F> beta) {
    constexpr int A = 1 << LOG_ARITY;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    Ef4<F> v[A];
    const uint4 *src = reinterpret_cast<const uint4 *>(in + row * (size_t)A * 4);
#pragma unroll
    for (int k = 0; k < A; k++) {
        const uint4 q = __ldg(src + k);
        v[k].c[0] = q.x; v[k].c[1] = q.y; v[k].c[2] = q.z; v[k].c[3] = q.w;
    }
    Ef4<F> b = beta;
#pragma unroll
    for (int step = 0; step < LOG_ARITY; step++) {
        const int n = A >> (step + 1);  // outputs of this step within the row
#pragma unroll
        for (int l = 0; l < n; l++) {
            const u32 t = __ldg(T + row * (size_t)n + l);
            Ef4<F> d, o;
#pragma unroll
            for (int k = 0; k < 4; k++) d.c[k] = fp_sub<F>(v[2 * l].c[k], v[2 * l + 1].c[k]);
            const Ef4<F> db = ef_mul<F>(d, b);
#pragma unroll
            for (int k = 0; k < 4; k++)
                o.c[k] = fp_add<F>(fp_halve<F>(fp_add<F>(v[2 * l].c[k], v[2 * l + 1].c[k])), mont_mul<F>(db.c[k], t));
            v[l] = o;
        }
        if (step + 1 < LOG_ARITY) b = ef_square<F>(b);
    }
    *reinterpret_cast<uint4 *>(out + row * 4) = make_uint4(v[0].c[0], v[0].c[1], v[0].c[2], v[0].c[3]);
}

// folded[i] += s * x[i] over EF4 (commit_phase roll-in of a shorter input, fri/src/prover.rs:258-265: s = beta^arity)
template <int F> __global__ void __launch_bounds__(256) ef_axpy_kernel(u32 *acc, const u32 *x, size_t n, const Ef4<F> s) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 a = reinterpret_cast<const uint4 *>(acc)[i], b = __ldg(reinterpret_cast<const uint4 *>(x) + i);
    Ef4<F> xv; xv.c[0] = b.x; xv.c[1] = b.y; xv.c[2] = b.z; xv.c[3] = b.w;
    const Ef4<F> p = ef_mul<F>(s, xv);
    reinterpret_cast<uint4 *>(acc)[i] = make_uint4(fp_add<F>(a.x, p.c[0]), fp_add<F>(a.y, p.c[1]), fp_add<F>(a.z, p.c[2]), fp_add<F>(a.w, p.c[3]));
}

template <int F> static int32_t ensure_fold_table(p3gpu_ctx *ctx, size_t len) {
    if (ctx->fold_table_len[F] >= len) return P3GPU_OK;
    size_t cap = 1;
    while (cap < len) cap <<= 1;
    u32 *T = nullptr;
    P3_CUDA(cudaMalloc(&T, cap * 4));
    FoldTableArgs a;
    for (u32 k = 0; k < 32; k++) a.ginv[k] = k <= Fp<F>::TWO_ADICITY ? fp_inv<F>(two_adic_generator<F>(k)) : Fp<F>::ONE;
    a.half = fp_halve<F>(Fp<F>::ONE);
    gen_fold_table<F><<<(unsigned)((cap + 255) / 256), 256, 0, ctx->stream>>>(T, cap, a);
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    if (ctx->fold_table[F]) {
        P3_CUDA(cudaStreamSynchronize(ctx->stream));  // earlier folds may still read the old table
        cudaFree(ctx->fold_table[F]);
    }
    ctx->fold_table[F] = T;
    ctx->fold_table_len[F] = cap;
    return P3GPU_OK;
}

template <int F>
static int32_t fold_impl(p3gpu_ctx *ctx, const u32 *d_in, size_t rows, unsigned log_arity, const u32 beta[4], u32 *d_out) {
    // first step has rows * arity/2 outputs => table indices < rows << (log_arity - 1)
    P3_TRY(ensure_fold_table<F>(ctx, rows << (log_arity - 1)));
    Ef4<F> b;
    for (int k = 0; k < 4; k++) b.c[k] = beta[k];
    const unsigned g = (unsigned)((rows + 127) / 128);
    const u32 *T = ctx->fold_table[F];
    switch (log_arity) {
        case 1: fri_fold_kernel<F, 1><<<g, 128, 0, ctx->stream>>>(d_in, d_out, rows, T, b); break;
        case 2: fri_fold_kernel<F, 2><<<g, 128, 0, ctx->stream>>>(d_in, d_out, rows, T, b); break;
        case 3: fri_fold_kernel<F, 3><<<g, 128, 0, ctx->stream>>>(d_in, d_out, rows, T, b); break;
        default: fri_fold_kernel<F, 4><<<g, 128, 0, ctx->stream>>>(d_in, d_out, rows, T, b); break;
    }
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}

int32_t fri_fold(p3gpu_ctx *ctx, int field, const u32 *d_in, size_t rows, unsigned log_arity, const u32 beta[4], u32 *d_out) {
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    P3_CHECK(log_arity >= 1 && log_arity <= 4, P3GPU_EINVAL, "log_arity %u out of range 1..4", log_arity);
    P3_CHECK(is_pow2(rows), P3GPU_EINVAL, "fold: row count %zu is not a power of two", rows);
    const unsigned adicity = field == BABY_BEAR ? Fp<BABY_BEAR>::TWO_ADICITY : Fp<KOALA_BEAR>::TWO_ADICITY;
    P3_CHECK(log2_floor(rows) + log_arity <= adicity, P3GPU_EINVAL, "fold: vector longer than the two-adic subgroup");
    return field == BABY_BEAR ? fold_impl<BABY_BEAR>(ctx, d_in, rows, log_arity, beta, d_out)
                              : fold_impl<KOALA_BEAR>(ctx, d_in, rows, log_arity, beta, d_out);
}

int32_t fri_ef_axpy(p3gpu_ctx *ctx, int field, u32 *d_acc, const u32 *d_x, size_t n, const u32 s[4]) {
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    if (n == 0) return P3GPU_OK;
    const unsigned g = (unsigned)((n + 255) / 256);
    if (field == BABY_BEAR) { Ef4<BABY_BEAR> v; for (int k = 0; k < 4; k++) v.c[k] = s[k]; ef_axpy_kernel<BABY_BEAR><<<g, 256, 0, ctx->stream>>>(d_acc, d_x, n, v); }
    else { Ef4<KOALA_BEAR> v; for (int k = 0; k < 4; k++) v.c[k] = s[k]; ef_axpy_kernel<KOALA_BEAR><<<g, 256, 0, ctx->stream>>>(d_acc, d_x, n, v); }
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}

}  // namespace p3
