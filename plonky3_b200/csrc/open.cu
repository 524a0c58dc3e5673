// TwoAdicFriPcs::open — the pre-FRI work on the device (fri/src/two_adic_pcs.rs:413-662).
//
// After the commitments, `open` reads every committed LDE again: once to evaluate each column at the out-of-domain point(s)
// by barycentric interpolation (a column-wise dot product of the low coset with EF4 weights, matrix/src/interpolation.rs:161-193)
// and once to compress each row with powers of alpha (row-wise dot product) and accumulate the quotient
// (Mred(z) - Mred(x)) / (z - x) into the FRI input vector.  With the LDEs resident in HBM these are streaming reductions:
// 4 multiply-accumulates per matrix element (base x EF4).  They use LAZY 64-bit accumulation: acc += m * v as one IMAD.WIDE,
// followed by one VIADDMNMX on the high word that conditionally subtracts p*2^32 (invariant acc < p*2^32, so acc + m*v < 2^64
// never overflows); a single Montgomery reduction per accumulator happens at the end.  3 pipe slots per term instead of 9,
// which puts both kernels under the HBM roofline (16 B... 4 B read per 12 slots).
#include "common.h"

namespace p3 {

template <int F> __device__ __forceinline__ void lazy_mac(u64 &acc, u32 m, u32 v) {
    acc += (u64)m * v;
    u32 hi = (u32)(acc >> 32);
    const u32 hs = hi - Fp<F>::P;
    hi = hi < hs ? hi : hs;                       // unsigned min: subtract p*2^32 iff acc >= p*2^32
    acc = ((u64)hi << 32) | (u32)acc;
}
// value = acc * 2^-32 (Montgomery reduction); the accumulated products m*v with v in Montgomery form and m in Montgomery form
// therefore come out as Montgomery(m*v): exactly sum of mont_mul(m, v).
template <int F> __device__ __forceinline__ u32 lazy_finish(u64 acc) { return mont_redc<F>(acc); }

template <int F> struct EfInvArgs { u32 zeta; };  // W^((p-1)/4), Montgomery: Frobenius X -> zeta * X

template <int F> __device__ inline Ef4<F> ef_inv_dev(const Ef4<F> &a, u32 zeta) {
    const u32 a1z = mont_mul<F>(a.c[1], zeta), a3z = mont_mul<F>(a.c[3], zeta);
    Ef4<F> c1, c2, c3;
    c1.c[0] = a.c[0]; c1.c[1] = a1z;               c1.c[2] = fp_neg<F>(a.c[2]); c1.c[3] = fp_neg<F>(a3z);
    c2.c[0] = a.c[0]; c2.c[1] = fp_neg<F>(a.c[1]); c2.c[2] = a.c[2];            c2.c[3] = fp_neg<F>(a.c[3]);
    c3.c[0] = a.c[0]; c3.c[1] = fp_neg<F>(a1z);    c3.c[2] = fp_neg<F>(a.c[2]); c3.c[3] = a3z;
    Ef4<F> b = ef_mul<F>(ef_mul<F>(c1, c2), c3);
    const Ef4<F> n = ef_mul<F>(a, b);              // the norm lies in the base field
    const u32 ninv = fp_inv<F>(n.c[0]);
    Ef4<F> o;
#pragma unroll
    for (int k = 0; k < 4; k++) o.c[k] = mont_mul<F>(b.c[k], ninv);
    return o;
}

// compute_inverse_denominators (two_adic_pcs.rs:743-780): out[i] = 1/(z - x_i), x_i = GENERATOR * w^bitrev(i).
// Optional adj[i] = out[i] - 1/z (compute_adjusted_weights).  x_i = g * prod_{bit b of i} gen(b+1).
struct CosetArgs { u32 gens[32]; u32 g; };
template <int F>
__global__ void __launch_bounds__(128) inv_denoms_kernel(u32 *out, u32 *adj, size_t n, const Ef4<F> z, const Ef4<F> zinv, const CosetArgs ca,
                                                         u32 zeta) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 x = ca.g;
    for (int b = 0; (i >> b) != 0; b++)
        if ((i >> b) & 1) x = mont_mul<F>(x, ca.gens[b + 1]);
    Ef4<F> d = z;
    d.c[0] = fp_sub<F>(z.c[0], x);
    const Ef4<F> r = ef_inv_dev<F>(d, zeta);
    reinterpret_cast<uint4 *>(out)[i] = make_uint4(r.c[0], r.c[1], r.c[2], r.c[3]);
    if (adj) reinterpret_cast<uint4 *>(adj)[i] = make_uint4(fp_sub<F>(r.c[0], zinv.c[0]), fp_sub<F>(r.c[1], zinv.c[1]),
                                                           fp_sub<F>(r.c[2], zinv.c[2]), fp_sub<F>(r.c[3], zinv.c[3]));
}

// columnwise_dot_product: partial[chunk][j] = sum_{i in chunk} mat[i][j] * v[i].  grid = (column groups of 256, row chunks).
constexpr int COL_THREADS = 256;
constexpr int COL_ROWS = 2048;      // rows per chunk
constexpr int COL_STAGE = 256;      // rows of v staged in shared memory at a time
template <int F>
__global__ void __launch_bounds__(COL_THREADS) columnwise_dot_kernel(const u32 *mat, size_t h, size_t w, const u32 *v, u32 *partial) {
    __shared__ uint4 vs[COL_STAGE];
    const size_t j = (size_t)blockIdx.x * COL_THREADS + threadIdx.x;
    const size_t r0 = (size_t)blockIdx.y * COL_ROWS, r1 = min(h, r0 + COL_ROWS);
    u64 acc[4] = {0, 0, 0, 0};
    for (size_t rs = r0; rs < r1; rs += COL_STAGE) {
        const size_t n = min((size_t)COL_STAGE, r1 - rs);
        __syncthreads();
        if (threadIdx.x < n) vs[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(v) + rs + threadIdx.x);
        __syncthreads();
        if (j < w) {
            const u32 *mp = mat + rs * w + j;
#pragma unroll 4
            for (size_t i = 0; i < n; i++) {
                const u32 m = __ldg(mp + i * w);
                const uint4 e = vs[i];
                lazy_mac<F>(acc[0], m, e.x); lazy_mac<F>(acc[1], m, e.y); lazy_mac<F>(acc[2], m, e.z); lazy_mac<F>(acc[3], m, e.w);
            }
        }
    }
    if (j < w)
        reinterpret_cast<uint4 *>(partial)[(size_t)blockIdx.y * w + j] =
            make_uint4(lazy_finish<F>(acc[0]), lazy_finish<F>(acc[1]), lazy_finish<F>(acc[2]), lazy_finish<F>(acc[3]));
}
// Vectorised variant (w % 4 == 0, 16-byte aligned rows): a thread owns 4 adjacent columns (one uint4 load per row, 16 lazy
// accumulators); a warp-load covers 512 contiguous bytes and 4 rows are in flight per thread.
template <int F>
__global__ void __launch_bounds__(128) columnwise_dot_vec_kernel(const u32 *mat, size_t h, size_t w, const u32 *v, u32 *partial) {
    __shared__ uint4 vs[COL_STAGE];
    const size_t j4 = (size_t)blockIdx.x * 128 + threadIdx.x;     // index of the 4-column group
    const size_t w4 = w >> 2;
    const size_t r0 = (size_t)blockIdx.y * COL_ROWS, r1 = min(h, r0 + COL_ROWS);
    u64 acc[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[c][k] = 0;
    for (size_t rs = r0; rs < r1; rs += COL_STAGE) {
        const size_t n = min((size_t)COL_STAGE, r1 - rs);
        __syncthreads();
        for (size_t t = threadIdx.x; t < n; t += 128) vs[t] = __ldg(reinterpret_cast<const uint4 *>(v) + rs + t);
        __syncthreads();
        if (j4 < w4) {
            const uint4 *mp = reinterpret_cast<const uint4 *>(mat + rs * w) + j4;
#pragma unroll 4
            for (size_t i = 0; i < n; i++) {
                const uint4 m = __ldg(mp + i * w4);
                const uint4 e = vs[i];
                const u32 mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    lazy_mac<F>(acc[c][0], mm[c], e.x); lazy_mac<F>(acc[c][1], mm[c], e.y);
                    lazy_mac<F>(acc[c][2], mm[c], e.z); lazy_mac<F>(acc[c][3], mm[c], e.w);
                }
            }
        }
    }
    if (j4 < w4) {
#pragma unroll
        for (int c = 0; c < 4; c++)
            reinterpret_cast<uint4 *>(partial)[(size_t)blockIdx.y * w + 4 * j4 + c] =
                make_uint4(lazy_finish<F>(acc[c][0]), lazy_finish<F>(acc[c][1]), lazy_finish<F>(acc[c][2]), lazy_finish<F>(acc[c][3]));
    }
}
// out[j] = scale * sum_chunks partial[chunk][j]
template <int F>
__global__ void columnwise_finish_kernel(const u32 *partial, size_t n_chunks, size_t w, u32 *out, const Ef4<F> scale, int has_scale) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    Ef4<F> s; s.c[0] = s.c[1] = s.c[2] = s.c[3] = 0;
    for (size_t c = 0; c < n_chunks; c++) {
        const uint4 e = reinterpret_cast<const uint4 *>(partial)[c * w + j];
        s.c[0] = fp_add<F>(s.c[0], e.x); s.c[1] = fp_add<F>(s.c[1], e.y); s.c[2] = fp_add<F>(s.c[2], e.z); s.c[3] = fp_add<F>(s.c[3], e.w);
    }
    if (has_scale) s = ef_mul<F>(scale, s);
    reinterpret_cast<uint4 *>(out)[j] = make_uint4(s.c[0], s.c[1], s.c[2], s.c[3]);
}

// rowwise dot with powers of alpha: out[i] = sum_j alpha^j * mat[i][j].  One warp per row, lanes stride over the columns;
// the powers table (w EF4 values) is read through L1/L2 (it is shared by every row).
constexpr int ROW_THREADS = 256;
template <int F>
__global__ void __launch_bounds__(ROW_THREADS) rowwise_dot_kernel(const u32 *mat, size_t h, size_t w, const u32 *pw, u32 *out) {
    const size_t row = (size_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned lane = threadIdx.x & 31;
    if (row >= h) return;
    const u32 *mp = mat + row * w;
    u64 acc[4] = {0, 0, 0, 0};
    for (size_t j = lane; j < w; j += 32) {
        const u32 m = __ldg(mp + j);
        const uint4 e = __ldg(reinterpret_cast<const uint4 *>(pw) + j);
        lazy_mac<F>(acc[0], m, e.x); lazy_mac<F>(acc[1], m, e.y); lazy_mac<F>(acc[2], m, e.z); lazy_mac<F>(acc[3], m, e.w);
    }
    u32 r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = lazy_finish<F>(acc[k]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = fp_add<F>(r[k], __shfl_down_sync(0xffffffffu, r[k], off));
    if (lane == 0) reinterpret_cast<uint4 *>(out)[row] = make_uint4(r[0], r[1], r[2], r[3]);
}
// Vectorised variant (w % 4 == 0): the powers table sits in shared memory; a warp owns RW consecutive rows and each lane
// 4-column groups, so one set of 4 powers (4 x LDS.128) feeds 4*RW elements and every global load is a 16-byte uint4.
constexpr int RW = 4;
template <int F>
__global__ void __launch_bounds__(ROW_THREADS) rowwise_dot_vec_kernel(const u32 *mat, size_t h, size_t w, const u32 *pw, u32 *out) {
    extern __shared__ uint4 pws[];
    for (size_t t = threadIdx.x; t < w; t += ROW_THREADS) pws[t] = __ldg(reinterpret_cast<const uint4 *>(pw) + t);
    __syncthreads();
    const size_t row0 = ((size_t)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5)) * RW;
    const unsigned lane = threadIdx.x & 31;
    if (row0 >= h) return;
    const size_t w4 = w >> 2;
    u64 acc[RW][4];
#pragma unroll
    for (int r = 0; r < RW; r++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[r][k] = 0;
    for (size_t j4 = lane; j4 < w4; j4 += 32) {
        uint4 m[RW];
#pragma unroll
        for (int r = 0; r < RW; r++)
            m[r] = (row0 + r < h) ? __ldg(reinterpret_cast<const uint4 *>(mat + (row0 + r) * w) + j4) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint4 e = pws[4 * j4 + c];
#pragma unroll
            for (int r = 0; r < RW; r++) {
                const u32 mm = c == 0 ? m[r].x : c == 1 ? m[r].y : c == 2 ? m[r].z : m[r].w;
                lazy_mac<F>(acc[r][0], mm, e.x); lazy_mac<F>(acc[r][1], mm, e.y); lazy_mac<F>(acc[r][2], mm, e.z); lazy_mac<F>(acc[r][3], mm, e.w);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RW; r++) {
        u32 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = lazy_finish<F>(acc[r][k]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = fp_add<F>(v[k], __shfl_down_sync(0xffffffffu, v[k], off));
        if (lane == 0 && row0 + r < h) reinterpret_cast<uint4 *>(out)[row0 + r] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}
// alpha powers table: pw[j] = alpha^j (sequential per block of 64 with a precomputed alpha^64 stride would be faster;
// w <= a few thousand, so one thread per block of 32 entries is plenty)
template <int F> __global__ void alpha_powers_kernel(u32 *pw, size_t w, const Ef4<F> alpha) {
    const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t j0 = blk * 32;
    if (j0 >= w) return;
    // alpha^j0 by square and multiply
    Ef4<F> cur; cur.c[0] = Fp<F>::ONE; cur.c[1] = cur.c[2] = cur.c[3] = 0;
    Ef4<F> base = alpha;
    for (size_t e = j0; e; e >>= 1) { if (e & 1) cur = ef_mul<F>(cur, base); base = ef_mul<F>(base, base); }
    for (size_t j = j0; j < w && j < j0 + 32; j++) {
        reinterpret_cast<uint4 *>(pw)[j] = make_uint4(cur.c[0], cur.c[1], cur.c[2], cur.c[3]);
        cur = ef_mul<F>(cur, alpha);
    }
}

// ro[i] += coeff * (yred - r[i]) * inv_denom[i]   (two_adic_pcs.rs:640-657)
template <int F>
__global__ void __launch_bounds__(256) open_reduce_kernel(u32 *ro, const u32 *r, const u32 *invd, size_t h, const Ef4<F> coeff, const Ef4<F> yred) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    const uint4 rv = __ldg(reinterpret_cast<const uint4 *>(r) + i), dv = __ldg(reinterpret_cast<const uint4 *>(invd) + i);
    uint4 acc = reinterpret_cast<uint4 *>(ro)[i];
    Ef4<F> d, inv;
    d.c[0] = fp_sub<F>(yred.c[0], rv.x); d.c[1] = fp_sub<F>(yred.c[1], rv.y); d.c[2] = fp_sub<F>(yred.c[2], rv.z); d.c[3] = fp_sub<F>(yred.c[3], rv.w);
    inv.c[0] = dv.x; inv.c[1] = dv.y; inv.c[2] = dv.z; inv.c[3] = dv.w;
    const Ef4<F> t = ef_mul<F>(ef_mul<F>(coeff, d), inv);
    reinterpret_cast<uint4 *>(ro)[i] = make_uint4(fp_add<F>(acc.x, t.c[0]), fp_add<F>(acc.y, t.c[1]), fp_add<F>(acc.z, t.c[2]), fp_add<F>(acc.w, t.c[3]));
}

// ---- host side ---------------------------------------------------------------------------------
template <int F> static Ef4<F> to_ef(const u32 *p) { Ef4<F> e; for (int k = 0; k < 4; k++) e.c[k] = p[k]; return e; }
static inline unsigned nb(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

template <int F> static u32 frobenius_zeta() { return fp_pow<F>(to_monty<F>(Fp<F>::EXT_W), ((u64)Fp<F>::P - 1) / 4); }

template <int F>
static int32_t inv_denoms_impl(p3gpu_ctx *ctx, unsigned log_h, const u32 *z, const u32 *zinv, u32 *d_out, u32 *d_adj) {
    CosetArgs ca;
    for (u32 k = 0; k < 32; k++) ca.gens[k] = k <= Fp<F>::TWO_ADICITY ? two_adic_generator<F>(k) : Fp<F>::ONE;
    ca.g = to_monty<F>(Fp<F>::GEN);
    const size_t n = (size_t)1 << log_h;
    Ef4<F> zi; for (int k = 0; k < 4; k++) zi.c[k] = zinv ? zinv[k] : 0;
    inv_denoms_kernel<F><<<nb(n, 128), 128, 0, ctx->stream>>>(d_out, d_adj, n, to_ef<F>(z), zi, ca, frobenius_zeta<F>());
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}
int32_t open_inv_denoms(p3gpu_ctx *ctx, int field, unsigned log_h, const u32 *z, const u32 *zinv, u32 *d_out, u32 *d_adj) {
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    const unsigned adicity = field == BABY_BEAR ? Fp<BABY_BEAR>::TWO_ADICITY : Fp<KOALA_BEAR>::TWO_ADICITY;
    P3_CHECK(log_h <= adicity, P3GPU_EINVAL, "coset of size 2^%u exceeds the two-adicity", log_h);
    P3_CHECK(d_adj == nullptr || zinv != nullptr, P3GPU_EINVAL, "adjusted weights need 1/z");
    return field == BABY_BEAR ? inv_denoms_impl<BABY_BEAR>(ctx, log_h, z, zinv, d_out, d_adj)
                              : inv_denoms_impl<KOALA_BEAR>(ctx, log_h, z, zinv, d_out, d_adj);
}

template <int F>
static int32_t columnwise_impl(p3gpu_ctx *ctx, const u32 *d_mat, size_t h, size_t w, const u32 *d_vec, u32 *d_out, const u32 *scale) {
    const size_t n_chunks = (h + COL_ROWS - 1) / COL_ROWS;
    void *partial = nullptr;
    P3_TRY(ctx_scratch2(ctx, n_chunks * w * 16, &partial));
    if (w % 4 == 0 && reinterpret_cast<uintptr_t>(d_mat) % 16 == 0) {
        dim3 grid(nb(w / 4, 128), (unsigned)n_chunks);
        columnwise_dot_vec_kernel<F><<<grid, 128, 0, ctx->stream>>>(d_mat, h, w, d_vec, (u32 *)partial);
    } else {
        dim3 grid(nb(w, COL_THREADS), (unsigned)n_chunks);
        columnwise_dot_kernel<F><<<grid, COL_THREADS, 0, ctx->stream>>>(d_mat, h, w, d_vec, (u32 *)partial);
    }
    Ef4<F> s; for (int k = 0; k < 4; k++) s.c[k] = scale ? scale[k] : 0;
    columnwise_finish_kernel<F><<<nb(w, 128), 128, 0, ctx->stream>>>((const u32 *)partial, n_chunks, w, d_out, s, scale != nullptr);
    ctx->launches += 2;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}
int32_t open_columnwise_dot(p3gpu_ctx *ctx, int field, const u32 *d_mat, size_t h, size_t w, const u32 *d_vec, u32 *d_out, const u32 *scale) {
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    P3_CHECK(h >= 1 && w >= 1 && w < (1ull << 31), P3GPU_EINVAL, "bad matrix shape %zu x %zu", h, w);
    return field == BABY_BEAR ? columnwise_impl<BABY_BEAR>(ctx, d_mat, h, w, d_vec, d_out, scale)
                              : columnwise_impl<KOALA_BEAR>(ctx, d_mat, h, w, d_vec, d_out, scale);
}

template <int F>
static int32_t rowwise_impl(p3gpu_ctx *ctx, const u32 *d_mat, size_t h, size_t w, const u32 *alpha, u32 *d_out) {
    void *pw = nullptr;
    P3_TRY(ctx_scratch2(ctx, w * 16, &pw));
    alpha_powers_kernel<F><<<nb((w + 31) / 32, 64), 64, 0, ctx->stream>>>((u32 *)pw, w, to_ef<F>(alpha));
    if (w % 4 == 0 && reinterpret_cast<uintptr_t>(d_mat) % 16 == 0 && w * 16 <= 200 * 1024) {
        auto kern = rowwise_dot_vec_kernel<F>;
        const size_t smem = w * 16;
        if (smem > 48 * 1024) P3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<nb(h, (ROW_THREADS / 32) * RW), ROW_THREADS, smem, ctx->stream>>>(d_mat, h, w, (const u32 *)pw, d_out);
    } else {
        rowwise_dot_kernel<F><<<nb(h, ROW_THREADS / 32), ROW_THREADS, 0, ctx->stream>>>(d_mat, h, w, (const u32 *)pw, d_out);
    }
    ctx->launches += 2;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}
int32_t open_rowwise_dot(p3gpu_ctx *ctx, int field, const u32 *d_mat, size_t h, size_t w, const u32 *alpha, u32 *d_out) {
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    P3_CHECK(h >= 1 && w >= 1 && w < (1ull << 31), P3GPU_EINVAL, "bad matrix shape %zu x %zu", h, w);
    return field == BABY_BEAR ? rowwise_impl<BABY_BEAR>(ctx, d_mat, h, w, alpha, d_out) : rowwise_impl<KOALA_BEAR>(ctx, d_mat, h, w, alpha, d_out);
}

int32_t open_reduce(p3gpu_ctx *ctx, int field, u32 *d_ro, const u32 *d_r, const u32 *d_invd, size_t h, const u32 *coeff, const u32 *yred) {
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    if (h == 0) return P3GPU_OK;
    if (field == BABY_BEAR) open_reduce_kernel<BABY_BEAR><<<nb(h, 256), 256, 0, ctx->stream>>>(d_ro, d_r, d_invd, h, to_ef<BABY_BEAR>(coeff), to_ef<BABY_BEAR>(yred));
    else open_reduce_kernel<KOALA_BEAR><<<nb(h, 256), 256, 0, ctx->stream>>>(d_ro, d_r, d_invd, h, to_ef<KOALA_BEAR>(coeff), to_ef<KOALA_BEAR>(yred));
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}

}  // namespace p3
