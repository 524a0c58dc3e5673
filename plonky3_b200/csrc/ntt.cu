// Batched NTT / coset LDE for row-major matrices over BabyBear / KoalaBear on sm_100a.
//
// Replaces Radix2DitParallel (dft/src/radix_2_dit_parallel.rs:30-515) behind TwoAdicSubgroupDft
// (dft/src/traits.rs:28-291).  Not a port: the reference runs two cache-blocked half networks separated by
// row bit-reversals on CPU threads; here ONE kernel family implements a Cooley-Tukey network that maps
// natural-order input to bit-reversed-order output ("network order"), executed as 1-3 passes over HBM:
//
//   * a pass owns the butterfly layers [l0, l1) of the size-2^n network.  A CTA takes a tile of R = 2^(l1-l0)
//     rows (all rows that agree on the top l0 and the low n-l1 index bits) x CT adjacent columns, stages it in
//     shared memory, runs the layers as radix-16 register steps (4 layers per shared-memory round trip), and
//     writes the tile back.  Row segments of CT*4 bytes are contiguous, so any row permutation (bit reversal on
//     input or output) is free: it only changes which 64/128-byte segments a tile touches.
//   * layer l uses one twiddle per block q (the reference's "twiddles with the coset shift baked in",
//     radix_2_dit_parallel.rs:80-115):  z_l[q] = shift^(N/2^(l+1)) * w_(2^(l+1))^bitrev_l(q).  They live in a
//     heap-ordered table Z[2^l + q]; a tile needs R-1 of them (contiguous runs per layer) and stages them in
//     shared memory next to the data.
//   * butterflies use Shoup multiplication by the precomputed twiddle and lazy [0, 2p) reduction:
//     3 multiply-pipe + 6 ALU-pipe instructions each (field.cuh: ct_butterfly).
//
// coset_lde_batch = inverse network (root^-1, scale 1/h) producing coefficients in network (bit-reversed) order,
// then per coset a forward network reading those coefficients through a bit-reversed row map and leaving the
// evaluations in network order — which is exactly the bit-reversed row order the reference leaves in memory
// (radix_2_dit_parallel.rs:245, fri/src/two_adic_pcs.rs:313-318).  No standalone bit-reversal or scaling pass exists.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"

namespace p3 {

// Phase-breakdown instrumentation (profiles/README.md) is compiled in only with -DP3GPU_NTT_PROFILE; the env switches
// P3GPU_NTT_NOBFLY / NOLOAD / NOSTORE are ignored by the production build.
#ifdef P3GPU_NTT_PROFILE
#define P3_SKIP(flag) (flag)
#else
#define P3_SKIP(flag) false
#endif

struct PassArgs {
    const u32 *in;
    u32 *out;
    u32 w;         // row pitch in elements
    u32 col0;      // first column of this launch
    u32 n_ctiles;  // column tiles (of CT columns) in this launch
    u32 ct;        // tile width of the generic-width kernel variant
    u32 n_cosets;  // cosets batched in this launch (fastest-varying part of blockIdx.x)
    u32 vec16;     // fast kernel: row segments are 16-byte aligned (cp.async 16)
    u32 skip_load, skip_store;  // profiling experiments only
    u32 skip_bfly; // profiling experiment only (P3GPU_NTT_NOBFLY=1): move the data, skip the butterflies
    int log_n, l0, l1;
    const uint2 *tw;  // heap-ordered twiddles of coset 0
    int in_bitrev, out_bitrev;
    int out_sh;
    u32 out_add;  // output row = (maybe_bitrev(i) << out_sh) + out_add
    uint2 scale;
    int has_scale, final_reduce;
    size_t tw_stride;   // uint2 elements between consecutive cosets' heaps
    size_t out_stride;  // u32 elements between consecutive cosets' output blocks
    size_t in_stride;   // u32 elements between consecutive cosets' input blocks
};

template <int LOG_CT> __device__ __forceinline__ u32 sidx(u32 row, u32 c) {
    // XOR swizzle so that narrow column tiles (CT < 32) stay bank-conflict free in the stride-1 radix step
    if (LOG_CT < 5) row ^= (row >> 4) & ((32u >> LOG_CT) - 1u);
    return (row << LOG_CT) + c;
}

template <int F, int LOG_CT, int THREADS, int Q>
__device__ __forceinline__ void radix_step(u32 *data, const uint2 *tws, int r, int lam0) {
    constexpr u32 CT = 1u << LOG_CT;
    constexpr int E = 1 << Q;
    const int logD = r - lam0 - Q;
    const u32 D = 1u << logD;
    const u32 items = (1u << (r - Q)) << LOG_CT;
    for (u32 it = threadIdx.x; it < items; it += THREADS) {
        const u32 c = it & (CT - 1), g = it >> LOG_CT;
        const u32 lo = g & (D - 1), hi = g >> logD;
        const u32 base = (hi << (logD + Q)) + lo;
        const u32 node = (1u << lam0) + hi;
        u32 x[E];
#pragma unroll
        for (int m = 0; m < E; m++) x[m] = data[sidx<LOG_CT>(base + m * D, c)];
#pragma unroll
        for (int j = 0; j < Q; j++) {
            const int half = E >> (j + 1);
#pragma unroll
            for (int grp = 0; grp < (1 << j); grp++) {
                const uint2 z = tws[(node << j) + grp];
#pragma unroll
                for (int t = 0; t < half; t++) ct_butterfly<F>(x[grp * 2 * half + t], x[grp * 2 * half + t + half], z);
            }
        }
#pragma unroll
        for (int m = 0; m < E; m++) data[sidx<LOG_CT>(base + m * D, c)] = x[m];
    }
    __syncthreads();
}

template <int F, int LOG_CT, int THREADS, bool VEC>
__global__ void __launch_bounds__(THREADS) ntt_pass_kernel(const PassArgs a) {
    constexpr u32 CT = 1u << LOG_CT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int r = a.l1 - a.l0;
    const u32 R = 1u << r;
    u32 *data = reinterpret_cast<u32 *>(smem_raw);
    uint2 *tws = reinterpret_cast<uint2 *>(data + ((size_t)R << LOG_CT));

    const u32 coset = blockIdx.x % a.n_cosets, bx = blockIdx.x / a.n_cosets;
    const u32 ctile = bx % a.n_ctiles, tile = bx / a.n_ctiles;
    const int lowbits = a.log_n - a.l1;
    const u32 L = tile & ((1u << lowbits) - 1u), T = tile >> lowbits;
    const u32 col = a.col0 + ctile * CT;
    const uint2 *tw = a.tw + (size_t)coset * a.tw_stride;
    const u32 *in = a.in + (size_t)coset * a.in_stride;
    u32 *out = a.out + (size_t)coset * a.out_stride;
    const u32 out_add = a.out_add;
    const u32 ibase = (a.l0 == 0 ? 0u : (T << (a.log_n - a.l0))) | L;
    const int brsh = 32 - a.log_n;

    // stage this tile's R-1 twiddles: tws[2^lam + ql] = Z[2^(l0+lam) + T*2^lam + ql]
    for (u32 k = threadIdx.x + 1; k < R; k += THREADS) {
        const int lam = 31 - __clz(k);
        const u32 ql = k - (1u << lam);
        tws[k] = tw[((size_t)1 << (a.l0 + lam)) + ((size_t)T << lam) + ql];
    }
    // gather the tile
    if (VEC) {
        constexpr u32 CV = CT >= 4 ? CT / 4 : 1;
        for (u32 it = threadIdx.x; it < R * CV; it += THREADS) {
            const u32 c4 = it % CV, rho = it / CV;
            const u32 i = ibase | (rho << lowbits);
            const u32 row = a.in_bitrev ? (__brev(i) >> brsh) : i;
            uint4 v = *reinterpret_cast<const uint4 *>(in + (size_t)row * a.w + col + 4 * c4);
            if (a.has_scale) {
                v.x = shoup_mul<F>(v.x, a.scale); v.y = shoup_mul<F>(v.y, a.scale);
                v.z = shoup_mul<F>(v.z, a.scale); v.w = shoup_mul<F>(v.w, a.scale);
            }
            *reinterpret_cast<uint4 *>(data + sidx<LOG_CT>(rho, 4 * c4)) = v;
        }
    } else {
        for (u32 it = threadIdx.x; it < (R << LOG_CT); it += THREADS) {
            const u32 c = it & (CT - 1), rho = it >> LOG_CT;
            const u32 i = ibase | (rho << lowbits);
            const u32 row = a.in_bitrev ? (__brev(i) >> brsh) : i;
            u32 v = in[(size_t)row * a.w + col + c];
            if (a.has_scale) v = shoup_mul<F>(v, a.scale);
            data[sidx<LOG_CT>(rho, c)] = v;
        }
    }
    __syncthreads();

    int lam0 = 0;
    const int q0 = (r & 3) ? (r & 3) : 4;
    switch (q0) {
        case 1: radix_step<F, LOG_CT, THREADS, 1>(data, tws, r, 0); break;
        case 2: radix_step<F, LOG_CT, THREADS, 2>(data, tws, r, 0); break;
        case 3: radix_step<F, LOG_CT, THREADS, 3>(data, tws, r, 0); break;
        default: radix_step<F, LOG_CT, THREADS, 4>(data, tws, r, 0); break;
    }
    for (lam0 = q0; lam0 < r; lam0 += 4) radix_step<F, LOG_CT, THREADS, 4>(data, tws, r, lam0);

    // scatter the tile
    if (VEC) {
        constexpr u32 CV = CT >= 4 ? CT / 4 : 1;
        for (u32 it = threadIdx.x; it < R * CV; it += THREADS) {
            const u32 c4 = it % CV, rho = it / CV;
            const u32 i = ibase | (rho << lowbits);
            const u32 row = ((a.out_bitrev ? (__brev(i) >> brsh) : i) << a.out_sh) + out_add;
            uint4 v = *reinterpret_cast<const uint4 *>(data + sidx<LOG_CT>(rho, 4 * c4));
            if (a.final_reduce) { v.x = fp_reduce<F>(v.x); v.y = fp_reduce<F>(v.y); v.z = fp_reduce<F>(v.z); v.w = fp_reduce<F>(v.w); }
            *reinterpret_cast<uint4 *>(out + (size_t)row * a.w + col + 4 * c4) = v;
        }
    } else {
        for (u32 it = threadIdx.x; it < (R << LOG_CT); it += THREADS) {
            const u32 c = it & (CT - 1), rho = it >> LOG_CT;
            const u32 i = ibase | (rho << lowbits);
            const u32 row = ((a.out_bitrev ? (__brev(i) >> brsh) : i) << a.out_sh) + out_add;
            u32 v = data[sidx<LOG_CT>(rho, c)];
            if (a.final_reduce) v = fp_reduce<F>(v);
            out[(size_t)row * a.w + col + c] = v;
        }
    }
}

// ---- fast path: the whole pass as TWO register networks with one shared-memory exchange ------------------------
// For 7 <= r <= 10 the r layers split as Q1 + Q2 (Q2 = ceil(r/2) <= 5).  Step 1 loads 2^Q1 rows per thread straight from
// global memory (stride 2^Q2 local rows), runs Q1 layers in registers and parks the results in shared memory; step 2
// reads 2^Q2 consecutive local rows, runs Q2 layers and stores straight to global memory.  Per element and pass that is
// one shared store + one shared load (the generic kernel does 2 per radix step plus the staging copy).
// Shared layout: local row rho, column c at (rho >> Q2) * gstride + (rho & (2^Q2-1)) * CT + c with gstride = 2^Q2*CT + pad,
// pad chosen so that gstride = CT (mod 32): both access patterns are bank-conflict free for ANY tile width CT, which lets
// one launch class take a non-power-of-two remainder tile (e.g. 100 = 5 x 16 + 20 columns) without sector over-fetch.
template <int Q> __host__ __device__ constexpr u32 brev_const(u32 m) {
    u32 r = 0;
    for (int b = 0; b < Q; b++) r |= ((m >> b) & 1u) << (Q - 1 - b);
    return r;
}

template <int F, int Q>
__device__ __forceinline__ void reg_network(u32 (&x)[1 << Q], const uint2 *tws, u32 node) {
    constexpr int E = 1 << Q;
#pragma unroll
    for (int j = 0; j < Q; j++) {
        const int half = E >> (j + 1);
#pragma unroll
        for (int grp = 0; grp < (1 << j); grp++) {
            const uint2 z = tws[(node << j) + grp];
#pragma unroll
            for (int t = 0; t < half; t++) ct_butterfly<F>(x[grp * 2 * half + t], x[grp * 2 * half + t + half], z);
        }
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Persistent, double-buffered pass kernel.  A CTA (one per SM) walks over tiles of 2^R_LOG rows x `ct` columns:
//   * tile k+1 (rows as 16-byte cp.async/LDGSTS copies, plus its 2^R_LOG - 1 twiddles) streams into the second shared
//     buffer while tile k is computed, so HBM latency overlaps the integer work;
//   * step 1 runs Q1 layers in registers IN PLACE in shared memory, step 2 runs Q2 layers and streams the results to
//     global memory (per-row TMA bulk stores were tried and rejected: UBLKCP is a warp-uniform instruction, so one
//     copy per lane serialises into a 32-iteration R2UR/PLOP3 loop per warp: +25 % instructions, profiles/README.md);
//   * all tiles of a pass have the same runtime width ct (16/20/24 columns; w = 100 -> 5 x 20) so that ONE launch covers
//     every column and neighbouring tiles share DRAM bursts through L2.
template <int F, int R_LOG, int CT_T, int THREADS, int NBUF>   // CT_T: compile-time tile width (16/20/24) or 0 = runtime a.ct
__global__ void __launch_bounds__(THREADS, NBUF == 2 ? 1 : (CT_T == 16 && THREADS == 256) ? 3 : 2) ntt_pass_fast_kernel(const __grid_constant__ PassArgs a) {
    constexpr int Q2 = (R_LOG + 1) / 2, Q1 = R_LOG - Q2;
    constexpr u32 E1 = 1u << Q1, E2 = 1u << Q2, R = 1u << R_LOG;
    const u32 CT = CT_T ? (u32)CT_T : a.ct;
    const u32 padw = (CT + 32u - ((E2 * CT) & 31u)) & 31u;
    const u32 gstride = E2 * CT + padw;
    const u32 buf_words = (E1 * gstride + 3u) & ~3u;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32 *data0 = reinterpret_cast<u32 *>(smem_raw);
    uint2 *tws0 = reinterpret_cast<uint2 *>(data0 + NBUF * buf_words);

    const int lowbits = a.log_n - a.l1;
    const int brsh = 32 - a.log_n;
    const u32 n_row_tiles = 1u << (a.log_n - R_LOG);
    const u32 total = n_row_tiles * a.n_ctiles * a.n_cosets;
    const bool vec16 = a.vec16 != 0;       // row segments 16-byte aligned in global memory (loads AND stores)
    const bool shared_tw = (a.l0 == 0);    // first pass of a network: every tile uses the same twiddles

    auto decode = [&](u32 t, u32 &coset, u32 &col, u32 &cw, u32 &T, u32 &ibase) {
        coset = t % a.n_cosets;
        const u32 bx = t / a.n_cosets;
        const u32 ctile = bx % a.n_ctiles, tile = bx / a.n_ctiles;
        const u32 L = tile & ((1u << lowbits) - 1u);
        T = tile >> lowbits;
        col = ctile * CT;
        cw = min(CT, a.w - col);
        ibase = (a.l0 == 0 ? 0u : (T << (a.log_n - a.l0))) | L;
    };
    auto issue_twiddles = [&](u32 coset, u32 T, uint2 *tws) {
        const uint2 *tw = a.tw + (size_t)coset * a.tw_stride;
        for (u32 k = threadIdx.x + 1; k < R; k += THREADS) {
            const int lam = 31 - __clz(k);
            const u32 ql = k - (1u << lam);
            cp_async8(tws + k, tw + ((size_t)1 << (a.l0 + lam)) + ((size_t)T << lam) + ql);
        }
    };
    auto issue = [&](u32 t, u32 buf) {
        u32 coset, col, cw, T, ibase;
        decode(t, coset, col, cw, T, ibase);
        u32 *data = data0 + buf * buf_words;
        const u32 *in = a.in + (size_t)coset * a.in_stride + col;
        if (!shared_tw || a.n_cosets > 1) issue_twiddles(coset, T, tws0 + buf * R);
        // chunk = 16 bytes (4 columns) when aligned, else one element
        const u32 cpr = vec16 ? (cw >> 2) : cw;                  // chunks per row segment
        const u32 rs = (THREADS / cpr) & ~(E2 - 1u);             // rows per sweep: a multiple of E2 keeps the shared address linear
        const u32 nthr = rs * cpr;
        if (rs == 0) {
            // fewer than E2 whole rows per sweep (wide unaligned tiles): plain index arithmetic per chunk
            for (u32 it = threadIdx.x; it < R * cpr; it += THREADS) {
                const u32 rho = it / cpr, ch = it - rho * cpr;
                const u32 e = vec16 ? 4u * ch : ch;
                const u32 i = ibase | (rho << lowbits);
                const u32 row = a.in_bitrev ? (__brev(i) >> brsh) : i;
                u32 *dst = data + (rho >> Q2) * gstride + (rho & (E2 - 1u)) * CT + e;
                const u32 *src = in + (size_t)row * a.w + e;
                if (vec16) cp_async16(dst, src); else cp_async4(dst, src);
            }
        } else if (threadIdx.x < nthr) {
            const u32 rho0 = threadIdx.x / cpr, ch = threadIdx.x - rho0 * cpr;
            const u32 e = vec16 ? 4u * ch : ch;
            u32 *dst = data + (rho0 >> Q2) * gstride + (rho0 & (E2 - 1u)) * CT + e;
            const u32 dstep = (rs >> Q2) * gstride;
            if (!a.in_bitrev) {
                const u32 *src = in + (size_t)(ibase | (rho0 << lowbits)) * a.w + e;
                const size_t sstep = ((size_t)rs << lowbits) * a.w;
                for (u32 rho = rho0; rho < R; rho += rs, dst += dstep, src += sstep) {
                    if (vec16) cp_async16(dst, src); else cp_async4(dst, src);
                }
            } else {
                for (u32 rho = rho0; rho < R; rho += rs, dst += dstep) {
                    const u32 row = __brev(ibase | (rho << lowbits)) >> brsh;
                    const u32 *src = in + (size_t)row * a.w + e;
                    if (vec16) cp_async16(dst, src); else cp_async4(dst, src);
                }
            }
        }
        cp_async_commit();
    };

    u32 t = blockIdx.x;
    if (t >= total) return;
    if (shared_tw && a.n_cosets == 1) issue_twiddles(0, 0, tws0);   // once per CTA, lands with the first tile's group
    if (NBUF == 2) issue(t, 0);
    for (u32 k = 0; t < total; t += gridDim.x, k++) {
        const u32 buf = NBUF == 2 ? (k & 1u) : 0u;
        __syncthreads();   // every warp is done reading the buffer that is refilled next
        if (NBUF == 2) {
            if (t + gridDim.x < total) { issue(t + gridDim.x, buf ^ 1u); cp_async_wait<1>(); }
            else cp_async_wait<0>();
        } else {   // single buffer: other resident CTAs of this SM compute while this one waits for its tile
            if (!P3_SKIP(a.skip_load)) issue(t, 0);
            cp_async_wait<0>();
        }
        __syncthreads();
        u32 *data = data0 + buf * buf_words;
        const uint2 *tws = (shared_tw && a.n_cosets == 1) ? tws0 : tws0 + buf * R;  // NBUF == 1: buf == 0
        u32 coset, col, cw, T, ibase;
        decode(t, coset, col, cw, T, ibase);
        const u32 dg = THREADS / cw, dc = THREADS - dg * cw;
        // ---- step 1 (in place in shared memory): item (g, c) holds local rows g + m*E2, m < E1
        {
            u32 g = threadIdx.x / cw, c = threadIdx.x - g * cw;
            for (; g < E2; ) {
                u32 *sp = data + g * CT + c;
                u32 x[E1];
#pragma unroll
                for (u32 m = 0; m < E1; m++) x[m] = sp[m * gstride];
                if (a.has_scale) {
#pragma unroll
                    for (u32 m = 0; m < E1; m++) x[m] = shoup_mul<F>(x[m], a.scale);
                }
                if (!P3_SKIP(a.skip_bfly)) reg_network<F, Q1>(x, tws, 1u);
#pragma unroll
                for (u32 m = 0; m < E1; m++) sp[m * gstride] = x[m];
                c += dc; g += dg;
                if (c >= cw) { c -= cw; g++; }
            }
        }
        __syncthreads();
        // ---- step 2: item (g, c) holds local rows g*E2 + m, m < E2
        {
            u32 *out = a.out + (size_t)coset * a.out_stride + col;
            // out row(m) = ((row0 + K_m * S) << out_sh) + out_add: natural: K_m = m, S = 1 << lowbits;
            //                                                        bit-reversed: K_m = brev_Q2(m), S = 1 << (l0+Q1)
            const size_t sstride = ((size_t)(a.out_bitrev ? (1u << (a.l0 + Q1)) : (1u << lowbits)) << a.out_sh) * a.w;
            u32 g = threadIdx.x / cw, c = threadIdx.x - g * cw;
            for (; g < E1; ) {
                u32 *sp = data + g * gstride + c;
                u32 x[E2];
#pragma unroll
                for (u32 m = 0; m < E2; m++) x[m] = sp[m * CT];
                if (!P3_SKIP(a.skip_bfly)) reg_network<F, Q2>(x, tws, E1 + g);
                if (a.final_reduce) {
#pragma unroll
                    for (u32 m = 0; m < E2; m++) x[m] = fp_reduce<F>(x[m]);
                }
                {
                    const u32 i0 = ibase | (g << (lowbits + Q2));
                    const u32 row0 = ((a.out_bitrev ? (__brev(i0) >> brsh) : i0) << a.out_sh) + a.out_add;
                    u32 *p = out + (size_t)row0 * a.w + c;
                    if (P3_SKIP(a.skip_store)) { u32 acc = 0;
#pragma unroll
                        for (u32 m = 0; m < E2; m++) acc ^= x[m];
                        if (acc == 0x12345678u) p[0] = acc;
                    } else if (a.out_bitrev) {
#pragma unroll
                        for (u32 m = 0; m < E2; m++) p[brev_const<Q2>(m) * sstride] = x[m];
                    } else {
#pragma unroll
                        for (u32 m = 0; m < E2; m++) p[m * sstride] = x[m];
                    }
                }
                c += dc; g += dg;
                if (c >= cw) { c -= cw; g++; }
            }
        }
    }
}

// ---- twiddle heaps ---------------------------------------------------------------------------
struct TwGenArgs {
    u32 sigma[32];  // sigma[l] = shift^(N/2^(l+1)), Montgomery
    u32 roots[32];  // roots[k] = primitive 2^k-th root (or its inverse), Montgomery
};
template <int F> __global__ void gen_twiddle_heap(uint2 *Z, int log_n, const TwGenArgs a) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ((size_t)1 << log_n)) return;
    if (idx == 0) { Z[0] = make_uint2(0, 0); return; }
    const int l = 63 - __clzll((long long)idx);
    const u32 q = (u32)(idx - ((size_t)1 << l));
    u32 z = a.sigma[l];
    for (int b = 0; b < l; b++)
        if ((q >> b) & 1u) z = mont_mul<F>(z, a.roots[b + 2]);
    Z[idx] = shoup_pair<F>(from_monty<F>(z));
}

// row i *= base^i  (dft/src/util.rs:32-55 coset_shift_cols), used by coset_idft_batch only
struct PowArgs { u32 pw[32]; };  // pw[k] = base^(2^k), Montgomery
template <int F> __global__ void scale_rows_by_powers(u32 *m, size_t h, size_t w, const PowArgs a) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= h * w) return;
    size_t row = idx / w;
    u32 s = Fp<F>::ONE;
    for (int k = 0; row; k++, row >>= 1)
        if (row & 1) s = mont_mul<F>(s, a.pw[k]);
    m[idx] = mont_mul<F>(m[idx], s);
}
__global__ void broadcast_row(const u32 *in, u32 *out, size_t rows, size_t w) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < rows * w) out[idx] = in[idx % w];
}

// Heap(s) for the size-2^log_n network.  added_bits = 0: one heap for (shift, inverse).  added_bits > 0 (LDE): 2^added_bits
// heaps back to back, block cb for the coset shift * g_big^bitrev(cb).
template <int F>
static int32_t get_twiddles(p3gpu_ctx *ctx, int log_n, int added_bits, u32 shift, int inverse, const uint2 **out) {
    TwiddleKey key{F, log_n, shift, inverse + 2 * added_bits};
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        auto it = ctx->twiddles.find(key);
        if (it != ctx->twiddles.end()) { *out = it->second; return P3GPU_OK; }
    }
    const size_t N = (size_t)1 << log_n, n_cosets = (size_t)1 << added_bits;
    uint2 *Z = nullptr;
    P3_CUDA(cudaMalloc(&Z, n_cosets * N * sizeof(uint2)));
    const u32 g_big = two_adic_generator<F>((u32)(log_n + added_bits));
    for (size_t cb = 0; cb < n_cosets; cb++) {
        size_t c = 0;
        for (int b = 0; b < added_bits; b++) c |= ((cb >> b) & 1) << (added_bits - 1 - b);
        const u32 s = mont_mul<F>(shift, fp_pow<F>(g_big, c));
        TwGenArgs a;
        for (int l = 0; l < 32; l++) { a.sigma[l] = Fp<F>::ONE; a.roots[l] = Fp<F>::ONE; }
        for (int l = 0; l < log_n; l++) a.sigma[l] = fp_pow<F>(s, (u64)(N >> (l + 1)));
        for (u32 k = 0; k <= (u32)log_n && k <= Fp<F>::TWO_ADICITY; k++) {
            u32 g = two_adic_generator<F>(k);
            a.roots[k] = inverse ? fp_inv<F>(g) : g;
        }
        const unsigned blocks = (unsigned)((N + 255) / 256);
        gen_twiddle_heap<F><<<blocks, 256, 0, ctx->stream>>>(Z + cb * N, log_n, a);
        ctx->launches++;
        P3_CUDA(cudaGetLastError());
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    auto ins = ctx->twiddles.emplace(key, Z);
    if (!ins.second) { cudaFree(Z); Z = ins.first->second; }  // lost a race: keep the first table
    else ctx->twiddle_bytes += n_cosets * N * sizeof(uint2);
    *out = Z;
    return P3GPU_OK;
}

static int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return s ? atoi(s) : dflt;
}

template <int F, int LOG_CT, bool VEC>
static int32_t launch_pass_ct(p3gpu_ctx *ctx, const PassArgs &a) {
    constexpr int THREADS = 256;
    const int r = a.l1 - a.l0;
    const size_t smem = (((size_t)1 << r) << LOG_CT) * 4 + ((size_t)1 << r) * sizeof(uint2);
    auto kern = ntt_pass_kernel<F, LOG_CT, THREADS, VEC>;
    if (smem > 48 * 1024) P3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t tiles = ((size_t)1 << (a.log_n - r)) * a.n_ctiles * a.n_cosets;
    P3_CHECK(tiles < (1ull << 31), P3GPU_EINVAL, "ntt: grid too large");
    kern<<<(unsigned)tiles, THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}

template <int F, int R_LOG, int CT_T, int THREADS, int NBUF>
static int32_t launch_fast_rct(p3gpu_ctx *ctx, const PassArgs &a) {
    constexpr int Q2 = (R_LOG + 1) / 2, Q1 = R_LOG - Q2;
    const u32 ct = a.ct;
    const u32 e2 = 1u << Q2, e1 = 1u << Q1;
    const u32 padw = (ct + 32u - ((e2 * ct) & 31u)) & 31u;
    const size_t buf_words = ((size_t)e1 * (e2 * ct + padw) + 3) & ~(size_t)3;
    const size_t smem = NBUF * buf_words * 4 + NBUF * ((size_t)1 << R_LOG) * sizeof(uint2);
    auto kern = ntt_pass_fast_kernel<F, R_LOG, CT_T, THREADS, NBUF>;
    P3_CHECK(smem <= 227 * 1024, P3GPU_EINVAL, "ntt: tile does not fit shared memory");
    static size_t smem_set[64] = {0};   // per instantiation and device: raise the dynamic shared memory limit once per size
    if (smem > 48 * 1024 && smem > smem_set[ctx->device & 63]) {
        P3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[ctx->device & 63] = smem;
    }
    const size_t tiles = ((size_t)1 << (a.log_n - R_LOG)) * a.n_ctiles * a.n_cosets;
    P3_CHECK(tiles < (1ull << 31), P3GPU_EINVAL, "ntt: grid too large");
    // persistent grid: one CTA per SM (more when the tile is small enough for several to be resident)
    size_t per_sm = std::min<size_t>(NBUF == 1 ? 2048 / THREADS : 2, (227 * 1024) / (smem + 1024));
    static int num_regs = 0;   // per instantiation; benign race (same value)
    if (num_regs == 0) {
        cudaFuncAttributes fa;
        P3_CUDA(cudaFuncGetAttributes(&fa, kern));
        num_regs = std::max(fa.numRegs, 16);
    }
    const size_t by_regs = 65536 / ((size_t)THREADS * (size_t)num_regs);   // register file of the SM
    if (per_sm > by_regs) per_sm = by_regs;
    if (per_sm < 1) per_sm = 1;
    const size_t grid = std::min(tiles, per_sm * (size_t)ctx->sm_count);
    kern<<<(unsigned)grid, THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}
template <int F, int R_LOG, int CT_T>
static int32_t launch_fast_rc(p3gpu_ctx *ctx, const PassArgs &a) {
    static const int threads = env_int("P3GPU_NTT_THREADS", 256);
    if (threads == 512) return launch_fast_rct<F, R_LOG, CT_T, 512, 2>(ctx, a);   // 1 double-buffered CTA per SM
    // default: 2-3 single-buffered 256-thread CTAs per SM (one loads its tile while the others compute).  Measured on the
    // 2^20 x 100 LDE: 1.82 ms, vs 2.17 ms for 1 x 512 double-buffered; block sizes 128/192/320/384 give 1.88/1.82/1.91/1.92 ms.
    // Rejected experiments (profiles/README.md): two columns per thread with 64-bit shared accesses (2.01 ms), L2 prefetch
    // of the next tile (1.99 ms), three passes of 7+7+6 layers with small tiles (3.1 ms).
    return launch_fast_rct<F, R_LOG, CT_T, 256, 1>(ctx, a);
}
template <int F, int R_LOG>
static int32_t launch_fast_r(p3gpu_ctx *ctx, const PassArgs &a) {
    switch (a.ct) {   // compile-time widths keep every shared-memory offset an immediate
        case 16: return launch_fast_rc<F, R_LOG, 16>(ctx, a);
        case 20: return launch_fast_rc<F, R_LOG, 20>(ctx, a);
        case 24: return launch_fast_rc<F, R_LOG, 24>(ctx, a);
        default: return launch_fast_rc<F, R_LOG, 0>(ctx, a);
    }
}
template <int F>
static int32_t launch_fast(p3gpu_ctx *ctx, const PassArgs &a) {
    switch (a.l1 - a.l0) {
        case 6: return launch_fast_r<F, 6>(ctx, a);
        case 7: return launch_fast_r<F, 7>(ctx, a);
        case 8: return launch_fast_r<F, 8>(ctx, a);
        case 9: return launch_fast_r<F, 9>(ctx, a);
        default: return launch_fast_r<F, 10>(ctx, a);
    }
}

// Column tile width of the fast kernel: all tiles of a launch share one width (a ragged last tile is allowed).
// Prefer exact divisors that keep 16-byte alignment (16, 20, 24 columns = 64/80/96-byte row segments).
static u32 choose_tile_width(u32 w) {
    if (w <= 24) return w;
    static const int forced = env_int("P3GPU_NTT_CT", 0);
    if (forced) return (u32)forced;
    for (u32 ct : {16u, 20u, 24u, 12u})
        if (w % ct == 0) return ct;
    const u32 n = (w + 19) / 20;                 // ~20 columns per tile, nearly equal tiles
    u32 ct = (w + n - 1) / n;
    ct = (ct + 3) & ~3u;
    return ct > 24 ? 24 : ct;
}

// One pass over all columns.
//   fast path (7 <= r <= 10): ONE launch, tiles of choose_tile_width(w) columns (16/20/24; ragged last tile allowed).
//   generic path: columns are split greedily into power-of-two tiles of main_ct, main_ct/2, ... columns.
template <int F>
static int32_t launch_pass(p3gpu_ctx *ctx, PassArgs a, unsigned n_cosets, int main_log_ct) {
    a.n_cosets = n_cosets;
    const int r = a.l1 - a.l0;
    if (r >= 6 && r <= 10 && !env_int("P3GPU_NTT_GENERIC", 0)) {
        const u32 ct = choose_tile_width(a.w);
        // 16-byte cp.async / TMA bulk stores need every row segment of every tile 16-byte aligned on both sides
        const bool al16 = (a.w % 4 == 0) && (ct % 4 == 0) &&
                          ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out)) % 16 == 0) &&
                          ((a.in_stride | a.out_stride) % 4 == 0);
        a.col0 = 0; a.ct = ct; a.n_ctiles = (a.w + ct - 1) / ct; a.vec16 = al16;
        a.skip_bfly = env_int("P3GPU_NTT_NOBFLY", 0);
        a.skip_load = env_int("P3GPU_NTT_NOLOAD", 0); a.skip_store = env_int("P3GPU_NTT_NOSTORE", 0);
        return launch_fast<F>(ctx, a);
    }
    const bool aligned = (a.w % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.in) | reinterpret_cast<uintptr_t>(a.out)) % 16 == 0) &&
                         ((a.in_stride | a.out_stride) % 4 == 0);
    u32 col = 0, rem = a.w;
    for (int lct = main_log_ct; lct >= 0 && rem; lct--) {
        const u32 ct = 1u << lct;
        const u32 n = rem >> lct;
        if (!n) continue;
        a.col0 = col; a.n_ctiles = n; a.ct = ct;
        const bool vec = aligned && lct >= 2 && (col % 4 == 0);
        int32_t rc;
        switch (lct) {
            case 5: rc = vec ? launch_pass_ct<F, 5, true>(ctx, a) : launch_pass_ct<F, 5, false>(ctx, a); break;
            case 4: rc = vec ? launch_pass_ct<F, 4, true>(ctx, a) : launch_pass_ct<F, 4, false>(ctx, a); break;
            case 3: rc = vec ? launch_pass_ct<F, 3, true>(ctx, a) : launch_pass_ct<F, 3, false>(ctx, a); break;
            case 2: rc = vec ? launch_pass_ct<F, 2, true>(ctx, a) : launch_pass_ct<F, 2, false>(ctx, a); break;
            case 1: rc = launch_pass_ct<F, 1, false>(ctx, a); break;
            default: rc = launch_pass_ct<F, 0, false>(ctx, a); break;
        }
        P3_TRY(rc);
        col += n * ct; rem -= n * ct;
    }
    return P3GPU_OK;
}

struct NetworkPlan {
    int n_passes;
    int bounds[8];  // layer boundaries: pass k covers [bounds[k], bounds[k+1])
};
static NetworkPlan plan_passes(int log_n, int max_r) {
    NetworkPlan p;
    p.n_passes = (log_n + max_r - 1) / max_r;
    if (p.n_passes < 1) p.n_passes = 1;
    int base = log_n / p.n_passes, extra = log_n % p.n_passes;
    p.bounds[0] = 0;
    for (int k = 0; k < p.n_passes; k++) p.bounds[k + 1] = p.bounds[k] + base + (k < extra ? 1 : 0);
    return p;
}

// Runs the size-2^log_n network on n_cosets (input, output, twiddle heap) triples laid out at fixed strides.
//   src: input rows, natural order unless in_bitrev (then element i of the network is read from row bitrev(i))
//   dst: output.  Default: network order (row i = network position i).  With out_bitrev / out_sh / out_add the last pass
//        writes position i to row (bitrev(i) << out_sh) + out_add, i.e. natural order (optionally interleaved);
//        that remap cannot run in place, so multi-pass plans then keep intermediate data in tmp (h*w words per coset).
template <int F>
static int32_t run_network(p3gpu_ctx *ctx, int log_n, size_t w, const uint2 *tw, size_t tw_stride, unsigned n_cosets,
                           const u32 *src, size_t src_stride, int in_bitrev, u32 *dst, size_t dst_stride, int out_bitrev,
                           int out_sh, u32 out_add, u32 *tmp, bool has_scale, uint2 scale, bool final_reduce) {
    const int max_r = std::min(12, std::max(4, env_int("P3GPU_NTT_MAXR", 10)));
    const int main_log_ct = std::min(5, std::max(0, env_int("P3GPU_NTT_LOGCT", 4)));
    const NetworkPlan plan = plan_passes(log_n, max_r);
    const bool remap = out_bitrev || out_sh != 0 || out_add != 0;
    const size_t hw = ((size_t)1 << log_n) * w;
    if (remap && plan.n_passes > 1) P3_CHECK(tmp != nullptr, P3GPU_EINVAL, "ntt: scratch missing");
    for (int k = 0; k < plan.n_passes; k++) {
        PassArgs a;
        memset(&a, 0, sizeof a);
        const bool first = (k == 0), last = (k == plan.n_passes - 1);
        a.w = (u32)w; a.log_n = log_n; a.l0 = plan.bounds[k]; a.l1 = plan.bounds[k + 1];
        a.tw = tw; a.tw_stride = tw_stride;
        u32 *mid = remap ? tmp : dst;
        const size_t mid_stride = remap ? hw : dst_stride;
        a.in = first ? src : mid;
        a.in_stride = first ? src_stride : mid_stride;
        a.in_bitrev = first ? in_bitrev : 0;
        if (last) {
            a.out = dst; a.out_stride = dst_stride;
            a.out_bitrev = out_bitrev; a.out_sh = out_sh; a.out_add = out_add;
            a.final_reduce = final_reduce;
        } else {
            a.out = mid; a.out_stride = mid_stride;
        }
        if (first) { a.has_scale = has_scale; a.scale = scale; }
        P3_TRY(launch_pass<F>(ctx, a, n_cosets, main_log_ct));
    }
    return P3GPU_OK;
}

template <int F> static uint2 inv_height_scale(size_t h) {
    return shoup_pair<F>(from_monty<F>(fp_inv<F>(to_monty<F>((u32)(h % Fp<F>::P)))));
}

template <int F>
static int32_t dft_batch_impl(p3gpu_ctx *ctx, int kind, const u32 *d_in, u32 *d_out, size_t h, size_t w, u32 shift) {
    const int log_n = (int)log2_floor(h);
    if (log_n == 0) {  // size-1 transform is the identity for every kind
        if (d_in != d_out) P3_CUDA(cudaMemcpyAsync(d_out, d_in, w * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return P3GPU_OK;
    }
    const bool inverse = (kind == P3GPU_IDFT || kind == P3GPU_COSET_IDFT);
    const u32 tw_shift = (kind == P3GPU_COSET_DFT) ? shift : Fp<F>::ONE;
    const uint2 *tw = nullptr;
    P3_TRY(get_twiddles<F>(ctx, log_n, 0, tw_shift, inverse, &tw));
    void *tmp = nullptr;
    P3_TRY(ctx_scratch(ctx, h * w * 4, &tmp));
    P3_TRY(run_network<F>(ctx, log_n, w, tw, 0, 1, d_in, 0, 0, d_out, 0, /*out_bitrev=*/1, 0, 0, (u32 *)tmp, inverse,
                          inverse ? inv_height_scale<F>(h) : make_uint2(0, 0), true));
    if (kind == P3GPU_COSET_IDFT) {  // traits.rs:145-155: coefficient i *= shift^-i
        PowArgs pa;
        u32 b = fp_inv<F>(shift);
        for (int k = 0; k < 32; k++) { pa.pw[k] = b; b = mont_mul<F>(b, b); }
        const size_t n = h * w;
        scale_rows_by_powers<F><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_out, h, w, pa);
        ctx->launches++;
        P3_CUDA(cudaGetLastError());
    }
    return P3GPU_OK;
}

template <int F>
static int32_t coset_lde_impl(p3gpu_ctx *ctx, const u32 *d_in, size_t h, size_t w, unsigned added_bits, u32 shift, u32 *d_out,
                              int bitrev_rows) {
    const int log_n = (int)log2_floor(h);
    const size_t n_cosets = (size_t)1 << added_bits;
    if (log_n == 0) {  // a constant polynomial: every evaluation equals the single input row
        const size_t n = n_cosets * w;
        broadcast_row<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_in, d_out, n_cosets, w);
        ctx->launches++;
        P3_CUDA(cudaGetLastError());
        return P3GPU_OK;
    }
    // 1) inverse network: evaluations on H (natural) -> coefficients in network (bit-reversed) order, scaled by 1/h
    const uint2 *tw_inv = nullptr;
    P3_TRY(get_twiddles<F>(ctx, log_n, 0, Fp<F>::ONE, 1, &tw_inv));
    void *coef = nullptr;
    P3_TRY(ctx_scratch(ctx, h * w * 4, &coef));
    P3_TRY(run_network<F>(ctx, log_n, w, tw_inv, 0, 1, d_in, 0, 0, (u32 *)coef, 0, 0, 0, 0, nullptr, true, inv_height_scale<F>(h), false));

    // 2) forward networks, one per coset.  Memory block cb (h rows) holds the coset with natural index c = bitrev(cb):
    //    points shift * g_big^c * H  (radix_2_dit_parallel.rs:226-239).  The heaps of all cosets are one allocation
    //    (block cb at offset cb*h) so that the cosets run as grid.y of a single launch and share the coefficient reads in L2.
    const uint2 *tw = nullptr;
    P3_TRY(get_twiddles<F>(ctx, log_n, (int)added_bits, shift, 0, &tw));
    if (bitrev_rows) {
        P3_TRY(run_network<F>(ctx, log_n, w, tw, h, (unsigned)n_cosets, (const u32 *)coef, 0, 1, d_out, h * w, 0, 0, 0, nullptr, false,
                              make_uint2(0, 0), true));
    } else {
        void *tmp = nullptr;
        P3_TRY(ctx_scratch2(ctx, h * w * 4, &tmp));
        for (size_t cb = 0; cb < n_cosets; cb++) {
            size_t c = 0;
            for (unsigned b = 0; b < added_bits; b++) c |= ((cb >> b) & 1) << (added_bits - 1 - b);
            // natural LDE row of (coset c, evaluation index j) is j * n_cosets + c
            P3_TRY(run_network<F>(ctx, log_n, w, tw + cb * h, 0, 1, (const u32 *)coef, 0, 1, d_out, 0, 1, (int)added_bits, (u32)c,
                                  (u32 *)tmp, false, make_uint2(0, 0), true));
        }
    }
    return P3GPU_OK;
}

static int32_t check_shape(int field, size_t h, size_t w, unsigned extra_bits) {
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    P3_CHECK(w >= 1 && w < (1ull << 31), P3GPU_EINVAL, "matrix width %zu out of range", w);
    P3_CHECK(is_pow2(h), P3GPU_EINVAL, "matrix height %zu is not a power of two", h);
    const unsigned adicity = field == BABY_BEAR ? Fp<BABY_BEAR>::TWO_ADICITY : Fp<KOALA_BEAR>::TWO_ADICITY;
    P3_CHECK(log2_floor(h) + extra_bits <= adicity, P3GPU_EINVAL, "height 2^%u (+%u bits) exceeds the field's two-adicity %u",
             log2_floor(h), extra_bits, adicity);
    P3_CHECK((h << extra_bits) * w < (1ull << 40), P3GPU_EINVAL, "matrix too large");
    return P3GPU_OK;
}

int32_t ntt_dft_batch(p3gpu_ctx *ctx, int field, int kind, const u32 *d_in, u32 *d_out, size_t h, size_t w, u32 shift) {
    P3_TRY(check_shape(field, h, w, 0));
    P3_CHECK(kind >= P3GPU_DFT && kind <= P3GPU_COSET_IDFT, P3GPU_EINVAL, "unknown transform kind %d", kind);
    return field == BABY_BEAR ? dft_batch_impl<BABY_BEAR>(ctx, kind, d_in, d_out, h, w, shift)
                              : dft_batch_impl<KOALA_BEAR>(ctx, kind, d_in, d_out, h, w, shift);
}

int32_t ntt_coset_lde(p3gpu_ctx *ctx, int field, const u32 *d_in, size_t h, size_t w, unsigned added_bits, u32 shift, u32 *d_out,
                      int bitrev_rows) {
    P3_CHECK(added_bits <= 8, P3GPU_EINVAL, "added_bits %u too large", added_bits);
    P3_TRY(check_shape(field, h, w, added_bits));
    P3_CHECK(d_in != d_out, P3GPU_EINVAL, "coset_lde_batch cannot run in place");
    return field == BABY_BEAR ? coset_lde_impl<BABY_BEAR>(ctx, d_in, h, w, added_bits, shift, d_out, bitrev_rows)
                              : coset_lde_impl<KOALA_BEAR>(ctx, d_in, h, w, added_bits, shift, d_out, bitrev_rows);
}

}  // namespace p3
