// Poseidon2 AIR on the device: trace generation and quotient evaluation for
// VectorizedPoseidon2Air<KoalaBear, WIDTH 16, S-box degree 3, 0 S-box registers, 4 + 20 + 4 rounds, VECTOR_LEN permutations per row>
// — the AIR of `prove_prime_field_31 --field koala-bear --objective poseidon-2-permutations` (BASELINE config 5,
// examples/examples/prove_prime_field_31.rs:150-165).
//
//   trace generation   poseidon2-air/src/generation.rs:14-70,184-253: one permutation -> 164 columns
//                      inputs[16] | 4 x post[16] | rounds_p x post_sbox | 4 x post[16]; a vectorised row is VECTOR_LEN of them
//   constraints        poseidon2-air/src/air.rs:173-296 (one assert_eq per post / post_sbox column, all of degree 3, no selectors),
//                      vectorised poseidon2-air/src/vectorized.rs:297-311
//   quotient           uni-stark/src/prover.rs:462-827: fold the constraints with powers of alpha (the first asserted constraint
//                      gets the highest power, uni-stark/src/folder.rs), multiply by 1/Z_H (commit/src/domain.rs:321-361)
//
// Both kernels are integer bound like the leaf sponge (one Poseidon2 evaluation per permutation); the quotient kernel also streams
// the 11 GB trace LDE once.  Mapping of the quotient kernel: 8 lanes per LDE row (one per permutation of the row, each reading its
// 656 contiguous bytes), constraints folded with lazy 64-bit multiply-accumulates against an alpha-power table in shared memory,
// 3-step shuffle reduction across the row's lanes.
#include "common.h"
#include "hash_core.cuh"

namespace p3 {

struct AirConsts {
    u32 beg[64], end[64], part[32];
    int rounds_p;
};
static_assert(sizeof(AirConsts) <= 1024, "kernel parameter budget");

constexpr int AIR_W = 16;

template <int F> __device__ __forceinline__ void air_internal_layer(u32 (&s)[AIR_W]) {
    u32 part = s[1];
#pragma unroll
    for (int i = 2; i < AIR_W; i++) part = fp_add<F>(part, s[i]);
    const u32 sum = fp_add<F>(part, s[0]);
    s[0] = fp_sub<F>(part, s[0]);                       // V_0 = -2: -2 s0 + sum
    DiagLoop<F, AIR_W, 1>::run(s, sum);
}

// ---- trace generation: one thread per permutation, rows written through a per-warp shared-memory transpose ------------------
// A thread produces its permutation's 164 columns 16 (or rounds_p) at a time.  Written straight from the thread, a warp's store
// instruction touches 32 different rows 16 bytes each — partial sectors, and ncu shows the kernel stuck on the store queue
// (lg_throttle 34, 1.5 Gperm/s).  Instead every group of n values per permutation goes through a [32][n + 1] tile per warp and is
// written back with consecutive lanes on consecutive words of a row: whole 64/80-byte row segments per instruction.
template <int F>
__global__ void __launch_bounds__(128) p2air_generate_kernel(const u32 *inputs, size_t n_perms, u32 *trace, const __grid_constant__ AirConsts k) {
    __shared__ u32 tiles[4][32 * 33];
    u32 *tile = tiles[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) - lane;       // first permutation of this warp
    if (p0 >= n_perms) return;
    const size_t p = p0 + lane;
    const bool live = p < n_perms;
    const size_t cols = 144 + (size_t)k.rounds_p;
    const unsigned n_warp = (unsigned)min((size_t)32, n_perms - p0);
    // write `n` values per permutation (held as tile[perm * (n + 1) + i]) to columns [off, off + n) of the warp's rows
    auto flush = [&](unsigned n, size_t off) {
        __syncwarp();
        for (unsigned idx = lane; idx < n_warp * n; idx += 32) {
            const unsigned perm = idx / n, i = idx - perm * n;
            trace[(p0 + perm) * cols + off + i] = tile[perm * (n + 1) + i];
        }
        __syncwarp();
    };
    u32 s[AIR_W];
    if (live) {
        const uint4 *ip = reinterpret_cast<const uint4 *>(inputs + p * 16);
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint4 v = __ldg(ip + i); s[4 * i] = v.x; s[4 * i + 1] = v.y; s[4 * i + 2] = v.z; s[4 * i + 3] = v.w; }
    } else {
#pragma unroll
        for (int i = 0; i < AIR_W; i++) s[i] = 0;
    }
    auto put16 = [&](size_t off) {
#pragma unroll
        for (int i = 0; i < AIR_W; i++) tile[lane * 17 + i] = s[i];
        flush(16, off);
    };
    size_t off = 0;
    put16(off); off += 16;
    mds_light<F, AIR_W>(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < AIR_W; i++) s[i] = sbox<F>(fp_add<F>(s[i], k.beg[r * 16 + i]));
        mds_light<F, AIR_W>(s);
        put16(off); off += 16;
    }
    const unsigned rp = (unsigned)k.rounds_p;
#pragma unroll 1
    for (unsigned r = 0; r < rp; r++) {
        s[0] = sbox<F>(fp_add<F>(s[0], k.part[r]));
        tile[lane * (rp + 1) + r] = s[0];
        air_internal_layer<F>(s);
    }
    flush(rp, off); off += rp;
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < AIR_W; i++) s[i] = sbox<F>(fp_add<F>(s[i], k.end[r * 16 + i]));
        mds_light<F, AIR_W>(s);
        put16(off); off += 16;
    }
}

// ---- quotient ----------------------------------------------------------------------------------------------------------
template <int F> __device__ __forceinline__ void qmac(u64 (&acc)[4], u32 c, const uint4 a) {
    // acc += c * a (base x EF4), lazily: invariant acc < p * 2^32 (open.cu lazy_mac)
    const u32 av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int d = 0; d < 4; d++) {
        acc[d] += (u64)c * av[d];
        u32 hi = (u32)(acc[d] >> 32);
        const u32 hs = hi - Fp<F>::P;
        hi = hi < hs ? hi : hs;
        acc[d] = ((u64)hi << 32) | (u32)acc[d];
    }
}

struct QuotArgs {
    const u32 *lde;      // H x (vec_len * cols), bit-reversed rows
    u32 *q;              // H x 4, natural order over the quotient domain
    const u32 *apow;     // (vec_len * n_constraints) EF4: alpha^j
    const u32 *invz;     // 2^rate_bits inverse vanishing values
    unsigned log_h, rate_mask;
    int vec_len;
};

template <int F>
__global__ void __launch_bounds__(128) p2air_quotient_kernel(const QuotArgs a, const __grid_constant__ AirConsts k) {
    // alpha powers, one padded row per permutation of the vector: constraint kk of permutation v (global index j = v * nc + kk,
    // multiplied by alpha^(n_all - 1 - j)) sits at ap[v * (nc + 1) + kk].  The row stride of nc + 1 = 149 entries (596 words = 20
    // mod 32 banks) spreads the 8 permutations a warp works on over disjoint banks; without the pad they collide 4 ways (ncu).
    extern __shared__ uint4 ap[];
    const int nc = 128 + k.rounds_p, n_all = nc * a.vec_len;
    for (int t = threadIdx.x; t < n_all; t += blockDim.x) {
        const int j = n_all - 1 - t;
        ap[(j / nc) * (nc + 1) + (j % nc)] = __ldg(reinterpret_cast<const uint4 *>(a.apow) + t);
    }
    __syncthreads();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lanes = a.vec_len;                                    // power of two <= 32 (checked by the host)
    const size_t i = t / lanes;
    const int v = (int)(t % lanes);
    const bool live = i < ((size_t)1 << a.log_h);
    u64 acc[4] = {0, 0, 0, 0};
    if (live) {
        const size_t cols = 144 + (size_t)k.rounds_p;
        const size_t m = (size_t)(__brevll((unsigned long long)i) >> (64 - a.log_h));
        const u32 *c = a.lde + (m * lanes + v) * cols;
        const uint4 *apv = ap + v * (nc + 1);
        // a permutation's 164 columns start 16-byte aligned (656 = 41 x 16 bytes): 16-byte loads throughout
        const uint4 *c4 = reinterpret_cast<const uint4 *>(c);
        u32 s[AIR_W];
        auto ld16 = [&](u32 (&dst)[AIR_W]) {
#pragma unroll
            for (int x = 0; x < 4; x++) { const uint4 v4 = __ldg(c4 + x); dst[4 * x] = v4.x; dst[4 * x + 1] = v4.y; dst[4 * x + 2] = v4.z; dst[4 * x + 3] = v4.w; }
            c4 += 4;
        };
        ld16(s);
        mds_light<F, AIR_W>(s);
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int x = 0; x < AIR_W; x++) s[x] = sbox<F>(fp_add<F>(s[x], k.beg[r * 16 + x]));
            mds_light<F, AIR_W>(s);
            u32 post[AIR_W];
            ld16(post);
#pragma unroll
            for (int x = 0; x < AIR_W; x++) { qmac<F>(acc, fp_sub<F>(s[x], post[x]), apv[x]); s[x] = post[x]; }
            apv += 16;
        }
        {
            const u32 *cp = reinterpret_cast<const u32 *>(c4);
#pragma unroll 1
            for (int r = 0; r < k.rounds_p; r += 4) {           // rounds_p % 4 == 0 is checked by the host (20 for KoalaBear width 16)
                const uint4 v4 = __ldg(reinterpret_cast<const uint4 *>(cp + r));
                const u32 pv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int t2 = 0; t2 < 4; t2++) {
                    const u32 x3 = sbox<F>(fp_add<F>(s[0], k.part[r + t2]));
                    qmac<F>(acc, fp_sub<F>(x3, pv[t2]), *apv++);
                    s[0] = pv[t2];
                    air_internal_layer<F>(s);
                }
            }
            c4 = reinterpret_cast<const uint4 *>(cp + k.rounds_p);
        }
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int x = 0; x < AIR_W; x++) s[x] = sbox<F>(fp_add<F>(s[x], k.end[r * 16 + x]));
            mds_light<F, AIR_W>(s);
            u32 post[AIR_W];
            ld16(post);
#pragma unroll
            for (int x = 0; x < AIR_W; x++) { qmac<F>(acc, fp_sub<F>(s[x], post[x]), apv[x]); s[x] = post[x]; }
            apv += 16;
        }
    }
    u32 r[4];
#pragma unroll
    for (int d = 0; d < 4; d++) r[d] = mont_redc<F>(acc[d]);
    for (int off = 1; off < lanes; off <<= 1)
#pragma unroll
        for (int d = 0; d < 4; d++) r[d] = fp_add<F>(r[d], __shfl_xor_sync(0xffffffffu, r[d], off));
    if (live && v == 0) {
        const u32 z = __ldg(a.invz + (i & a.rate_mask));
        reinterpret_cast<uint4 *>(a.q)[i] = make_uint4(mont_mul<F>(r[0], z), mont_mul<F>(r[1], z), mont_mul<F>(r[2], z), mont_mul<F>(r[3], z));
    }
}

// alpha^j, j < n (sequential per block of 32 with a square-and-multiply start)
template <int F> __global__ void ef_powers_kernel(u32 *pw, size_t n, const Ef4<F> alpha) {
    const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t j0 = blk * 32;
    if (j0 >= n) return;
    Ef4<F> cur; cur.c[0] = Fp<F>::ONE; cur.c[1] = cur.c[2] = cur.c[3] = 0;
    Ef4<F> base = alpha;
    for (size_t e = j0; e; e >>= 1) { if (e & 1) cur = ef_mul<F>(cur, base); base = ef_mul<F>(base, base); }
    for (size_t j = j0; j < n && j < j0 + 32; j++) {
        reinterpret_cast<uint4 *>(pw)[j] = make_uint4(cur.c[0], cur.c[1], cur.c[2], cur.c[3]);
        cur = ef_mul<F>(cur, alpha);
    }
}

static int32_t air_consts(p3gpu_ctx *ctx, int field, const AirConsts **out) {
    P3_CHECK(field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "Poseidon2 AIR: only the KoalaBear instance (degree-3 S-box, no S-box registers) is built");
    P3_CHECK(ctx->air_set, P3GPU_ESTATE, "Poseidon2 AIR round constants not set (p3gpu_p2air_set_constants)");
    *out = reinterpret_cast<const AirConsts *>(ctx->air_consts);
    return P3GPU_OK;
}

int32_t air_set_constants(p3gpu_ctx *ctx, int field, const u32 *beg, const u32 *part, int rounds_p, const u32 *end) {
    P3_CHECK(field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "Poseidon2 AIR: only the KoalaBear instance (degree-3 S-box, no S-box registers) is built");
    P3_CHECK(rounds_p >= 4 && rounds_p <= 32 && rounds_p % 4 == 0, P3GPU_EINVAL, "rounds_p %d must be a multiple of 4 in 4..32", rounds_p);
    static_assert(sizeof(AirConsts) <= sizeof(ctx->air_consts), "context storage for the AIR constants");
    AirConsts k;
    memset(&k, 0, sizeof k);
    for (int i = 0; i < 64; i++) {
        P3_CHECK(beg[i] < Fp<KOALA_BEAR>::P && end[i] < Fp<KOALA_BEAR>::P, P3GPU_EINVAL, "round constant not in canonical Montgomery range");
        k.beg[i] = beg[i]; k.end[i] = end[i];
    }
    for (int i = 0; i < rounds_p; i++) { P3_CHECK(part[i] < Fp<KOALA_BEAR>::P, P3GPU_EINVAL, "round constant not in canonical Montgomery range"); k.part[i] = part[i]; }
    k.rounds_p = rounds_p;
    memcpy(ctx->air_consts, &k, sizeof k);
    ctx->air_set = 1;
    return P3GPU_OK;
}

int32_t air_generate_trace(p3gpu_ctx *ctx, int field, const u32 *d_inputs, size_t n_perms, u32 *d_trace) {
    const AirConsts *k;
    P3_TRY(air_consts(ctx, field, &k));
    if (n_perms == 0) return P3GPU_OK;
    p2air_generate_kernel<KOALA_BEAR><<<(unsigned)((n_perms + 127) / 128), 128, 0, ctx->stream>>>(d_inputs, n_perms, d_trace, *k);
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}

int32_t air_quotient(p3gpu_ctx *ctx, int field, int vec_len, const u32 *d_lde, unsigned log_h, unsigned log_n, const u32 *alpha, u32 *d_q) {
    constexpr int F = KOALA_BEAR;
    const AirConsts *k;
    P3_TRY(air_consts(ctx, field, &k));
    P3_CHECK(vec_len >= 1 && vec_len <= 32 && (vec_len & (vec_len - 1)) == 0, P3GPU_EINVAL, "vector length %d must be a power of two <= 32", vec_len);
    P3_CHECK(log_h >= log_n && log_h <= Fp<F>::TWO_ADICITY && log_h - log_n <= 8, P3GPU_EINVAL, "bad domain sizes 2^%u / 2^%u", log_h, log_n);
    P3_CHECK(reinterpret_cast<uintptr_t>(d_lde) % 16 == 0 && reinterpret_cast<uintptr_t>(d_q) % 16 == 0, P3GPU_EINVAL, "quotient: buffers must be 16-byte aligned");
    const int nc = 128 + k->rounds_p, n_all = nc * vec_len;
    const unsigned rate_bits = log_h - log_n;
    const size_t nz = (size_t)1 << rate_bits;
    void *tab = nullptr;
    P3_TRY(ctx_scratch2(ctx, (size_t)n_all * 16 + nz * 4, &tab));
    u32 *apow = (u32 *)tab, *invz = apow + (size_t)n_all * 4;
    Ef4<F> al; for (int d = 0; d < 4; d++) al.c[d] = alpha[d];
    ef_powers_kernel<F><<<(unsigned)(((n_all + 31) / 32 + 63) / 64), 64, 0, ctx->stream>>>(apow, (size_t)n_all, al);
    // 1 / Z_H on the coset GENERATOR * K (domain.rs:326-360): Z_H(x_i) = g^N * w^(i mod 2^rate_bits) - 1, w of order 2^rate_bits
    u32 hz[256];
    const u32 s_pow_n = fp_pow<F>(to_monty<F>(Fp<F>::GEN), (u64)1 << log_n), wr = two_adic_generator<F>(rate_bits);
    u32 wp = Fp<F>::ONE;
    for (size_t j = 0; j < nz; j++) { hz[j] = fp_inv<F>(fp_sub<F>(mont_mul<F>(s_pow_n, wp), Fp<F>::ONE)); wp = mont_mul<F>(wp, wr); }
    P3_CUDA(cudaMemcpyAsync(invz, hz, nz * 4, cudaMemcpyHostToDevice, ctx->stream));
    QuotArgs qa;
    qa.lde = d_lde; qa.q = d_q; qa.apow = apow; qa.invz = invz; qa.log_h = log_h; qa.rate_mask = (unsigned)(nz - 1); qa.vec_len = vec_len;
    const size_t threads = ((size_t)1 << log_h) * vec_len;
    const size_t smem = (size_t)vec_len * (nc + 1) * 16;
    auto kern = p2air_quotient_kernel<F>;
    if (smem > 48 * 1024) P3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)((threads + 127) / 128), 128, smem, ctx->stream>>>(qa, *k);
    ctx->launches += 2;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}

}  // namespace p3
