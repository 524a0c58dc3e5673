// The two permutations of the hot path as device functions: Poseidon2 over the 31-bit Montgomery fields (width 16 / 24) and
// Keccak-f[1600].  Kept apart from the kernels (hash.cu) so that the same source can also be compiled as plain C++ and executed
// on the host against the CPU oracle (tests/cpp/hash_core_host.cpp): g++ ignores the CUDA attributes, and the one intrinsic used
// here gets a host body there.
#pragma once
#include "field.cuh"
#include "poseidon2_consts.h"

namespace p3 {

// =================================================================================================
// Poseidon2
// =================================================================================================
// Montgomery product left in (0, 2p): hi(ab) - hi(t p) + p is one IADD3, no conditional correction.  Safe as ONE factor of
// a following product (2p * p < p * 2^32), which then returns to the canonical range.
template <int F> __device__ __forceinline__ u32 mont_mul_lazy(u32 a, u32 b) { return mont_redc_lazy<F>((u64)a * b) + Fp<F>::P; }

template <int F> __device__ __forceinline__ u32 sbox(u32 x) {
    if (Fp<F>::SBOX_D == 3) return mont_mul<F>(mont_mul_lazy<F>(x, x), x);          // x^3: the square stays lazy
    const u32 x2 = mont_mul<F>(x, x);                                                // x^7 = x^4 * x^3, x^3 lazy
    const u32 x3 = mont_mul_lazy<F>(x2, x);
    const u32 x4 = mont_mul<F>(x2, x2);
    return mont_mul<F>(x4, x3);
}

// poseidon2/src/external.rs:60-74: circ(2,3,1,1)
template <int F> __device__ __forceinline__ void mat4(u32 &x0, u32 &x1, u32 &x2, u32 &x3) {
    const u32 t01 = fp_add<F>(x0, x1), t23 = fp_add<F>(x2, x3);
    const u32 t0123 = fp_add<F>(t01, t23);
    const u32 t01123 = fp_add<F>(t0123, x1), t01233 = fp_add<F>(t0123, x3);
    const u32 n3 = fp_add<F>(t01233, fp_double<F>(x0));
    const u32 n1 = fp_add<F>(t01123, fp_double<F>(x2));
    const u32 n0 = fp_add<F>(t01123, t01);
    const u32 n2 = fp_add<F>(t01233, t23);
    x0 = n0; x1 = n1; x2 = n2; x3 = n3;
}
// poseidon2/src/external.rs:113-159
template <int F, int W> __device__ __forceinline__ void mds_light(u32 (&s)[W]) {
#pragma unroll
    for (int i = 0; i < W; i += 4) mat4<F>(s[i], s[i + 1], s[i + 2], s[i + 3]);
    u32 sums[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        sums[k] = s[k];
#pragma unroll
        for (int j = 4; j < W; j += 4) sums[k] = fp_add<F>(sums[k], s[j + k]);
    }
#pragma unroll
    for (int i = 0; i < W; i++) s[i] = fp_add<F>(s[i], sums[i & 3]);
}

// Internal diagonal V (1 + Diag(V) is the internal matrix): koala-bear/src/poseidon2.rs:407-461,
// baby-bear/src/poseidon2.rs:394-450.  Encoded as (mul, shift): V_i = mul * 2^shift, shift <= 0.
struct DiagEntry { int mul, shift; };
template <int F, int W> struct Diag;
template <> struct Diag<BABY_BEAR, 16> { static __host__ __device__ constexpr DiagEntry at(int i) {
    constexpr DiagEntry d[16] = {{-2,0},{1,0},{2,0},{1,-1},{3,0},{4,0},{-1,-1},{-3,0},{-4,0},{1,-8},{1,-2},{1,-3},{1,-27},{-1,-8},{-1,-4},{-1,-27}};
    return d[i]; } };
template <> struct Diag<BABY_BEAR, 24> { static __host__ __device__ constexpr DiagEntry at(int i) {
    constexpr DiagEntry d[24] = {{-2,0},{1,0},{2,0},{1,-1},{3,0},{4,0},{-1,-1},{-3,0},{-4,0},{1,-8},{1,-2},{1,-3},{1,-4},{1,-7},{1,-9},{1,-27},{-1,-8},{-1,-2},{-1,-3},{-1,-4},{-1,-5},{-1,-6},{-1,-7},{-1,-27}};
    return d[i]; } };
template <> struct Diag<KOALA_BEAR, 16> { static __host__ __device__ constexpr DiagEntry at(int i) {
    constexpr DiagEntry d[16] = {{-2,0},{1,0},{2,0},{1,-1},{3,0},{4,0},{-1,-1},{-3,0},{-4,0},{1,-8},{1,-3},{1,-24},{-1,-8},{-1,-3},{-1,-4},{-1,-24}};
    return d[i]; } };
template <> struct Diag<KOALA_BEAR, 24> { static __host__ __device__ constexpr DiagEntry at(int i) {
    constexpr DiagEntry d[24] = {{-2,0},{1,0},{2,0},{1,-1},{3,0},{4,0},{-1,-1},{-3,0},{-4,0},{1,-8},{1,-2},{1,-3},{1,-4},{1,-5},{1,-6},{1,-24},{-1,-8},{-1,-3},{-1,-4},{-1,-5},{-1,-6},{-1,-7},{-1,-9},{-1,-24}};
    return d[i]; } };

// x * 2^-k (monty-31 div_2exp_u64).  Both primes are p = 2^31 - 2^L + 1 (L = 24 KoalaBear, 27 BabyBear), so p = 1 mod 2^k for
// k <= L and the exact quotient is (x + m*p) >> k with m = (-x) mod 2^k:
//     x / 2^k = ceil(x / 2^k) + m * (2^(31-k) - 2^(L-k))            (result < p for x < p)
// = 2 logic/shift ops + 1 add + 1 IMAD, no IMAD.HI (a Montgomery reduction of x << (32-k) costs IMAD + IMAD.HI + 4 ALU ops).
// It is representation-independent: dividing the Montgomery form by 2^k divides the value by 2^k.
template <int F, int K> __device__ __forceinline__ u32 div_2exp(u32 x) {
    constexpr int L = (F == KOALA_BEAR) ? 24 : 27;
    static_assert(K >= 1 && K <= L, "shift exceeds the 2-adic part of p - 1");
    constexpr u32 mask = (1u << K) - 1u;
    constexpr u32 C = (1u << (31 - K)) - (1u << (L - K));
    const u32 m = (0u - x) & mask;
    const u32 c = (x + mask) >> K;
    return c + m * C;
}
template <int F, int W, int I> __device__ __forceinline__ u32 diag_mul_add(u32 x, u32 sum) {
    constexpr DiagEntry d = Diag<F, W>::at(I);
    constexpr int am = d.mul < 0 ? -d.mul : d.mul;
    u32 v;
    if constexpr (d.shift == 0) {
        v = x;
        if (am == 2) v = fp_double<F>(x);
        if (am == 3) v = fp_add<F>(fp_double<F>(x), x);
        if (am == 4) v = fp_double<F>(fp_double<F>(x));
    } else {
        v = div_2exp<F, -d.shift>(x);   // also covers the halves (k = 1)
    }
    return d.mul < 0 ? fp_sub<F>(sum, v) : fp_add<F>(sum, v);
}
template <int F, int W, int I> struct DiagLoop {
    static __device__ __forceinline__ void run(u32 (&s)[W], u32 sum) {
        s[I] = diag_mul_add<F, W, I>(s[I], sum);
        DiagLoop<F, W, I + 1>::run(s, sum);
    }
};
template <int F, int W> struct DiagLoop<F, W, W> { static __device__ __forceinline__ void run(u32 (&)[W], u32) {} };

// One copy of the external-round body and one of the internal-round body (rounds are loops, not unrolled): the fully
// unrolled permutation is 50-300 KB of SASS and the leaf kernels then stall on instruction fetch (ncu: stall_no_instruction
// was the top reason); looped, a whole sponge kernel is 10-20 KB and stays resident in the instruction cache.
template <int F, int W>
__device__ __forceinline__ void poseidon2_permute(u32 (&s)[W], const Poseidon2Consts &k) {
    mds_light<F, W>(s);
#pragma unroll 1
    for (int r = 0; r < 8; r++) {
        if (r == 4) {
            // monty-31/src/poseidon2.rs:76-85
#pragma unroll 1
            for (int q = 0; q < k.rounds_p; q++) {
                s[0] = sbox<F>(fp_add<F>(s[0], k.rc_int[q]));
                u32 part = s[1];
#pragma unroll
                for (int i = 2; i < W; i++) part = fp_add<F>(part, s[i]);
                const u32 sum = fp_add<F>(part, s[0]);
                s[0] = fp_sub<F>(part, s[0]);
                DiagLoop<F, W, 1>::run(s, sum);
            }
        }
        // external round r (0-3 initial, 4-7 terminal): poseidon2/src/external.rs:288-336
#pragma unroll
        for (int i = 0; i < W; i++) s[i] = sbox<F>(fp_add<F>(s[i], k.rc_ext[r * W + i]));
        mds_light<F, W>(s);
    }
}

// =================================================================================================
// Keccak-f[1600]
// =================================================================================================
// The state is kept as 25 (lo, hi) pairs of 32-bit registers: every 64-bit rotation is two funnel shifts (SHF), theta's
// column parity + application and chi are single 3-input LOP3s, so a round is exactly 122 LOP3 + 58 SHF and no register
// moves (the compiler's 64-bit version needed 162 LOP3 + 52 SHF + 46 moves).  All of it runs on the ALU pipe: the kernel is
// bound by that pipe (ncu: 98 % busy).  The (lo, hi) split also matches the leaf packing, which pairs consecutive u32
// field elements into one u64 word (field/src/integers.rs:494-509): lo = first element, hi = second.
static __constant__ u32 KECCAK_RC_LO[24] = {0x00000001u, 0x00008082u, 0x0000808au, 0x80008000u, 0x0000808bu, 0x80000001u, 0x80008081u, 0x00008009u,
                                      0x0000008au, 0x00000088u, 0x80008009u, 0x8000000au, 0x8000808bu, 0x0000008bu, 0x00008089u, 0x00008003u,
                                      0x00008002u, 0x00000080u, 0x0000800au, 0x8000000au, 0x80008081u, 0x00008080u, 0x80000001u, 0x80008008u};
static __constant__ u32 KECCAK_RC_HI[24] = {0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u,
                                      0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x80000000u,
                                      0x80000000u, 0x80000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x00000000u, 0x80000000u};

// 64-bit rotate-left of (lo, hi) by the compile-time constant R
template <int R> __device__ __forceinline__ void rot64(u32 lo, u32 hi, u32 &olo, u32 &ohi) {
    if constexpr (R == 0) { olo = lo; ohi = hi; }
    else if constexpr (R == 32) { olo = hi; ohi = lo; }
    else if constexpr (R < 32) { olo = __funnelshift_l(hi, lo, R); ohi = __funnelshift_l(lo, hi, R); }
    else { olo = __funnelshift_l(lo, hi, R - 32); ohi = __funnelshift_l(hi, lo, R - 32); }
}

struct KState { u32 lo[25], hi[25]; };

__device__ __forceinline__ void keccak_f(KState &s) {
    u32 (&al)[25] = s.lo;
    u32 (&ah)[25] = s.hi;
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        u32 cl[5], ch[5], rl[5], rh[5];
#pragma unroll
        for (int x = 0; x < 5; x++) {
            cl[x] = al[x] ^ al[x + 5] ^ al[x + 10] ^ al[x + 15] ^ al[x + 20];
            ch[x] = ah[x] ^ ah[x + 5] ^ ah[x + 10] ^ ah[x + 15] ^ ah[x + 20];
        }
#pragma unroll
        for (int x = 0; x < 5; x++) rot64<1>(cl[x], ch[x], rl[x], rh[x]);
        u32 bl[25], bh[25];
        // theta + rho + pi: b[pi(i)] = rotl(a[i] ^ C[x-1] ^ rotl(C[x+1], 1), rho(i))
#define P3_TH(i, R, dst)                                                                             \
    rot64<R>(al[i] ^ cl[((i) % 5 + 4) % 5] ^ rl[((i) % 5 + 1) % 5], ah[i] ^ ch[((i) % 5 + 4) % 5] ^ rh[((i) % 5 + 1) % 5], \
             bl[dst], bh[dst])
        P3_TH(0, 0, 0);
        P3_TH(1, 1, 10);   P3_TH(2, 62, 20);  P3_TH(3, 28, 5);   P3_TH(4, 27, 15);
        P3_TH(5, 36, 16);  P3_TH(6, 44, 1);   P3_TH(7, 6, 11);   P3_TH(8, 55, 21);
        P3_TH(9, 20, 6);   P3_TH(10, 3, 7);   P3_TH(11, 10, 17); P3_TH(12, 43, 2);
        P3_TH(13, 25, 12); P3_TH(14, 39, 22); P3_TH(15, 41, 23); P3_TH(16, 45, 8);
        P3_TH(17, 15, 18); P3_TH(18, 21, 3);  P3_TH(19, 8, 13);  P3_TH(20, 18, 14);
        P3_TH(21, 2, 24);  P3_TH(22, 61, 9);  P3_TH(23, 56, 19); P3_TH(24, 14, 4);
#undef P3_TH
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
            for (int x = 0; x < 5; x++) {
                al[x + 5 * y] = bl[x + 5 * y] ^ (~bl[(x + 1) % 5 + 5 * y] & bl[(x + 2) % 5 + 5 * y]);
                ah[x + 5 * y] = bh[x + 5 * y] ^ (~bh[(x + 1) % 5 + 5 * y] & bh[(x + 2) % 5 + 5 * y]);
            }
        al[0] ^= KECCAK_RC_LO[round];
        ah[0] ^= KECCAK_RC_HI[round];
    }
}

}  // namespace p3
