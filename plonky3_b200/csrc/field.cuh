// Device arithmetic for the two 31-bit Montgomery fields of the hot path.
//
// Data representation is identical to the reference's MontyField31.value (monty-31/src/monty_31.rs:34-44):
// u32 = x * 2^32 mod p, canonical range [0, p).  Field parameters: baby-bear/src/baby_bear.rs:14-65,
// koala-bear/src/koala_bear.rs:14-91.  add/sub/mul restate monty-31/src/utils.rs:63-125 with branch-free
// unsigned-min corrections.
//
// Two multiplication flavours are provided:
//   * mont_mul(a, b)        general product of two Montgomery values (S-boxes, EF4 arithmetic)
//   * shoup_mul(v, {w, w'}) product by a *precomputed constant* w (canonical integer, w' = floor(w*2^32/p)):
//                           3 multiply-class instructions, no carries, accepts ANY u32 v and returns a value
//                           in [0, 2p).  Because w is the canonical value, v*w keeps v's Montgomery scaling,
//                           so twiddle tables are stored canonical and the data never leaves Montgomery form.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace p3 {

typedef uint32_t u32;
typedef uint64_t u64;

enum { BABY_BEAR = 0, KOALA_BEAR = 1 };

template <int FIELD> struct Fp;

template <> struct Fp<BABY_BEAR> {
    static constexpr u32 P = 0x78000001u;
    static constexpr u32 MU = 0x88000001u;        // p^-1 mod 2^32 (baby_bear.rs:20)
    static constexpr u32 ONE = 0x0ffffffeu;       // 2^32 mod p
    static constexpr u32 R2 = 0x45dddde3u;        // 2^64 mod p
    static constexpr u32 GEN = 31u;               // baby_bear.rs:28
    static constexpr u32 TWO_ADICITY = 27u;       // baby_bear.rs:44
    static constexpr u32 TOP_ROOT = 0x1a427a41u;  // canonical generator of the 2^27 subgroup (baby_bear.rs:48-53)
    static constexpr u32 EXT_W = 11u;             // baby_bear.rs:68
    static constexpr int SBOX_D = 7;              // baby-bear/src/poseidon1.rs:38
};
template <> struct Fp<KOALA_BEAR> {
    static constexpr u32 P = 0x7f000001u;
    static constexpr u32 MU = 0x81000001u;        // koala_bear.rs:23
    static constexpr u32 ONE = 0x01fffffeu;
    static constexpr u32 R2 = 0x17f7efe4u;
    static constexpr u32 GEN = 3u;                // koala_bear.rs:53
    static constexpr u32 TWO_ADICITY = 24u;       // koala_bear.rs:69
    static constexpr u32 TOP_ROOT = 0x6ac49f88u;  // koala_bear.rs:73-78
    static constexpr u32 EXT_W = 3u;              // koala_bear.rs:94
    static constexpr int SBOX_D = 3;              // koala-bear/src/poseidon1.rs:27
};

// ---- canonical-range helpers ---------------------------------------------------------------
template <int F> __host__ __device__ __forceinline__ u32 fp_reduce(u32 x) {  // [0,2p) -> [0,p)
    u32 y = x - Fp<F>::P;
    return x < y ? x : y;  // unsigned min: if x < p the subtraction wraps to a huge value
}
template <int F> __host__ __device__ __forceinline__ u32 fp_add(u32 a, u32 b) { return fp_reduce<F>(a + b); }
template <int F> __host__ __device__ __forceinline__ u32 fp_sub(u32 a, u32 b) {
    u32 d = a - b, e = d + Fp<F>::P;
    return d < e ? d : e;  // a>=b: d in [0,p) < d+p ; a<b: d wrapped (huge), d+p is the answer
}
template <int F> __host__ __device__ __forceinline__ u32 fp_neg(u32 a) { return fp_sub<F>(0u, a); }
template <int F> __host__ __device__ __forceinline__ u32 fp_double(u32 a) { return fp_reduce<F>(a + a); }
// monty-31/src/utils.rs:92-97
template <int F> __host__ __device__ __forceinline__ u32 fp_halve(u32 a) {
    return (a >> 1) + ((a & 1u) ? ((Fp<F>::P + 1u) >> 1) : 0u);
}

// Montgomery reduction of x < p*2^32 to the signed-wrapped value hi(x) - hi(t*p) in (-p, p)  (utils.rs:105-125)
template <int F> __host__ __device__ __forceinline__ u32 mont_redc_lazy(u64 x) {
    u32 t = (u32)x * Fp<F>::MU;
#ifdef __CUDA_ARCH__
    u32 u = __umulhi(t, Fp<F>::P);
#else
    u32 u = (u32)(((u64)t * Fp<F>::P) >> 32);
#endif
    return (u32)(x >> 32) - u;
}
template <int F> __host__ __device__ __forceinline__ u32 mont_redc(u64 x) {
    u32 r = mont_redc_lazy<F>(x), s = r + Fp<F>::P;
    return r < s ? r : s;
}
template <int F> __host__ __device__ __forceinline__ u32 mont_mul(u32 a, u32 b) { return mont_redc<F>((u64)a * b); }
template <int F> __host__ __device__ __forceinline__ u32 to_monty(u32 canonical) { return mont_mul<F>(canonical, Fp<F>::R2); }
template <int F> __host__ __device__ __forceinline__ u32 from_monty(u32 m) { return mont_redc<F>((u64)m); }

template <int F> __host__ __device__ inline u32 fp_pow(u32 a, u64 e) {
    u32 r = Fp<F>::ONE;
    while (e) { if (e & 1) r = mont_mul<F>(r, a); a = mont_mul<F>(a, a); e >>= 1; }
    return r;
}
template <int F> __host__ __device__ inline u32 fp_inv(u32 a) { return fp_pow<F>(a, (u64)Fp<F>::P - 2); }
// monty_31.rs:709-726
template <int F> __host__ __device__ inline u32 two_adic_generator(u32 bits) {
    u32 g = to_monty<F>(Fp<F>::TOP_ROOT);
    for (u32 i = bits; i < Fp<F>::TWO_ADICITY; i++) g = mont_mul<F>(g, g);
    return g;
}

// ---- Shoup constant multiplication ---------------------------------------------------------
// tw.x = w (canonical, < p), tw.y = floor(w * 2^32 / p).  Result in [0, 2p) for any u32 v.
template <int F> __device__ __forceinline__ u32 shoup_mul(u32 v, uint2 tw) {
    u32 q = __umulhi(v, tw.y);
    return v * tw.x - q * Fp<F>::P;
}
template <int F> __host__ __device__ inline uint2 shoup_pair(u32 w_canonical) {
    uint2 r;
    r.x = w_canonical;
    r.y = (u32)((((u64)w_canonical) << 32) / Fp<F>::P);
    return r;
}

// Cooley-Tukey butterfly on lazily reduced data: a, b in [0, 2p) (b may be any u32), outputs in [0, 2p).
//   a' = a + w*b ,  b' = a - w*b          (dft/src/butterflies.rs:118 DitButterfly)
template <int F> __device__ __forceinline__ void ct_butterfly(u32 &a, u32 &b, uint2 tw) {
    u32 u = fp_reduce<F>(a);
    u32 r = fp_reduce<F>(shoup_mul<F>(b, tw));
    a = u + r;
    b = u - r + Fp<F>::P;
}

// ---- EF4 = F[X]/(X^4 - W)  (field/src/extension/binomial_extension.rs:724-770) --------------
template <int F> struct Ef4 { u32 c[4]; };

template <int F> __host__ __device__ __forceinline__ u32 mul_w(u32 a) {
    if (F == KOALA_BEAR) return fp_add<F>(fp_double<F>(a), a);  // W = 3 (koala_bear.rs:97-99)
    return mont_mul<F>(a, to_monty<F>(Fp<F>::EXT_W));
}
template <int F> __host__ __device__ inline Ef4<F> ef_mul(const Ef4<F> &a, const Ef4<F> &b) {
    // schoolbook with u64 accumulation of Montgomery products: each product < p^2 < 2^62, so sums of 4 overflow;
    // reduce every product instead (this is not a hot loop: one EF mul per folded element).
    u32 r[7] = {0, 0, 0, 0, 0, 0, 0};
    #pragma unroll
    for (int i = 0; i < 4; i++)
        #pragma unroll
        for (int j = 0; j < 4; j++) r[i + j] = fp_add<F>(r[i + j], mont_mul<F>(a.c[i], b.c[j]));
    Ef4<F> o;
    #pragma unroll
    for (int i = 0; i < 3; i++) o.c[i] = fp_add<F>(r[i], mul_w<F>(r[i + 4]));
    o.c[3] = r[3];
    return o;
}

}  // namespace p3
