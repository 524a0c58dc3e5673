// Host execution of the DEVICE arithmetic in plonky3_b200/csrc/field.cuh (compiled as plain C++: g++ ignores the CUDA function
// attributes, the one intrinsic gets a host body).  Checks the lazy-range contracts the NTT kernels rely on against 64-bit
// reference arithmetic: Montgomery multiply, Shoup multiply by a constant for ANY 32-bit input, and the Cooley-Tukey butterfly on
// [0, 2p) data.  Prints "ok <n>" or the first violation.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
#include "../../plonky3_b200/csrc/field.cuh"
using namespace p3;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <int F> static int run(const char *name) {
    const uint64_t P = Fp<F>::P;
    const uint64_t RINV = [] { uint64_t r = 1, b = ((uint64_t)1 << 32) % Fp<F>::P, e = Fp<F>::P - 2; while (e) { if (e & 1) r = r * b % Fp<F>::P; b = b * b % Fp<F>::P; e >>= 1; } return r; }();
    long n = 0;
    for (int it = 0; it < 400000; it++) {
        const u32 a = (u32)(rnd() % P), b = (u32)(rnd() % P);
        // Montgomery product: a*b*R^-1 mod p, canonical
        if (mont_mul<F>(a, b) != (u32)((uint64_t)a * b % P * RINV % P)) { printf("%s mont_mul(%u,%u)\n", name, a, b); return 1; }
        if (from_monty<F>(to_monty<F>(a)) != a) { printf("%s to/from monty %u\n", name, a); return 1; }
        // Shoup: any u32 v, canonical constant w -> v*w mod p in [0, 2p)
        const u32 v = (u32)rnd(), w = (u32)(rnd() % P);
        const u32 s = shoup_mul<F>(v, shoup_pair<F>(w));
        if (s >= 2 * P || s % P != (uint64_t)v * w % P) { printf("%s shoup_mul(%u,%u) = %u\n", name, v, w, s); return 1; }
        // butterfly on lazy data: x in [0,2p), y any u32 (the kernels feed [0,2p)); outputs in [0,2p), congruent to x +- w*y
        u32 x = (u32)(rnd() % (2 * P)), y = (it & 1) ? (u32)(rnd() % (2 * P)) : (u32)rnd();
        const uint64_t wy = (uint64_t)y % P * w % P;
        u32 x2 = x, y2 = y;
        ct_butterfly<F>(x2, y2, shoup_pair<F>(w));
        if (x2 >= 2 * P || y2 >= 2 * P || x2 % P != (x % P + wy) % P || y2 % P != (x % P + P - wy) % P) {
            printf("%s butterfly(%u,%u,w=%u) = %u,%u\n", name, x, y, w, x2, y2); return 1;
        }
        // canonical helpers
        if (fp_add<F>(a, b) != (a + (uint64_t)b) % P || fp_sub<F>(a, b) != (a + P - b) % P || fp_halve<F>(fp_double<F>(a)) != a) { printf("%s add/sub/halve\n", name); return 1; }
        n += 5;
    }
    // EF4 multiplication against schoolbook mod (X^4 - W) on canonical integers
    for (int it = 0; it < 20000; it++) {
        Ef4<F> a, b; uint64_t ca[4], cb[4], acc[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 4; k++) { ca[k] = rnd() % P; cb[k] = rnd() % P; a.c[k] = to_monty<F>((u32)ca[k]); b.c[k] = to_monty<F>((u32)cb[k]); }
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i + j] = (acc[i + j] + ca[i] * cb[j]) % P;
        const Ef4<F> r = ef_mul<F>(a, b);
        for (int k = 0; k < 4; k++) {
            const uint64_t want = (acc[k] + (k < 3 ? acc[k + 4] * Fp<F>::EXT_W % P : 0)) % P;
            if (from_monty<F>(r.c[k]) != want) { printf("%s ef_mul coefficient %d\n", name, k); return 1; }
        }
        n++;
    }
    // two-adic generators: order exactly 2^bits
    for (u32 bits = 1; bits <= Fp<F>::TWO_ADICITY; bits++) {
        u32 g = two_adic_generator<F>(bits), h = g;
        for (u32 i = 1; i < bits; i++) h = mont_mul<F>(h, h);
        if (h == Fp<F>::ONE || mont_mul<F>(h, h) != Fp<F>::ONE) { printf("%s two_adic_generator(%u)\n", name, bits); return 1; }
        n++;
    }
    printf("ok %s %ld\n", name, n);
    return 0;
}

int main() { return run<BABY_BEAR>("baby_bear") | run<KOALA_BEAR>("koala_bear"); }
