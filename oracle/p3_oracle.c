/*
 * p3_oracle.c — CPU restatement of the Plonky3 prover hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for the B200 kernels: a plain-C restatement of the reference
 * algorithms.  It is NOT product code.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product path (plonky3_b200/) never does.
 *
 * The reference is 100% Rust and no Rust toolchain exists in this image, so the reference itself
 * cannot be compiled here (oracle/_ref is therefore absent, cpu_baseline.kind = "port").
 * The oracle is pinned against the reference's own known-answer tests and its committed proof
 * fixture (tests/test_oracle_*.py):  field KATs, two-adic generator tower, Poseidon2 KATs for both
 * fields / widths 16+24, and a full replay of uni-stark/tests/fixtures/uni_stark_two_adic_v1.postcard
 * (trace cap, quotient cap, openings, FRI round cap, final poly).  The Keccak-f path has no literal
 * vector in the reference ("parity unpinned" there); it is pinned with FIPS-202 vectors (hashlib).
 *
 * Every function cites the reference file:line it restates (paths relative to the Plonky3 tree).
 * All field elements are u32 in Montgomery form x*2^32 mod p, exactly MontyField31.value
 * (monty-31/src/monty_31.rs:34-44).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef uint32_t u32;
typedef uint64_t u64;

/* ------------------------------------------------------------------------------------------------
 * Field parameters: baby-bear/src/baby_bear.rs:14-65, koala-bear/src/koala_bear.rs:14-91
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    u32 p;          /* prime */
    u32 mu;         /* p^-1 mod 2^32 (positive convention, monty-31/src/utils.rs:105-125) */
    u32 gen;        /* multiplicative generator (canonical) */
    u32 two_adicity;
    u32 top_root;   /* generator of the 2^two_adicity subgroup (canonical) */
    u32 ext_w;      /* X^4 - W (canonical) */
    u32 sbox_d;     /* Poseidon2 S-box degree */
} field_t;

static const field_t FIELDS[2] = {
    /* BabyBear  */ {0x78000001u, 0x88000001u, 31u, 27u, 0x1a427a41u, 11u, 7u},
    /* KoalaBear */ {0x7f000001u, 0x81000001u, 3u, 24u, 0x6ac49f88u, 3u, 3u},
};

static inline const field_t *F(int f) { return &FIELDS[f]; }

/* monty-31/src/utils.rs:63-71 */
static inline u32 f_add(const field_t *f, u32 a, u32 b) { u32 s = a + b; return s >= f->p ? s - f->p : s; }
/* monty-31/src/utils.rs:81-86 */
static inline u32 f_sub(const field_t *f, u32 a, u32 b) { u32 d = a - b; return a < b ? d + f->p : d; }
/* monty-31/src/utils.rs:105-125 (monty_reduce), monty_31.rs:757-764 (Mul) */
static inline u32 f_redc(const field_t *f, u64 x) {
    u64 t = (u32)((u32)x * f->mu);
    u64 u = t * (u64)f->p;
    u64 d = x - u;
    u32 hi = (u32)(d >> 32);
    return x < u ? hi + f->p : hi;
}
static inline u32 f_mul(const field_t *f, u32 a, u32 b) { return f_redc(f, (u64)a * b); }
/* monty-31/src/utils.rs:7-9 */
static inline u32 f_to_monty(const field_t *f, u32 x) { return (u32)((((u64)x) << 32) % f->p); }
static inline u32 f_from_monty(const field_t *f, u32 x) { return f_redc(f, x); }
static inline u32 f_one(const field_t *f) { return f_to_monty(f, 1); }
/* monty-31/src/utils.rs:92-97 */
static inline u32 f_halve(const field_t *f, u32 a) { u32 s = a >> 1; return (a & 1) ? s + ((f->p + 1) >> 1) : s; }

static u32 f_pow(const field_t *f, u32 a, u64 e) {
    u32 r = f_one(f);
    while (e) { if (e & 1) r = f_mul(f, r, a); a = f_mul(f, a, a); e >>= 1; }
    return r;
}
static inline u32 f_inv(const field_t *f, u32 a) { return f_pow(f, a, (u64)f->p - 2); }

/* monty_31.rs:709-726: two_adic_generator(bits) = top_root^(2^(two_adicity-bits)) */
static u32 f_two_adic_generator(const field_t *f, u32 bits) {
    u32 g = f_to_monty(f, f->top_root);
    for (u32 i = bits; i < f->two_adicity; i++) g = f_mul(f, g, g);
    return g;
}

/* exported scalar API (all Montgomery in / Montgomery out unless stated) */
u32 p3o_prime(int f) { return F(f)->p; }
u32 p3o_add(int f, u32 a, u32 b) { return f_add(F(f), a, b); }
u32 p3o_sub(int f, u32 a, u32 b) { return f_sub(F(f), a, b); }
u32 p3o_mul(int f, u32 a, u32 b) { return f_mul(F(f), a, b); }
u32 p3o_pow(int f, u32 a, u64 e) { return f_pow(F(f), a, e); }
u32 p3o_inv(int f, u32 a) { return f_inv(F(f), a); }
u32 p3o_halve(int f, u32 a) { return f_halve(F(f), a); }
u32 p3o_to_monty(int f, u32 x) { return f_to_monty(F(f), x % F(f)->p); }
u32 p3o_from_monty(int f, u32 x) { return f_from_monty(F(f), x); }
u32 p3o_two_adic_generator(int f, u32 bits) { return f_two_adic_generator(F(f), bits); }
u32 p3o_generator(int f) { return f_to_monty(F(f), F(f)->gen); }
void p3o_to_monty_vec(int f, u32 *v, size_t n) { for (size_t i = 0; i < n; i++) v[i] = f_to_monty(F(f), v[i] % F(f)->p); }
void p3o_from_monty_vec(int f, u32 *v, size_t n) { for (size_t i = 0; i < n; i++) v[i] = f_from_monty(F(f), v[i]); }

/* ------------------------------------------------------------------------------------------------
 * Bit reversal: util/src/lib.rs:203-214 (reverse_bits_len), matrix/src/util.rs:36-57
 * ---------------------------------------------------------------------------------------------- */
static inline size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static unsigned log2_strict(size_t n) {
    unsigned l = 0;
    while (((size_t)1 << l) < n) l++;
    if (((size_t)1 << l) != n) { fprintf(stderr, "p3_oracle: %zu is not a power of two\n", n); abort(); }
    return l;
}
/* memcpy and geometric tables split over the OpenMP team (large matrices: a serial copy / first touch of ~0.4 GB costs as much
 * as a butterfly sweep, and the pages should be touched by the threads that use them) */
static void par_copy(u32 *dst, const u32 *src, size_t n) {
    #pragma omp parallel for schedule(static)
    for (size_t blk = 0; blk < (n + 65535) / 65536; blk++) {
        size_t o = blk * 65536, len = n - o < 65536 ? n - o : 65536;
        memcpy(dst + o, src + o, len * 4);
    }
}
static void f_powers(const field_t *f, u32 *pw, size_t n, u32 base) {   /* pw[i] = base^i */
    const size_t B = 4096;
    #pragma omp parallel for schedule(static)
    for (size_t blk = 0; blk < (n + B - 1) / B; blk++) {
        size_t o = blk * B, e = o + B < n ? o + B : n;
        u32 v = f_pow(f, base, o);
        for (size_t i = o; i < e; i++) { pw[i] = v; v = f_mul(f, v, base); }
    }
}
void p3o_reverse_matrix_index_bits(u32 *mat, size_t h, size_t w) {
    unsigned lh = log2_strict(h);
    #pragma omp parallel
    {
        u32 *tmp = (u32 *)malloc(w * sizeof(u32));
        #pragma omp for schedule(static)
        for (size_t i = 0; i < h; i++) {
            size_t j = bitrev(i, lh);
            if (i < j) {   /* each unordered pair is swapped by exactly one iteration: no races (matrix/src/util.rs:46-56) */
                memcpy(tmp, mat + i * w, w * 4); memcpy(mat + i * w, mat + j * w, w * 4); memcpy(mat + j * w, tmp, w * 4);
            }
        }
        free(tmp);
    }
}

/* ------------------------------------------------------------------------------------------------
 * DFT family.  Semantics: dft/src/naive.rs:9-32 (definition), dft/src/traits.rs:62-259
 * (coset / inverse / lde definitions), memory layout of coset_lde_batch:
 * dft/src/radix_2_dit_parallel.rs:181-246 + fri/src/two_adic_pcs.rs:312-318.
 * Any correct DFT gives identical (canonical) results, so the algorithm here is a textbook
 * in-place radix-2 DIT on rows, parallelised with OpenMP (the reference uses rayon).
 * ---------------------------------------------------------------------------------------------- */

/* O(h^2) definition: y_i = sum_j c_j w^(ij)   (dft/src/naive.rs:9-32) */
void p3o_naive_dft(int fi, const u32 *in, size_t h, size_t w, u32 *out) {
    const field_t *f = F(fi);
    unsigned lh = log2_strict(h);
    u32 g = f_two_adic_generator(f, lh);
    u32 *pw = (u32 *)malloc(h * 4);
    pw[0] = f_one(f);
    for (size_t i = 1; i < h; i++) pw[i] = f_mul(f, pw[i - 1], g);
    for (size_t i = 0; i < h; i++)
        for (size_t c = 0; c < w; c++) {
            u32 acc = 0;
            for (size_t j = 0; j < h; j++) acc = f_add(f, acc, f_mul(f, in[j * w + c], pw[(i * j) & (h - 1)]));
            out[i * w + c] = acc;
        }
    free(pw);
}

static inline void bf_rows(const field_t *f, u32 *a, u32 *b, size_t w, u32 t, int trivial) {
    if (trivial) {
        for (size_t c = 0; c < w; c++) { u32 x = a[c], y = b[c]; a[c] = f_add(f, x, y); b[c] = f_sub(f, x, y); }
    } else {
        for (size_t c = 0; c < w; c++) { u32 x = a[c], y = f_mul(f, b[c], t); a[c] = f_add(f, x, y); b[c] = f_sub(f, x, y); }
    }
}
/* in-place forward DFT of every column with the given primitive h-th root; natural in, natural out.
 * Textbook radix-2 DIT after a row bit-reversal, run as two cache-blocked halves like the reference's
 * Radix2DitParallel (dft/src/radix_2_dit_parallel.rs:22-28): layers 0..mid-1 inside blocks of 2^mid consecutive rows,
 * layers mid..log_h-1 inside the 2^mid interleaved row sets {q*2^mid + r}. Each half is one sweep over the matrix. */
static void dft_rows(const field_t *f, u32 *mat, size_t h, size_t w, u32 root) {
    if (h <= 1) return;
    unsigned lh = log2_strict(h);
    p3o_reverse_matrix_index_bits(mat, h, w);
    u32 *tw = (u32 *)malloc((h / 2) * 4);
    f_powers(f, tw, h / 2, root);
    unsigned mid = (lh + 1) / 2;
    size_t B = (size_t)1 << mid, nblk = h >> mid;
    #pragma omp parallel for schedule(static)
    for (size_t blk = 0; blk < nblk; blk++) {
        u32 *base = mat + blk * B * w;
        for (unsigned layer = 0; layer < mid; layer++) {
            size_t half = (size_t)1 << layer, step = h >> (layer + 1);
            for (size_t s = 0; s < B; s += 2 * half)
                for (size_t j = 0; j < half; j++)
                    bf_rows(f, base + (s + j) * w, base + (s + j + half) * w, w, tw[j * step], j == 0);
        }
    }
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < B; r++) {
        for (unsigned layer = mid; layer < lh; layer++) {
            unsigned t = layer - mid;
            size_t halfq = (size_t)1 << t, step = h >> (layer + 1);
            for (size_t qs = 0; qs < nblk; qs += 2 * halfq)
                for (size_t qj = 0; qj < halfq; qj++) {
                    size_t i = (qs + qj) * B + r, j = qj * B + r;
                    bf_rows(f, mat + i * w, mat + (i + (halfq << mid)) * w, w, tw[j * step], j == 0);
                }
        }
    }
    free(tw);
}

/* dft/src/traits.rs:62 — dft_batch, natural order output */
void p3o_dft_batch(int fi, u32 *mat, size_t h, size_t w) {
    const field_t *f = F(fi);
    dft_rows(f, mat, h, w, f_two_adic_generator(f, log2_strict(h)));
}
/* dft/src/util.rs:32-55 — coset_shift_cols: row i *= shift^i */
static void coset_shift_rows(const field_t *f, u32 *mat, size_t h, size_t w, u32 shift) {
    u32 *pw = (u32 *)malloc(h * 4);
    f_powers(f, pw, h, shift);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < h; i++) for (size_t c = 0; c < w; c++) mat[i * w + c] = f_mul(f, mat[i * w + c], pw[i]);
    free(pw);
}
/* dft/src/traits.rs:84-92 */
void p3o_coset_dft_batch(int fi, u32 *mat, size_t h, size_t w, u32 shift) {
    coset_shift_rows(F(fi), mat, h, w, shift);
    p3o_dft_batch(fi, mat, h, w);
}
/* dft/src/traits.rs:112-123: dft, divide by h, reverse rows 1..h-1  (== dft with the inverse root, scaled) */
void p3o_idft_batch(int fi, u32 *mat, size_t h, size_t w) {
    const field_t *f = F(fi);
    unsigned lh = log2_strict(h);
    dft_rows(f, mat, h, w, f_inv(f, f_two_adic_generator(f, lh)));
    u32 hinv = f_inv(f, f_to_monty(f, (u32)(h % f->p)));
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < h * w; i++) mat[i] = f_mul(f, mat[i], hinv);
}
/* dft/src/traits.rs:145-155 */
void p3o_coset_idft_batch(int fi, u32 *mat, size_t h, size_t w, u32 shift) {
    p3o_idft_batch(fi, mat, h, w);
    coset_shift_rows(F(fi), mat, h, w, f_inv(F(fi), shift));
}
/* coset_lde_batch: dft/src/traits.rs:227-259 (definition), radix_2_dit_parallel.rs:181-246 (layout).
 * in: h x w evaluations on H (natural order).  out: (h<<added_bits) x w evaluations on shift*K.
 * bitrev_out != 0: memory row m holds the evaluation at shift*w_K^bitrev(m)  — the layout that
 * Radix2DitParallel leaves in memory and TwoAdicFriPcs::commit commits (two_adic_pcs.rs:313-318).
 * bitrev_out == 0: natural order (what .to_row_major_matrix() of the returned view yields). */
void p3o_coset_lde_batch(int fi, const u32 *in, size_t h, size_t w, unsigned added_bits, u32 shift, u32 *out, int bitrev_out) {
    const field_t *f = F(fi);
    unsigned lh = log2_strict(h);
    size_t nc = (size_t)1 << added_bits;
    u32 *coeffs = (u32 *)malloc(h * w * 4);
    par_copy(coeffs, in, h * w);
    p3o_idft_batch(fi, coeffs, h, w);
    u32 g_big = f_two_adic_generator(f, lh + added_bits);
    u32 *tmp = (u32 *)malloc(h * w * 4);
    for (size_t c = 0; c < nc; c++) {
        /* coset c (natural coset index): points shift * g_big^c * H */
        u32 s = f_mul(f, shift, f_pow(f, g_big, c));
        par_copy(tmp, coeffs, h * w);
        p3o_coset_dft_batch(fi, tmp, h, w, s);
        /* natural LDE index of (coset c, j) is j*nc + c */
        #pragma omp parallel for schedule(static)
        for (size_t j = 0; j < h; j++) {
            size_t nat = j * nc + c;
            size_t row = bitrev_out ? bitrev(nat, lh + added_bits) : nat;
            memcpy(out + row * w, tmp + j * w, w * 4);
        }
    }
    free(tmp); free(coeffs);
}

/* ------------------------------------------------------------------------------------------------
 * Poseidon2: poseidon2/src/lib.rs:131-147, poseidon2/src/external.rs:60-74,113-159,288-336,
 * monty-31/src/poseidon2.rs:76-85, diagonals baby-bear/src/poseidon2.rs:394-450,
 * koala-bear/src/poseidon2.rs:407-461.  Round constants are runtime inputs (Poseidon2::new).
 * ---------------------------------------------------------------------------------------------- */
#define P2_MAXW 24
typedef struct {
    int field;                 /* 0 BabyBear, 1 KoalaBear */
    int width;                 /* 16 or 24 */
    int rounds_p;              /* number of internal rounds */
    u32 rc_init[4 * P2_MAXW];  /* 4 initial external rounds x width, Montgomery form */
    u32 rc_term[4 * P2_MAXW];  /* 4 terminal external rounds x width */
    u32 rc_int[32];            /* internal round constants */
} p3o_perm;

/* internal diagonal as exponents/signs.  kind: 0 => +2^k (k may be negative => division), special
 * entries for 1,2,3,4 etc. are expressed through small integer multipliers. */
typedef struct { int mul; int shift; } diag_t; /* value = mul * 2^shift, mul in {-4..4}, shift <= 0 */

static const diag_t DIAG_BB16[16] = {{-2,0},{1,0},{2,0},{1,-1},{3,0},{4,0},{-1,-1},{-3,0},{-4,0},{1,-8},{1,-2},{1,-3},{1,-27},{-1,-8},{-1,-4},{-1,-27}};
static const diag_t DIAG_BB24[24] = {{-2,0},{1,0},{2,0},{1,-1},{3,0},{4,0},{-1,-1},{-3,0},{-4,0},{1,-8},{1,-2},{1,-3},{1,-4},{1,-7},{1,-9},{1,-27},{-1,-8},{-1,-2},{-1,-3},{-1,-4},{-1,-5},{-1,-6},{-1,-7},{-1,-27}};
static const diag_t DIAG_KB16[16] = {{-2,0},{1,0},{2,0},{1,-1},{3,0},{4,0},{-1,-1},{-3,0},{-4,0},{1,-8},{1,-3},{1,-24},{-1,-8},{-1,-3},{-1,-4},{-1,-24}};
static const diag_t DIAG_KB24[24] = {{-2,0},{1,0},{2,0},{1,-1},{3,0},{4,0},{-1,-1},{-3,0},{-4,0},{1,-8},{1,-2},{1,-3},{1,-4},{1,-5},{1,-6},{1,-24},{-1,-8},{-1,-3},{-1,-4},{-1,-5},{-1,-6},{-1,-7},{-1,-9},{-1,-24}};

static const diag_t *p2_diag(int field, int width) {
    if (field == 0) return width == 16 ? DIAG_BB16 : DIAG_BB24;
    return width == 16 ? DIAG_KB16 : DIAG_KB24;
}
static u32 diag_value(const field_t *f, diag_t d) {
    u32 m = d.mul >= 0 ? f_to_monty(f, (u32)d.mul) : f_sub(f, 0, f_to_monty(f, (u32)(-d.mul)));
    u32 half = f_inv(f, f_to_monty(f, 2));
    for (int i = 0; i < -d.shift; i++) m = f_mul(f, m, half);
    return m;
}
/* Montgomery-form internal diagonal V (exported for the GPU side and for tests) */
void p3o_poseidon2_diag(int field, int width, u32 *out) {
    const diag_t *d = p2_diag(field, width);
    for (int i = 0; i < width; i++) out[i] = diag_value(F(field), d[i]);
}

static inline u32 sbox(const field_t *f, u32 x) {
    u32 x2 = f_mul(f, x, x), x3 = f_mul(f, x2, x);
    if (f->sbox_d == 3) return x3;
    u32 x4 = f_mul(f, x2, x2);
    return f_mul(f, x4, x3); /* x^7 */
}
/* poseidon2/src/external.rs:60-74: circ(2,3,1,1) */
static inline void mat4(const field_t *f, u32 *x) {
    u32 t01 = f_add(f, x[0], x[1]), t23 = f_add(f, x[2], x[3]);
    u32 t0123 = f_add(f, t01, t23);
    u32 t01123 = f_add(f, t0123, x[1]), t01233 = f_add(f, t0123, x[3]);
    u32 n3 = f_add(f, t01233, f_add(f, x[0], x[0]));
    u32 n1 = f_add(f, t01123, f_add(f, x[2], x[2]));
    u32 n0 = f_add(f, t01123, t01);
    u32 n2 = f_add(f, t01233, t23);
    x[0] = n0; x[1] = n1; x[2] = n2; x[3] = n3;
}
/* poseidon2/src/external.rs:113-159 */
static void mds_light(const field_t *f, u32 *s, int w) {
    for (int i = 0; i < w; i += 4) mat4(f, s + i);
    u32 sums[4] = {0, 0, 0, 0};
    for (int i = 0; i < w; i++) sums[i & 3] = f_add(f, sums[i & 3], s[i]);
    for (int i = 0; i < w; i++) s[i] = f_add(f, s[i], sums[i & 3]);
}
void p3o_poseidon2_permute(const p3o_perm *pm, u32 *s) {
    const field_t *f = F(pm->field);
    int w = pm->width;
    u32 diag[P2_MAXW];
    p3o_poseidon2_diag(pm->field, w, diag);
    /* external.rs:316-336 initial: MDS-light, then 4 x {+rc, sbox, MDS-light} */
    mds_light(f, s, w);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < w; i++) s[i] = sbox(f, f_add(f, s[i], pm->rc_init[r * w + i]));
        mds_light(f, s, w);
    }
    /* monty-31/src/poseidon2.rs:76-85 internal: s0=(s0+rc)^d ; s_i = V_i*s_i + sum */
    for (int r = 0; r < pm->rounds_p; r++) {
        s[0] = sbox(f, f_add(f, s[0], pm->rc_int[r]));
        u32 sum = 0;
        for (int i = 0; i < w; i++) sum = f_add(f, sum, s[i]);
        for (int i = 0; i < w; i++) s[i] = f_add(f, f_mul(f, s[i], diag[i]), sum);
    }
    /* external.rs:288-310 terminal */
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < w; i++) s[i] = sbox(f, f_add(f, s[i], pm->rc_term[r * w + i]));
        mds_light(f, s, w);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Keccak-f[1600] (FIPS-202).  Reference: keccak/src/lib.rs:70-76 -> tiny_keccak::keccakf (third
 * party, tiny-keccak 2.0.2, not vendored); in-repo vectorised statement keccak/src/avx512.rs:12-365.
 * ---------------------------------------------------------------------------------------------- */
static const u64 KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static inline u64 rotl64(u64 x, int r) { return r ? (x << r) | (x >> (64 - r)) : x; }
void p3o_keccak_f(u64 *a) {
    for (int round = 0; round < 24; round++) {
        u64 c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(a[x + 5 * y], KECCAK_ROT[x + 5 * y]);
        for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= KECCAK_RC[round];
    }
}

/* ------------------------------------------------------------------------------------------------
 * Hashers.  Digest is always 8 x u32 (Poseidon2: [F;8]; Keccak: [u64;4] little-endian words).
 *   kind 0: leaf  = PaddingFreeSponge<leaf perm, WIDTH, RATE, 8>   (symmetric/src/sponge.rs:182-216)
 *           node  = TruncatedPermutation<comp perm, 2, 8, 16>       (symmetric/src/compression.rs:34-49)
 *   kind 1: leaf  = SerializingHasher<PaddingFreeSponge<KeccakF,25,17,4>> (serializing_hasher.rs:89-101,
 *                   field/src/integers.rs:494-509)
 *           node  = CompressionFunctionFromHasher<sponge,2,4>       (symmetric/src/compression.rs:60-70)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int kind;
    int leaf_rate;     /* Poseidon2 sponge rate (8 for width 16, 16 for width 24) */
    p3o_perm leaf;     /* leaf sponge permutation */
    p3o_perm comp;     /* width-16 compression permutation */
} p3o_hasher;

/* streaming sponge state so that several matrices' rows can be absorbed back to back */
typedef struct { u32 st[P2_MAXW]; int pos; } p2_sponge;
static void p2_absorb(const p3o_hasher *hs, p2_sponge *sp, const u32 *in, size_t n) {
    for (size_t i = 0; i < n; i++) {
        sp->st[sp->pos++] = in[i];
        if (sp->pos == hs->leaf_rate) { p3o_poseidon2_permute(&hs->leaf, sp->st); sp->pos = 0; }
    }
}
static void p2_finish(const p3o_hasher *hs, p2_sponge *sp, u32 *digest) {
    if (sp->pos != 0) p3o_poseidon2_permute(&hs->leaf, sp->st);
    memcpy(digest, sp->st, 32);
}
typedef struct { u64 st[25]; int pos; int have_lo; u32 lo; } k_sponge;
static void k_push_word(k_sponge *sp, u64 wd) {
    sp->st[sp->pos++] = wd;
    if (sp->pos == 17) { p3o_keccak_f(sp->st); sp->pos = 0; }
}
static void k_absorb(k_sponge *sp, const u32 *in, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (!sp->have_lo) { sp->lo = in[i]; sp->have_lo = 1; }
        else { k_push_word(sp, (u64)sp->lo | ((u64)in[i] << 32)); sp->have_lo = 0; }
    }
}
static void k_finish(k_sponge *sp, u32 *digest) {
    if (sp->have_lo) { k_push_word(sp, (u64)sp->lo); sp->have_lo = 0; }
    if (sp->pos != 0) p3o_keccak_f(sp->st);
    memcpy(digest, sp->st, 32);
}

/* hash the concatenation of n_parts slices (rows of several matrices, input order) */
void p3o_hash_slices(const p3o_hasher *hs, const u32 *const *parts, const size_t *lens, size_t n_parts, u32 *digest) {
    if (hs->kind == 0) {
        p2_sponge sp; memset(&sp, 0, sizeof sp);
        for (size_t k = 0; k < n_parts; k++) p2_absorb(hs, &sp, parts[k], lens[k]);
        p2_finish(hs, &sp, digest);
    } else {
        k_sponge sp; memset(&sp, 0, sizeof sp);
        for (size_t k = 0; k < n_parts; k++) k_absorb(&sp, parts[k], lens[k]);
        k_finish(&sp, digest);
    }
}
void p3o_hash_row(const p3o_hasher *hs, const u32 *row, size_t n, u32 *digest) {
    p3o_hash_slices(hs, &row, &n, 1, digest);
}
void p3o_compress(const p3o_hasher *hs, const u32 *left, const u32 *right, u32 *out) {
    if (hs->kind == 0) {
        u32 st[16];
        memcpy(st, left, 32); memcpy(st + 8, right, 32);
        p3o_poseidon2_permute(&hs->comp, st);
        memcpy(out, st, 32);
    } else {
        u64 st[25]; memset(st, 0, sizeof st);
        memcpy(st, left, 32); memcpy(st + 4, right, 32);  /* 8 words < rate 17: one permutation */
        p3o_keccak_f(st);
        memcpy(out, st, 32);
    }
}

/* ------------------------------------------------------------------------------------------------
 * MerkleTree::new for arity N=2 (merkle-tree/src/merkle_tree.rs:95-178, 268-338, 348-460, 473-538),
 * mixed heights included.  Output: all digest layers concatenated (layer 0 first); layer_lens[k]
 * receives the (padded) length of layer k; returns the number of layers.
 * ---------------------------------------------------------------------------------------------- */
static size_t next_pow2(size_t x) { size_t p = 1; while (p < x) p <<= 1; return p; }
/* merkle_tree.rs:473-481 with n = 2 */
static size_t padded_len2(size_t raw) { return raw <= 1 ? raw : (raw + 1) / 2 * 2; }

/* mmcs/geometry.rs:83-124; returns 0 if ok */
int p3o_validate_heights(const size_t *hs_, size_t n) {
    size_t maxh = 0;
    for (size_t i = 0; i < n; i++) if (hs_[i] > maxh) maxh = hs_[i];
    if (maxh == 0) return -1;
    unsigned lmax = 0; while (((size_t)1 << lmax) < maxh) lmax++;
    for (size_t i = 0; i < n; i++) {
        if (hs_[i] == 0) return -2;
        unsigned l = 0; while (((size_t)1 << l) < hs_[i]) l++;
        size_t expect = ((maxh - 1) >> (lmax - l)) + 1;
        if (hs_[i] != expect) return -2;
    }
    return 0;
}

size_t p3o_merkle_tree(const p3o_hasher *hs, size_t n_mats, const u32 *const *mats, const size_t *heights,
                       const size_t *widths, u32 *layers_out, size_t *layer_lens) {
    if (p3o_validate_heights(heights, n_mats) != 0) return 0;
    /* stable sort indices by height, tallest first (merkle_tree.rs:124-127) */
    size_t *order = (size_t *)malloc(n_mats * sizeof(size_t));
    for (size_t i = 0; i < n_mats; i++) order[i] = i;
    for (size_t i = 1; i < n_mats; i++) { /* insertion sort keeps input order within a height class */
        size_t k = order[i], j = i;
        while (j > 0 && heights[order[j - 1]] < heights[k]) { order[j] = order[j - 1]; j--; }
        order[j] = k;
    }
    size_t max_h = heights[order[0]];
    size_t next = 0; /* cursor into order[] */
    size_t grp_end = 0;
    while (grp_end < n_mats && heights[order[grp_end]] == max_h) grp_end++;

    /* first_digest_layer (merkle_tree.rs:268-338) */
    size_t len0 = padded_len2(max_h);
    u32 *cur = layers_out;
    memset(cur, 0, len0 * 32);
    {
        size_t gn = grp_end;
        #pragma omp parallel for schedule(static)
        for (size_t r = 0; r < max_h; r++) {
            const u32 *parts[64]; size_t lens[64];
            for (size_t k = 0; k < gn; k++) { parts[k] = mats[order[k]] + r * widths[order[k]]; lens[k] = widths[order[k]]; }
            p3o_hash_slices(hs, parts, lens, gn, cur + r * 8);
        }
    }
    next = grp_end;
    size_t n_layers = 1;
    layer_lens[0] = len0;
    size_t prev_len = len0;
    u32 *prev = cur;
    const u32 zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (prev_len > 1) {
        size_t raw_next = prev_len / 2;
        size_t next_layer_len = next_pow2(raw_next);
        /* matrices injected at this level: padded height == next_layer_len (merkle_tree.rs:150-152) */
        size_t inj_begin = next, inj_end = next;
        while (inj_end < n_mats && next_pow2(heights[order[inj_end]]) == next_layer_len) inj_end++;
        next = inj_end;
        size_t out_len = padded_len2(raw_next);
        u32 *out = prev + prev_len * 8;
        memset(out, 0, out_len * 32);
        size_t inj_h = inj_end > inj_begin ? heights[order[inj_begin]] : 0;
        size_t gi = inj_begin, gn = inj_end - inj_begin;
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < raw_next; i++) {
            u32 d[8];
            p3o_compress(hs, prev + (2 * i) * 8, prev + (2 * i + 1) * 8, d);
            if (gn > 0) { /* compress_and_inject (merkle_tree.rs:348-460) */
                if (i < inj_h) {
                    const u32 *parts[64]; size_t lens[64]; u32 rd[8];
                    for (size_t k = 0; k < gn; k++) { parts[k] = mats[order[gi + k]] + i * widths[order[gi + k]]; lens[k] = widths[order[gi + k]]; }
                    p3o_hash_slices(hs, parts, lens, gn, rd);
                    p3o_compress(hs, d, rd, out + i * 8);
                } else {
                    p3o_compress(hs, d, zero, out + i * 8);
                }
            } else {
                memcpy(out + i * 8, d, 32);
            }
        }
        layer_lens[n_layers++] = out_len;
        prev = out; prev_len = out_len;
    }
    free(order);
    return n_layers;
}

/* upper bound on total digests written by p3o_merkle_tree for a given max height */
size_t p3o_merkle_total_digests(size_t max_h) {
    size_t tot = 0, len = padded_len2(max_h);
    tot += len;
    while (len > 1) { len = padded_len2(len / 2); tot += len; }
    return tot;
}

/* ------------------------------------------------------------------------------------------------
 * EF4 = F[X]/(X^4 - W): field/src/extension/binomial_extension.rs:724-770 (quartic_mul)
 * ---------------------------------------------------------------------------------------------- */
static void ef_mul(const field_t *f, const u32 *a, const u32 *b, u32 *out) {
    u32 wm = f_to_monty(f, f->ext_w);
    u32 r[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i + j] = f_add(f, r[i + j], f_mul(f, a[i], b[j]));
    u32 o[4];
    for (int i = 0; i < 3; i++) o[i] = f_add(f, r[i], f_mul(f, wm, r[i + 4]));
    o[3] = r[3];
    memcpy(out, o, 16);
}
void p3o_ef_mul(int fi, const u32 *a, const u32 *b, u32 *out) { ef_mul(F(fi), a, b, out); }

/* ------------------------------------------------------------------------------------------------
 * TwoAdicFriFolding::fold_matrix: fri/src/two_adic_pcs.rs:134-213.
 * in: rows x arity EF4 elements (rows*arity*4 u32), bit-reversed evaluation order; out: rows EF4.
 * ---------------------------------------------------------------------------------------------- */
void p3o_fold_matrix(int fi, const u32 *in, size_t rows, unsigned log_arity, const u32 *beta, u32 *out) {
    const field_t *f = F(fi);
    size_t len = rows << log_arity;
    u32 *data = (u32 *)malloc(len * 16), *nxt = (u32 *)malloc(len * 8);
    memcpy(data, in, len * 16);
    u32 cur_beta[4]; memcpy(cur_beta, beta, 16);
    u32 half = f_halve(f, f_one(f));
    for (unsigned step = 0; step < log_arity; step++) {
        size_t height = len >> 1;
        unsigned lh = log2_strict(height);
        u32 g_inv = f_inv(f, f_two_adic_generator(f, lh + 1));
        /* halve_inv_powers[j] = (1/2) * g_inv^bitrev(j)   (two_adic_pcs.rs:151-155,174-192) */
        u32 *pw = (u32 *)malloc(height * 4);
        pw[0] = half;
        for (size_t i = 1; i < height; i++) pw[i] = f_mul(f, pw[i - 1], g_inv);
        #pragma omp parallel for schedule(static)
        for (size_t j = 0; j < height; j++) {
            const u32 *lo = data + (2 * j) * 4, *hi = lo + 4;
            u32 t = pw[bitrev(j, lh)];
            u32 sum[4], dif[4], db[4];
            for (int k = 0; k < 4; k++) { sum[k] = f_halve(f, f_add(f, lo[k], hi[k])); dif[k] = f_sub(f, lo[k], hi[k]); }
            ef_mul(f, dif, cur_beta, db);
            for (int k = 0; k < 4; k++) nxt[j * 4 + k] = f_add(f, sum[k], f_mul(f, db[k], t));
        }
        free(pw);
        memcpy(data, nxt, height * 16);
        len = height;
        ef_mul(f, cur_beta, cur_beta, cur_beta);
    }
    memcpy(out, data, rows * 16);
    free(data); free(nxt);
}

/* ------------------------------------------------------------------------------------------------
 * TwoAdicFriPcs::open — the pre-FRI device work:
 *   compute_inverse_denominators   fri/src/two_adic_pcs.rs:743-780
 *   columnwise_dot_product / interpolate_coset_with_precomputation   matrix/src/interpolation.rs:161-193
 *   rowwise dot with powers of alpha + quotient accumulation          fri/src/two_adic_pcs.rs:622-657
 * EF4 inverse through the Frobenius conjugates (any method gives the same canonical result).
 * ---------------------------------------------------------------------------------------------- */
static void ef_inv(const field_t *f, const u32 *a, u32 *out) {
    /* zeta = W^((p-1)/4): phi(X) = zeta * X, zeta^2 = -1 */
    u32 zeta = f_pow(f, f_to_monty(f, f->ext_w), ((u64)f->p - 1) / 4);
    u32 a1z = f_mul(f, a[1], zeta), a3z = f_mul(f, a[3], zeta);
    u32 c1[4] = {a[0], a1z, f_sub(f, 0, a[2]), f_sub(f, 0, a3z)};
    u32 c2[4] = {a[0], f_sub(f, 0, a[1]), a[2], f_sub(f, 0, a[3])};
    u32 c3[4] = {a[0], f_sub(f, 0, a1z), f_sub(f, 0, a[2]), a3z};
    u32 b[4], n[4];
    ef_mul(f, c1, c2, b); ef_mul(f, b, c3, b);
    ef_mul(f, a, b, n);            /* norm: lies in the base field (n[1..3] == 0) */
    u32 ninv = f_inv(f, n[0]);
    for (int k = 0; k < 4; k++) out[k] = f_mul(f, b[k], ninv);
}
void p3o_ef_inv(int fi, const u32 *a, u32 *out) { ef_inv(F(fi), a, out); }

/* out[i] = 1 / (z - x_i), x_i = GENERATOR * w^bitrev(i), i < 2^log_h   (two_adic_pcs.rs:486-494, 743-780) */
void p3o_open_inv_denoms(int fi, unsigned log_h, const u32 *z, u32 *out) {
    const field_t *f = F(fi);
    size_t h = (size_t)1 << log_h;
    u32 g = f_to_monty(f, f->gen), w = f_two_adic_generator(f, log_h);
    u32 *pts = (u32 *)malloc(h * 4);
    pts[0] = g;
    for (size_t i = 1; i < h; i++) pts[i] = f_mul(f, pts[i - 1], w);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < h; i++) {
        u32 d[4] = {f_sub(f, z[0], pts[bitrev(i, log_h)]), z[1], z[2], z[3]};
        ef_inv(f, d, out + 4 * i);
    }
    free(pts);
}
/* out[j] = sum_i mat[i][j] * v[i]   (v: h EF4 values; matrix/src/lib.rs columnwise_dot_product) */
void p3o_columnwise_dot(int fi, const u32 *mat, size_t h, size_t w, const u32 *v, u32 *out) {
    const field_t *f = F(fi);
    #pragma omp parallel for schedule(static)
    for (size_t j = 0; j < w; j++) {
        u32 acc[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < h; i++)
            for (int k = 0; k < 4; k++) acc[k] = f_add(f, acc[k], f_mul(f, mat[i * w + j], v[4 * i + k]));
        memcpy(out + 4 * j, acc, 16);
    }
}
/* out[i] = sum_j alpha^j * mat[i][j]   (rowwise_packed_dot_product with packed_ext_powers, two_adic_pcs.rs:622-626) */
void p3o_rowwise_dot(int fi, const u32 *mat, size_t h, size_t w, const u32 *alpha, u32 *out) {
    const field_t *f = F(fi);
    u32 *pw = (u32 *)malloc(w * 16);
    u32 cur[4] = {f_one(f), 0, 0, 0};
    for (size_t j = 0; j < w; j++) { memcpy(pw + 4 * j, cur, 16); ef_mul(f, cur, alpha, cur); }
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < h; i++) {
        u32 acc[4] = {0, 0, 0, 0};
        for (size_t j = 0; j < w; j++)
            for (int k = 0; k < 4; k++) acc[k] = f_add(f, acc[k], f_mul(f, mat[i * w + j], pw[4 * j + k]));
        memcpy(out + 4 * i, acc, 16);
    }
    free(pw);
}
/* ro[i] += coeff * (yred - r[i]) * inv_denom[i]   (two_adic_pcs.rs:640-657) */
void p3o_open_reduce(int fi, u32 *ro, const u32 *r, const u32 *inv_denoms, size_t h, const u32 *coeff, const u32 *yred) {
    const field_t *f = F(fi);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < h; i++) {
        u32 d[4], t[4];
        for (int k = 0; k < 4; k++) d[k] = f_sub(f, yred[k], r[4 * i + k]);
        ef_mul(f, coeff, d, t);
        ef_mul(f, t, inv_denoms + 4 * i, t);
        for (int k = 0; k < 4; k++) ro[4 * i + k] = f_add(f, ro[4 * i + k], t[k]);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Poseidon2 AIR: VectorizedPoseidon2Air with WIDTH 16, S-box degree 3 and 0 S-box registers
 * (the KoalaBear instance of prove_prime_field_31, examples/examples/prove_prime_field_31.rs:150-165).
 *   columns of one permutation (poseidon2-air/src/columns.rs:11-48):
 *     inputs[16] | 4 x post[16] (beginning full rounds) | rounds_p x post_sbox | 4 x post[16] (ending full rounds)
 *   trace generation: poseidon2-air/src/generation.rs:184-253 (+ full/partial round :430-553)
 *   constraints:      poseidon2-air/src/air.rs:173-240 (eval), :254-275 (full round), :277-296 (partial round),
 *                     vectorised: poseidon2-air/src/vectorized.rs:297-311 (VECTOR_LEN permutations per row, in order)
 *   folding:          uni-stark/src/folder.rs (first asserted constraint gets the highest power of alpha),
 *                     quotient = folded * inv_vanishing, uni-stark/src/prover.rs:462-827, commit/src/domain.rs:321-361
 * The linear layers are the permutation's own (GenericPoseidon2LinearLayersMonty31 == the optimised layers,
 * koala-bear/src/poseidon2.rs:575-610).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int field;          /* 1 = KoalaBear (degree-3 S-box, no registers); other instances are not restated */
    int rounds_p;       /* partial rounds (20 for KoalaBear width 16) */
    u32 beg[4 * 16];    /* beginning_full_round_constants, Montgomery */
    u32 part[32];       /* partial_round_constants */
    u32 end[4 * 16];    /* ending_full_round_constants */
} p3o_air;

size_t p3o_p2air_cols(const p3o_air *a) { return 16 + 64 + (size_t)a->rounds_p + 64; }
size_t p3o_p2air_constraints(const p3o_air *a) { return 64 + (size_t)a->rounds_p + 64; }

static void air_internal_layer(const field_t *f, const u32 *diag, u32 *s) {
    u32 sum = 0;
    for (int i = 0; i < 16; i++) sum = f_add(f, sum, s[i]);
    for (int i = 0; i < 16; i++) s[i] = f_add(f, f_mul(f, s[i], diag[i]), sum);
}

/* generate_trace_rows_for_perm for n_perms inputs; row-major n_perms x cols (the vectorised trace is the same buffer viewed as
 * (n_perms / VECTOR_LEN) x (VECTOR_LEN * cols), generation.rs:14-70) */
void p3o_p2air_generate(const p3o_air *a, const u32 *inputs, size_t n_perms, u32 *trace) {
    const field_t *f = F(a->field);
    const size_t cols = p3o_p2air_cols(a);
    u32 diag[P2_MAXW];
    p3o_poseidon2_diag(a->field, 16, diag);
    #pragma omp parallel for schedule(static)
    for (size_t p = 0; p < n_perms; p++) {
        u32 s[16];
        u32 *row = trace + p * cols;
        memcpy(s, inputs + p * 16, 64);
        memcpy(row, s, 64);
        row += 16;
        mds_light(f, s, 16);
        for (int r = 0; r < 4; r++) {
            for (int i = 0; i < 16; i++) s[i] = sbox(f, f_add(f, s[i], a->beg[r * 16 + i]));
            mds_light(f, s, 16);
            memcpy(row, s, 64); row += 16;
        }
        for (int r = 0; r < a->rounds_p; r++) {
            s[0] = sbox(f, f_add(f, s[0], a->part[r]));
            *row++ = s[0];
            air_internal_layer(f, diag, s);
        }
        for (int r = 0; r < 4; r++) {
            for (int i = 0; i < 16; i++) s[i] = sbox(f, f_add(f, s[i], a->end[r * 16 + i]));
            mds_light(f, s, 16);
            memcpy(row, s, 64); row += 16;
        }
    }
}

/* the constraints of ONE permutation on one row of column values (any field point of the LDE): out[k], k < constraints */
static void air_eval_perm(const p3o_air *a, const field_t *f, const u32 *diag, const u32 *c, u32 *out) {
    u32 s[16];
    memcpy(s, c, 64); c += 16;
    mds_light(f, s, 16);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox(f, f_add(f, s[i], a->beg[r * 16 + i]));
        mds_light(f, s, 16);
        for (int i = 0; i < 16; i++) { *out++ = f_sub(f, s[i], c[i]); s[i] = c[i]; }     /* assert_eq(state_i, post_i); state_i = post_i */
        c += 16;
    }
    for (int r = 0; r < a->rounds_p; r++) {
        const u32 x = sbox(f, f_add(f, s[0], a->part[r]));
        *out++ = f_sub(f, x, *c);                                                           /* assert_eq(state_0, post_sbox) */
        s[0] = *c++;
        air_internal_layer(f, diag, s);
    }
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox(f, f_add(f, s[i], a->end[r * 16 + i]));
        mds_light(f, s, 16);
        for (int i = 0; i < 16; i++) { *out++ = f_sub(f, s[i], c[i]); s[i] = c[i]; }
        c += 16;
    }
}

/* check_constraints-style helper for tests: number of non-zero constraints over a (vectorised) trace */
size_t p3o_p2air_check(const p3o_air *a, int vec_len, const u32 *trace, size_t rows) {
    const field_t *f = F(a->field);
    const size_t cols = p3o_p2air_cols(a), nc = p3o_p2air_constraints(a);
    u32 diag[P2_MAXW];
    p3o_poseidon2_diag(a->field, 16, diag);
    size_t bad = 0;
    for (size_t r = 0; r < rows; r++)
        for (int v = 0; v < vec_len; v++) {
            u32 out[64 + 32 + 64];
            air_eval_perm(a, f, diag, trace + (r * vec_len + v) * cols, out);
            for (size_t k = 0; k < nc; k++) bad += out[k] != 0;
        }
    return bad;
}

/* quotient_values (uni-stark/src/prover.rs:462-827) for this AIR.  lde: the committed trace LDE, H = 2^log_h rows in
 * BIT-REVERSED order (row m = evaluation at GENERATOR * w_H^bitrev(m)), width vec_len * cols; log_n: log2 of the trace height.
 * q: H x 4, NATURAL order over the quotient domain GENERATOR * K, |K| = H (requires quotient degree == blow-up, the fast path of
 * get_evaluations_on_domain, two_adic_pcs.rs:376-385).  alpha: 4 Montgomery words. */
void p3o_p2air_quotient(const p3o_air *a, int vec_len, const u32 *lde, unsigned log_h, unsigned log_n, const u32 *alpha, u32 *q) {
    const field_t *f = F(a->field);
    const size_t cols = p3o_p2air_cols(a), nc = p3o_p2air_constraints(a), n_all = nc * (size_t)vec_len, H = (size_t)1 << log_h;
    const unsigned rate_bits = log_h - log_n;
    u32 diag[P2_MAXW];
    p3o_poseidon2_diag(a->field, 16, diag);
    /* alpha powers: constraint j is multiplied by alpha^(n_all - 1 - j)  (Horner, folder.rs:374-376) */
    u32 *apow = (u32 *)malloc(n_all * 16);
    apow[0] = f_one(f); apow[1] = apow[2] = apow[3] = 0;
    for (size_t j = 1; j < n_all; j++) ef_mul(f, apow + 4 * (j - 1), alpha, apow + 4 * j);
    /* inv_vanishing on the coset (domain.rs:326-360): Z_H(x_i) = s^N * w_(2^rate_bits)^(i mod 2^rate_bits) - 1 */
    const size_t nz = (size_t)1 << rate_bits;
    u32 *invz = (u32 *)malloc(nz * 4);
    const u32 s_pow_n = f_pow(f, f_to_monty(f, f->gen), (u64)1 << log_n), wr = f_two_adic_generator(f, rate_bits);
    for (size_t k = 0; k < nz; k++) invz[k] = f_inv(f, f_sub(f, f_mul(f, s_pow_n, f_pow(f, wr, k)), f_one(f)));
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < H; i++) {
        const u32 *row = lde + bitrev(i, log_h) * (cols * vec_len);
        u32 acc[4] = {0, 0, 0, 0}, out[64 + 32 + 64];
        for (int v = 0; v < vec_len; v++) {
            air_eval_perm(a, f, diag, row + (size_t)v * cols, out);
            for (size_t k = 0; k < nc; k++) {
                const u32 *ap = apow + 4 * (n_all - 1 - ((size_t)v * nc + k));
                for (int d = 0; d < 4; d++) acc[d] = f_add(f, acc[d], f_mul(f, out[k], ap[d]));
            }
        }
        for (int d = 0; d < 4; d++) q[4 * i + d] = f_mul(f, acc[d], invz[i & (nz - 1)]);
    }
    free(apow); free(invz);
}

/* rand 0.10 SmallRng on 64-bit targets (xoshiro256++ seeded through SplitMix64) + the MontyField31 sampler
 * (monty-31/src/monty_31.rs:154-165: v = next_u32 >> 1, rejected if >= p, value taken AS the Montgomery representation).
 * Pinned by the fixture replay.  Fills out[n] with consecutive field samples; state (4 x u64) in/out. */
void p3o_smallrng_seed(u64 seed, u64 *st) {
    u64 x = seed;
    for (int i = 0; i < 4; i++) {
        x += 0x9e3779b97f4a7c15ULL;
        u64 z = x;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        st[i] = z ^ (z >> 31);
    }
}
void p3o_smallrng_field(int fi, u64 *s, u32 *out, size_t n) {
    const u32 p = F(fi)->p;
    for (size_t k = 0; k < n;) {
        const u64 r = rotl64(s[0] + s[3], 23) + s[0], t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
        const u32 v = (u32)(r >> 32) >> 1;
        if (v < p) out[k++] = v;
    }
}
