// Integer-pipe microbenchmarks for the butterfly / Poseidon2 inner loops on sm_100a (developer tool).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../plonky3_b200/csrc/field.cuh"
using namespace p3;
constexpr int F = KOALA_BEAR;
constexpr u32 P = Fp<F>::P;

#define ITERS 2048

template <int MODE> __global__ void k(u32 *out, u32 seed, uint2 tw0) {
    u32 x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = seed * (i + 1) + threadIdx.x;
    uint2 tw = tw0; tw.x += threadIdx.x & 1;
    const u32 c1 = seed | 1, c2 = seed * 3 + 7;
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {  // IMAD lo
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = x[i] * c1 + c2;
        } else if (MODE == 1) {  // IMAD.HI
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = __umulhi(x[i], c1) + c2;
        } else if (MODE == 2) {  // IADD3
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = x[i] + c1 + x[(i + 1) & 15];
        } else if (MODE == 3) {  // VIADDMNMX
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = min(x[i] ^ 0u, x[i] - P) + 0;
        } else if (MODE == 4) {  // LOP3
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = (x[i] ^ c1) & (x[(i + 1) & 15] | c2);
        } else if (MODE == 5) {  // shoup butterfly, 8 per iteration
#pragma unroll
            for (int i = 0; i < 8; i++) ct_butterfly<F>(x[i], x[i + 8], tw);
        } else if (MODE == 6) {  // montgomery mul
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = mont_mul<F>(x[i], x[(i + 5) & 15]);
        } else if (MODE == 7) {  // add mod (canonical)
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fp_add<F>(x[i], x[(i + 3) & 15]);
        } else if (MODE == 8) {  // SHF funnel (64-bit rotate halves)
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = __funnelshift_l(x[i], x[(i + 1) & 15], 13);
        } else if (MODE == 9) {  // mul.wide + use both halves
#pragma unroll
            for (int i = 0; i < 16; i++) { u64 p = (u64)x[i] * c1; x[i] = (u32)p ^ (u32)(p >> 32); }
        } else if (MODE == 10) {  // butterfly variant: everything after the 3 multiplies on the FMA pipe where possible
#pragma unroll
            for (int i = 0; i < 8; i++) {
                u32 a = x[i], b = x[i + 8];
                u32 u = fp_reduce<F>(a);
                u32 q = __umulhi(b, tw.y);
                u32 r = fp_reduce<F>(b * tw.x - q * P);
                u32 nr = r * 0xffffffffu + P;  // P - r on the multiply pipe
                x[i] = u + r; x[i + 8] = u + nr;
            }
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, double ops_per_iter, u32 *d) {
    for (int threads : {128, 256, 512, 1024}) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<MODE><<<148, threads>>>(d, 12345, make_uint2(1234567, 2489012));
        cudaEventRecord(a);
        k<MODE><<<148, threads>>>(d, 12345, make_uint2(1234567, 2489012));
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double total = ops_per_iter * ITERS * threads * 148.0;
        printf("%-28s threads/SM %4d: %8.3f ms  %7.1f Gops/s  = %6.2f ops/clk/SM @1.92GHz\n", name, threads, ms, total / ms / 1e6, total / (ms * 1e-3) / 148 / 1.92e9);
    }
}
int main() {
    u32 *d; cudaMalloc(&d, 148 * 1024 * 4);
    run<0>("IMAD.lo", 16, d); run<1>("IMAD.HI(+add)", 16, d); run<2>("IADD3", 16, d); run<3>("VIADDMNMX(min x,x-P)", 16, d);
    run<4>("LOP3 (x^c)&(y|c)", 16, d); run<8>("SHF funnel", 16, d); run<9>("mul.wide+xor", 16, d);
    run<7>("fp_add (canonical)", 16, d); run<6>("mont_mul", 16, d);
    run<5>("ct_butterfly (shoup lazy)", 8, d); run<10>("butterfly variant FMA-heavy", 8, d);
    return 0;
}
