// Probe: 5-D tiled TMA load of one NTT tile with a box LARGER than the tensor extent in dim 1 (zero-filled pad row).
// Layout wanted in shared memory: [G][mu (E2 + 1 rows, last one padding)][CT].   nvcc -arch=sm_100a tma_probe.cu -o tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef unsigned u32;
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int CT = 8, E2 = 32, E1 = 32;
constexpr int BOX_WORDS = CT * (E2 + 1) * E1;
__global__ void probe(const __grid_constant__ CUtensorMap tm, u32 *out, int c0, int c3, int c4, long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    u32 *buf = reinterpret_cast<u32 *>(smem);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + BOX_WORDS * 4);
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(bar), buf_a = (unsigned)__cvta_generic_to_shared(buf);
    for (int i = threadIdx.x; i < BOX_WORDS; i += blockDim.x) buf[i] = 0xdeadbeefu;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(BOX_WORDS * 4));
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(buf_a), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(c0), "r"(0), "r"(0), "r"(c3), "r"(c4), "r"(bar_a) : "memory");
    }
    unsigned done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    for (int i = threadIdx.x; i < BOX_WORDS; i += blockDim.x) out[(size_t)blockIdx.x * BOX_WORDS + i] = buf[i];
}
int main() {
    const int log_n = 14, l0 = 2, l1 = 12, w = 100;     // pass over layers [2,12): r = 10, lowbits = 2
    const int lowbits = log_n - l1;
    const size_t h = 1u << log_n;
    std::vector<u32> host(h * w);
    for (size_t i = 0; i < h * w; i++) host[i] = (u32)i;
    u32 *d; cudaMalloc(&d, h * w * 4); cudaMemcpy(d, host.data(), h * w * 4, cudaMemcpyHostToDevice);
    EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qres);
    printf("entry point: %d %d %p\n", (int)e, (int)qres, (void *)enc);
    CUtensorMap tm;
    const cuuint64_t pitch = (cuuint64_t)w * 4;
    cuuint64_t dims[5] = {(cuuint64_t)w, (cuuint64_t)E2, (cuuint64_t)E1, 1ull << lowbits, 1ull << l0};
    cuuint64_t strides[4] = {pitch << lowbits, pitch << (lowbits + 5), pitch, pitch << (log_n - l0)};
    cuuint32_t box[5] = {CT, E2 + 1, E1, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int nb = 3;
    u32 *out; cudaMalloc(&out, (size_t)nb * BOX_WORDS * 4);
    long long *cyc; cudaMalloc(&cyc, nb * 8);
    const int smem = BOX_WORDS * 4 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // tile: column tile at c0 = 96 (ragged: 4 real columns), L = 3, T = 2
    const int c0 = 96, L = 3, T = 2;
    probe<<<nb, 128, smem>>>(tm, out, c0, L, T, cyc);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<u32> res((size_t)nb * BOX_WORDS); cudaMemcpy(res.data(), out, res.size() * 4, cudaMemcpyDeviceToHost);
    long long hc[nb]; cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int G = 0; G < E1; G++) for (int mu = 0; mu <= E2; mu++) for (int c = 0; c < CT; c++) {
        const u32 got = res[((size_t)G * (E2 + 1) + mu) * CT + c];
        u32 want = 0;
        if (mu < E2 && c0 + c < w) {
            const size_t rho = (size_t)G * E2 + mu, row = ((size_t)T << (log_n - l0)) | (rho << lowbits) | L;
            want = host[row * w + c0 + c];
        }
        if (got != want && bad++ < 5) printf("mismatch G=%d mu=%d c=%d got %u want %u\n", G, mu, c, got, want);
    }
    printf("mismatches: %zu   load cycles: %lld %lld %lld\n", bad, hc[0], hc[1], hc[2]);
    return bad != 0;
}
