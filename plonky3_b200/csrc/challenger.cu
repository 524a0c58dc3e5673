// Device-resident Fiat-Shamir transcript: DuplexChallenger<F, Poseidon2, WIDTH, RATE> (challenger/src/duplex_challenger.rs:60-114,
// 168-300) and its proof-of-work grinding (challenger/src/grinding_challenger.rs:100-232).
//
// The reference keeps the transcript on the host.  With the prover's data resident on the GPU the values it has to absorb (Merkle
// caps, thousands of opened values) are produced on the device, so the sponge lives there too: its state, input and output
// buffers sit in device memory, `observe` is one single-thread kernel walking a device (or staged host) slice through the sponge —
// a few microseconds per permutation, no PCIe round trip per duplexing — and `sample` copies squeezed elements back.  Grinding is a
// parallel search over candidate witnesses (every thread permutes `transcript || candidate`, the smallest valid witness wins —
// what a serial reference build returns; parallel builds may return any valid one, SURVEY 8c "Determinism caveat").
// This is protocol plumbing for the config-5 prove driver, not part of the hot path.
#include "common.h"
#include "hash_core.cuh"

struct p3gpu_challenger {
    int field, width, rate;
    p3::u32 *state;     // device: [0, width) sponge state | [32, 32+rate) input buffer | [64, 64+rate) output buffer | [96] n_in | [97] n_out
    p3::u32 *stage;     // device staging for host observes / samples (4096 words)
};

namespace p3 {

constexpr int CH_IN = 32, CH_OUT = 64, CH_NIN = 96, CH_NOUT = 97, CH_WORDS = 128, CH_STAGE = 4096;

template <int F, int W>
__device__ void ch_duplexing(u32 *st, int rate, const Poseidon2Consts &k) {
    const u32 n = st[CH_NIN];
    u32 s[W];
#pragma unroll
    for (int i = 0; i < W; i++) s[i] = st[i];
    if (n > 0) {
#pragma unroll
        for (int i = 0; i < W; i++)
            if (i < rate) s[i] = (u32)i < n ? st[CH_IN + i] : 0u;           // overwrite the leading rate slots, clear the rest of the rate
        s[rate] = fp_add<F>(s[rate], to_monty<F>(n));                       // bind the absorbed length into the first capacity element
    }
    poseidon2_permute<F, W>(s, k);
#pragma unroll
    for (int i = 0; i < W; i++) st[i] = s[i];
    for (int i = 0; i < rate; i++) st[CH_OUT + i] = s[i];
    st[CH_NIN] = 0; st[CH_NOUT] = (u32)rate;
}

template <int F, int W>
__global__ void ch_observe_kernel(u32 *st, int rate, const u32 *vals, size_t n, const __grid_constant__ Poseidon2Consts k) {
    if (threadIdx.x | blockIdx.x) return;
    for (size_t j = 0; j < n; j++) {
        st[CH_NOUT] = 0;                                                      // any buffered output is now invalid
        const u32 m = st[CH_NIN];
        st[CH_IN + m] = vals[j];
        st[CH_NIN] = m + 1;
        if (m + 1 == (u32)rate) ch_duplexing<F, W>(st, rate, k);
    }
}

template <int F, int W>
__global__ void ch_sample_kernel(u32 *st, int rate, u32 *out, size_t n, const __grid_constant__ Poseidon2Consts k) {
    if (threadIdx.x | blockIdx.x) return;
    for (size_t j = 0; j < n; j++) {
        if (st[CH_NIN] != 0 || st[CH_NOUT] == 0) ch_duplexing<F, W>(st, rate, k);
        const u32 m = st[CH_NOUT] - 1;                                        // samples pop from the END of the output buffer
        out[j] = st[CH_OUT + m];
        st[CH_NOUT] = m;
    }
}

// candidate c is valid iff the sample after observing it has `bits` trailing zero bits (canonical value); best = smallest valid c
template <int F, int W>
__global__ void __launch_bounds__(128) ch_grind_kernel(const u32 *st, int rate, u32 base, u32 count, u32 mask, u32 *best, const __grid_constant__ Poseidon2Consts k) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const u32 cand = base + t;
    if (cand >= Fp<F>::P) return;
    const u32 widx = st[CH_NIN];
    u32 s[W];
#pragma unroll
    for (int i = 0; i < W; i++) {
        if (i < rate) s[i] = (u32)i < widx ? st[CH_IN + i] : 0u;
        else s[i] = st[i];
    }
#pragma unroll
    for (int i = 0; i < W; i++) if ((u32)i == widx) s[i] = to_monty<F>(cand);
    s[rate] = fp_add<F>(s[rate], to_monty<F>(widx + 1));
    poseidon2_permute<F, W>(s, k);
    u32 last = 0;
#pragma unroll
    for (int i = 0; i < W; i++) if (i == rate - 1) last = s[i];
    if ((from_monty<F>(last) & mask) == 0) atomicMin(best, cand);
}

template <typename Fn> static int32_t ch_dispatch(int field, int width, Fn &&fn) {
    if (field == BABY_BEAR && width == 16) return fn(std::integral_constant<int, BABY_BEAR>(), std::integral_constant<int, 16>());
    if (field == BABY_BEAR && width == 24) return fn(std::integral_constant<int, BABY_BEAR>(), std::integral_constant<int, 24>());
    if (field == KOALA_BEAR && width == 16) return fn(std::integral_constant<int, KOALA_BEAR>(), std::integral_constant<int, 16>());
    return fn(std::integral_constant<int, KOALA_BEAR>(), std::integral_constant<int, 24>());
}

static int32_t ch_consts(p3gpu_ctx *ctx, const p3gpu_challenger *ch, const Poseidon2Consts **k) {
    *k = &ctx->p2_host[ch->field][ch->width == 24];
    P3_CHECK((*k)->set, P3GPU_ESTATE, "Poseidon2 constants for field %d width %d not set (p3gpu_poseidon2_set_constants)", ch->field, ch->width);
    return P3GPU_OK;
}

int32_t challenger_new(p3gpu_ctx *ctx, int field, int width, int rate, p3gpu_challenger **out) {
    P3_CHECK(field == BABY_BEAR || field == KOALA_BEAR, P3GPU_EUNSUPPORTED, "unknown field %d", field);
    P3_CHECK(width == 16 || width == 24, P3GPU_EUNSUPPORTED, "challenger permutation width %d unsupported (16 or 24)", width);
    P3_CHECK(rate > 0 && rate < width && rate <= 24, P3GPU_EINVAL, "challenger rate %d out of range", rate);
    p3gpu_challenger *ch = new p3gpu_challenger();
    ch->field = field; ch->width = width; ch->rate = rate;
    if (cudaMalloc(&ch->state, (CH_WORDS + CH_STAGE) * 4) != cudaSuccess) { delete ch; cudaGetLastError(); set_error("cudaMalloc failed"); return P3GPU_ENOMEM; }
    ch->stage = ch->state + CH_WORDS;
    P3_CUDA(cudaMemsetAsync(ch->state, 0, CH_WORDS * 4, ctx->stream));
    *out = ch;
    return P3GPU_OK;
}
void challenger_free(p3gpu_ctx *ctx, p3gpu_challenger *ch) {
    if (!ch) return;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ch->state);
    delete ch;
}
int32_t challenger_clone(p3gpu_ctx *ctx, const p3gpu_challenger *src, p3gpu_challenger **out) {
    P3_TRY(challenger_new(ctx, src->field, src->width, src->rate, out));
    P3_CUDA(cudaMemcpyAsync((*out)->state, src->state, CH_WORDS * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return P3GPU_OK;
}

int32_t challenger_observe_dev(p3gpu_ctx *ctx, p3gpu_challenger *ch, const u32 *d_vals, size_t n) {
    if (n == 0) return P3GPU_OK;
    const Poseidon2Consts *k;
    P3_TRY(ch_consts(ctx, ch, &k));
    P3_TRY(ch_dispatch(ch->field, ch->width, [&](auto f, auto w) -> int32_t {
        ch_observe_kernel<decltype(f)::value, decltype(w)::value><<<1, 1, 0, ctx->stream>>>(ch->state, ch->rate, d_vals, n, *k);
        return P3GPU_OK;
    }));
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    return P3GPU_OK;
}
int32_t challenger_observe_host(p3gpu_ctx *ctx, p3gpu_challenger *ch, const u32 *h_vals, size_t n) {
    const u32 p = ch->field == BABY_BEAR ? Fp<BABY_BEAR>::P : Fp<KOALA_BEAR>::P;
    for (size_t i = 0; i < n; i++) P3_CHECK(h_vals[i] < p, P3GPU_EINVAL, "observed value not in canonical Montgomery range");
    for (size_t off = 0; off < n; off += CH_STAGE) {
        const size_t m = std::min<size_t>(CH_STAGE, n - off);
        P3_CUDA(cudaMemcpyAsync(ch->stage, h_vals + off, m * 4, cudaMemcpyHostToDevice, ctx->stream));   // pageable source: staged before return
        P3_TRY(challenger_observe_dev(ctx, ch, ch->stage, m));
    }
    return P3GPU_OK;
}
int32_t challenger_sample(p3gpu_ctx *ctx, p3gpu_challenger *ch, u32 *h_out, size_t n) {
    P3_CHECK(n <= (size_t)CH_STAGE, P3GPU_EINVAL, "too many samples in one call");
    if (n == 0) return P3GPU_OK;
    const Poseidon2Consts *k;
    P3_TRY(ch_consts(ctx, ch, &k));
    P3_TRY(ch_dispatch(ch->field, ch->width, [&](auto f, auto w) -> int32_t {
        ch_sample_kernel<decltype(f)::value, decltype(w)::value><<<1, 1, 0, ctx->stream>>>(ch->state, ch->rate, ch->stage, n, *k);
        return P3GPU_OK;
    }));
    ctx->launches++;
    P3_CUDA(cudaGetLastError());
    P3_CUDA(cudaMemcpyAsync(h_out, ch->stage, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    P3_CUDA(cudaStreamSynchronize(ctx->stream));
    return P3GPU_OK;
}
// grind(bits): smallest witness w (canonical integer; returned in Montgomery form) such that observe(w); sample_bits(bits) == 0.
// The witness is observed and the sample consumed, exactly like check_witness (grinding_challenger.rs:226-229).
int32_t challenger_grind(p3gpu_ctx *ctx, p3gpu_challenger *ch, unsigned bits, u32 *witness_monty) {
    P3_CHECK(bits < 31, P3GPU_EINVAL, "proof-of-work bits %u too large", bits);
    const u32 p = ch->field == BABY_BEAR ? Fp<BABY_BEAR>::P : Fp<KOALA_BEAR>::P;
    if (bits == 0) { *witness_monty = 0; return P3GPU_OK; }
    const Poseidon2Consts *k;
    P3_TRY(ch_consts(ctx, ch, &k));
    u32 *best = ch->stage + CH_STAGE - 1;
    const u32 mask = (1u << bits) - 1u, batch = 1u << std::min(20u, bits + 3);
    u32 found = 0xffffffffu;
    for (u64 base = 0; base < p && found == 0xffffffffu; base += batch) {
        P3_CUDA(cudaMemsetAsync(best, 0xff, 4, ctx->stream));
        P3_TRY(ch_dispatch(ch->field, ch->width, [&](auto f, auto w) -> int32_t {
            ch_grind_kernel<decltype(f)::value, decltype(w)::value><<<(batch + 127) / 128, 128, 0, ctx->stream>>>(ch->state, ch->rate, (u32)base, batch, mask, best, *k);
            return P3GPU_OK;
        }));
        ctx->launches++;
        P3_CUDA(cudaGetLastError());
        P3_CUDA(cudaMemcpyAsync(&found, best, 4, cudaMemcpyDeviceToHost, ctx->stream));
        P3_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    P3_CHECK(found != 0xffffffffu, P3GPU_EINVAL, "failed to find proof-of-work witness");
    const u32 wm = ch->field == BABY_BEAR ? to_monty<BABY_BEAR>(found) : to_monty<KOALA_BEAR>(found);
    P3_TRY(challenger_observe_host(ctx, ch, &wm, 1));
    u32 s = 0;
    P3_TRY(challenger_sample(ctx, ch, &s, 1));
    const u32 canon = ch->field == BABY_BEAR ? from_monty<BABY_BEAR>(s) : from_monty<KOALA_BEAR>(s);
    P3_CHECK((canon & mask) == 0, P3GPU_ECUDA, "proof-of-work witness failed the check");
    *witness_monty = wm;
    return P3GPU_OK;
}

}  // namespace p3
