// Compile/link check of the C++ host mirror (include/p3gpu.hpp) against libp3gpu.so.  On a box without a CUDA device
// the context constructor must fail loudly (no CPU fallback); with a device it runs one tiny DFT + commit.
#include <cstdio>
#include "p3gpu.hpp"
int main() {
    try {
        p3gpu::Context ctx(0);
        p3gpu::Radix2DitParallel dft(ctx, P3GPU_KOALA_BEAR);
        p3gpu::RowMajorMatrix m; m.width = 3; m.values.assign(8 * 3, 0u); m.values[0] = 0x01fffffeu;  // delta in column 0 (Montgomery ONE)
        auto r = dft.dft_batch(m);
        for (size_t i = 0; i < 8; i++) if (r.values[i * 3] != 0x01fffffeu || r.values[i * 3 + 1] != 0) { std::puts("FAIL dft"); return 2; }
        p3gpu::MerkleTreeMmcs mmcs(ctx, P3GPU_BABY_BEAR, P3GPU_HASH_KECCAK, 0);
        auto ct = mmcs.commit({&r});
        std::printf("gpu ok: %zu layers, cap words %zu\n", ct.second.digest_layers.size(), ct.first.size());
        return 0;
    } catch (const p3gpu::Error &e) {
        std::printf("no device: %s\n", e.what());
        return 3;
    }
}
