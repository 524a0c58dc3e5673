// Host-only check of p3gpu::prune_paths (include/p3gpu.hpp): reads "levels n" then n lines "index w0 .. w(levels*8-1)" (hex words),
// prints the pruned digests.  tests/test_abi.py compares with plonky3_b200.merkle_tree.prune_paths (pinned on the reference's fixture).
#include <cstdio>
#include "p3gpu.hpp"
int main() {
    size_t levels = 0, n = 0;
    if (std::scanf("%zu %zu", &levels, &n) != 2) return 2;
    std::vector<uint32_t> idx(n), paths(n * levels * 8);
    for (size_t q = 0; q < n; q++) {
        if (std::scanf("%u", &idx[q]) != 1) return 2;
        for (size_t k = 0; k < levels * 8; k++) if (std::scanf("%x", &paths[q * levels * 8 + k]) != 1) return 2;
    }
    const auto out = p3gpu::prune_paths(idx, paths, levels);
    for (size_t k = 0; k < out.size(); k++) std::printf("%08x%c", out[k], (k % 8 == 7) ? '\n' : ' ');
    return 0;
}
