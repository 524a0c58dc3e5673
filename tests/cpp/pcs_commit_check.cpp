// GPU run of the C++ host mirror (include/p3gpu.hpp): TwoAdicFriPcs::commit of a seeded trace + open_multi_batch with pruning.
// Prints the cap, the opened rows and the pruned multiproof as hex words; tests/test_gpu_parity.py compares them with the oracle.
//   usage: pcs_commit_check <field> <hash> <log_h> <w> <log_blowup> <cap_height> <idx,idx,...>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "p3gpu.hpp"

static void dump(const char *name, const std::vector<uint32_t> &v) {
    std::printf("%s", name);
    for (uint32_t x : v) std::printf(" %08x", x);
    std::printf("\n");
}

int main(int argc, char **argv) {
    if (argc != 8) { std::puts("usage"); return 2; }
    const int field = std::atoi(argv[1]), hash = std::atoi(argv[2]);
    const unsigned log_h = (unsigned)std::atoi(argv[3]), log_blowup = (unsigned)std::atoi(argv[5]);
    const size_t w = (size_t)std::atoll(argv[4]), cap_height = (size_t)std::atoll(argv[6]);
    std::vector<uint32_t> indices;
    for (char *tok = std::strtok(argv[7], ","); tok; tok = std::strtok(nullptr, ",")) indices.push_back((uint32_t)std::strtoul(tok, nullptr, 10));
    const uint64_t p = field == P3GPU_BABY_BEAR ? 0x78000001ull : 0x7f000001ull;
    try {
        p3gpu::Context ctx(0);
        p3gpu::RowMajorMatrix m;
        m.width = w;
        m.values.resize(((size_t)1 << log_h) * w);
        uint64_t s = 12345;                                           // the same LCG the Python side runs
        for (auto &v : m.values) { s = (s * 6364136223846793005ull + 1442695040888963407ull); v = (uint32_t)((s >> 33) % p); }
        if (hash != P3GPU_HASH_KECCAK) {                              // default-constant Poseidon2 is NOT implied: constants come from the caller
            std::puts("only the Keccak MMCS needs no constants; use hash 2"); return 2;
        }
        p3gpu::TwoAdicFriPcs pcs(ctx, field, hash, log_blowup, cap_height);
        auto cd = pcs.commit(m);
        dump("cap", cd.first);
        auto op = pcs.open_multi_batch(indices, cd.second);
        for (auto &row : op.opened_values) dump("row", row);
        dump("pruned", op.pruned_digests);
        return 0;
    } catch (const p3gpu::Error &e) {
        std::printf("error: %s\n", e.what());
        return 3;
    }
}
