// Host execution of the DEVICE permutations in plonky3_b200/csrc/hash_core.cuh (Poseidon2 with looped rounds, lazy S-box squares and
// the shift-based internal diagonal; Keccak-f on 32-bit halves), compiled as plain C++.  A filter: reads jobs from stdin, prints the
// permuted states; tests/test_abi.py compares them with the CPU oracle and the reference's known-answer vectors.
//   p2 <field> <width> <rounds_p>  <8*width rc_ext>  <rounds_p rc_int>  <n>  <n*width state words (Montgomery)>
//   keccak <n>  <25*n 64-bit words>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    return (unsigned)(((((unsigned long long)hi << 32) | lo) << (shift & 31)) >> 32);
}
#include "../../plonky3_b200/csrc/hash_core.cuh"
using namespace p3;

template <int F, int W> static void run_p2(const Poseidon2Consts &k, std::vector<u32> &st) {
    for (size_t i = 0; i + W <= st.size(); i += W) {
        u32 s[W];
        memcpy(s, &st[i], sizeof s);
        poseidon2_permute<F, W>(s, k);
        memcpy(&st[i], s, sizeof s);
    }
}

int main() {
    char cmd[16];
    while (scanf("%15s", cmd) == 1) {
        if (!strcmp(cmd, "p2")) {
            int field, width, rounds_p; size_t n;
            if (scanf("%d %d %d", &field, &width, &rounds_p) != 3) return 2;
            Poseidon2Consts k; memset(&k, 0, sizeof k);
            k.rounds_p = rounds_p; k.width = width; k.set = 1;
            for (int i = 0; i < 8 * width; i++) if (scanf("%u", &k.rc_ext[i]) != 1) return 2;
            for (int i = 0; i < rounds_p; i++) if (scanf("%u", &k.rc_int[i]) != 1) return 2;
            if (scanf("%zu", &n) != 1) return 2;
            std::vector<u32> st(n * width);
            for (auto &v : st) if (scanf("%u", &v) != 1) return 2;
            if (field == 0 && width == 16) run_p2<BABY_BEAR, 16>(k, st);
            else if (field == 0 && width == 24) run_p2<BABY_BEAR, 24>(k, st);
            else if (field == 1 && width == 16) run_p2<KOALA_BEAR, 16>(k, st);
            else if (field == 1 && width == 24) run_p2<KOALA_BEAR, 24>(k, st);
            else return 3;
            for (auto v : st) printf("%u\n", v);
        } else if (!strcmp(cmd, "keccak")) {
            size_t n;
            if (scanf("%zu", &n) != 1) return 2;
            for (size_t j = 0; j < n; j++) {
                KState s;
                for (int i = 0; i < 25; i++) { unsigned long long v; if (scanf("%llu", &v) != 1) return 2; s.lo[i] = (u32)v; s.hi[i] = (u32)(v >> 32); }
                keccak_f(s);
                for (int i = 0; i < 25; i++) printf("%llu\n", (unsigned long long)s.lo[i] | ((unsigned long long)s.hi[i] << 32));
            }
        } else return 4;
    }
    return 0;
}
