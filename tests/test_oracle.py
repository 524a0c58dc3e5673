"""Pins the CPU oracle (oracle/p3_oracle.c) against the reference's own known answers.

Everything here runs on CPU.  The oracle is what the GPU parity tests compare against, so it is
checked first against: field KATs, the two-adic generator tables, NaiveDft, Poseidon2 KATs, FIPS-202
(Keccak), structural Merkle identities from the reference's tests, and the committed proof fixture.
"""
import hashlib
import json
import pathlib

import numpy as np
import pytest

from oracle import p3_oracle as O
import fixture_replay as FR

GOLD = pathlib.Path(__file__).resolve().parent / "golden"
BB, KB = O.BABY_BEAR, O.KOALA_BEAR


def test_field_mul_kat():
    # baby-bear/src/baby_bear.rs:178-181, koala-bear/src/koala_bear.rs:182-185
    for f, exp in ((BB, 0x1B5C8046), (KB, 0x54B46B81)):
        m1, m2 = O.to_monty(f, 0x34167C58), O.to_monty(f, 0x61F3207B)
        assert O.from_monty(f, O.mul(f, m1, m2)) == exp


def test_monty_constants():
    # SURVEY Appendix A (derived from baby_bear.rs:17-20, koala_bear.rs:20-23)
    assert O.to_monty(BB, 1) == 0x0FFFFFFE and O.to_monty(KB, 1) == 0x01FFFFFE
    for f in (BB, KB):
        p = O.prime(f)
        for a, b in ((0, 0), (1, p - 1), (p - 1, p - 1), (12345, 678910)):
            am, bm = O.to_monty(f, a), O.to_monty(f, b)
            assert O.from_monty(f, O.add(f, am, bm)) == (a + b) % p
            assert O.from_monty(f, O.sub(f, am, bm)) == (a - b) % p
            assert O.from_monty(f, O.mul(f, am, bm)) == a * b % p
            assert O.from_monty(f, O.halve(f, am)) == a * pow(2, p - 2, p) % p


def test_two_adic_generators():
    # baby_bear.rs:48-53,150-158 ; koala_bear.rs:73-78,163-171
    gens = json.loads((GOLD / "two_adic_generators.json").read_text())
    for f, name in ((BB, "baby_bear"), (KB, "koala_bear")):
        for bits, g in enumerate(gens[name]):
            assert O.from_monty(f, O.two_adic_generator(f, bits)) == g


@pytest.mark.parametrize("f", [BB, KB])
@pytest.mark.parametrize("log_h,w", [(0, 3), (1, 2), (3, 5), (6, 3), (8, 1)])
def test_dft_matches_naive(f, log_h, w):
    # dft/tests/testing.rs:298-378: every backend must agree with NaiveDft
    m = O.random_matrix(f, 1 << log_h, w, seed=log_h * 10 + w)
    assert np.array_equal(O.dft_batch(f, m), O.naive_dft(f, m))


def test_naive_dft_literal():
    # dft/src/naive.rs:46-85 style: DFT of a delta is all-ones; DFT of ones is h*delta
    f = BB
    one = O.to_monty(f, 1)
    m = np.zeros((8, 1), dtype=np.uint32); m[0, 0] = one
    assert (O.naive_dft(f, m) == one).all()
    m[:] = one
    out = O.naive_dft(f, m)
    assert out[0, 0] == O.to_monty(f, 8) and (out[1:] == 0).all()


@pytest.mark.parametrize("f", [BB, KB])
def test_dft_roundtrips_and_cosets(f):
    # dft/tests/testing.rs:380-452 (round trips), traits.rs:84-155 (coset definitions)
    m = O.random_matrix(f, 64, 7, seed=3)
    assert np.array_equal(O.idft_batch(f, O.dft_batch(f, m)), m)
    s = O.to_monty(f, 0x1234567)
    assert np.array_equal(O.coset_idft_batch(f, O.coset_dft_batch(f, m, s), s), m)
    # coset dft by definition: evaluate the polynomial at s*w^i
    p = O.prime(f); co = O.from_monty_arr(f, m[:, 0]).astype(object)
    w = O.from_monty(f, O.two_adic_generator(f, 6)); sc = O.from_monty(f, s)
    ev = O.from_monty_arr(f, O.coset_dft_batch(f, m, s)[:, 0])
    for i in (0, 1, 17, 63):
        x = sc * pow(w, i, p) % p
        assert int(ev[i]) == sum(int(c) * pow(x, j, p) for j, c in enumerate(co)) % p


@pytest.mark.parametrize("f", [BB, KB])
@pytest.mark.parametrize("added_bits", [0, 1, 2, 3])
def test_coset_lde_layout(f, added_bits):
    # traits.rs:227-259 (definition) + radix_2_dit_parallel.rs:181-246 / two_adic_pcs.rs:313-318 (bit-reversed rows)
    h, w = 16, 3
    m = O.random_matrix(f, h, w, seed=7)
    shift = O.generator(f)
    nat = O.coset_lde_batch(f, m, added_bits, shift, bitrev_out=False)
    brv = O.coset_lde_batch(f, m, added_bits, shift, bitrev_out=True)
    assert np.array_equal(O.reverse_matrix_index_bits(nat), brv)
    # definition: idft, zero-pad, coset dft
    co = O.idft_batch(f, m)
    pad = np.zeros((h << added_bits, w), dtype=np.uint32); pad[:h] = co
    assert np.array_equal(O.coset_dft_batch(f, pad, shift), nat)
    # first h memory rows of the bit-reversed LDE are the evaluations on shift*H (bit-reversed)
    assert np.array_equal(brv[:h], O.reverse_matrix_index_bits(O.coset_dft_batch(f, co, shift)))


def test_poseidon2_kats():
    # koala-bear/src/poseidon2.rs:614-653, baby-bear/src/poseidon2.rs:599-639
    kats = json.loads((GOLD / "poseidon2_kat.json").read_text())
    for f, name in ((BB, "baby_bear"), (KB, "koala_bear")):
        for w in (16, 24):
            k = kats[f"{name}_{w}"]
            out = O.poseidon2_permute(O.default_perm(f, w), O.to_monty_arr(f, k["input"]))
            assert O.from_monty_arr(f, out).tolist() == k["expected"]


def test_poseidon2_diag_values():
    # koala-bear/src/poseidon2.rs:410-428: V16 = [-2,1,2,1/2,3,4,-1/2,-3,-4,2^-8,1/8,2^-24,-2^-8,-1/8,-1/16,-2^-24]
    p = O.prime(KB); i2 = lambda k: pow(pow(2, k, p), p - 2, p)
    exp = [p - 2, 1, 2, i2(1), 3, 4, p - i2(1), p - 3, p - 4, i2(8), i2(3), i2(24), p - i2(8), p - i2(3), p - i2(4), p - i2(24)]
    assert O.from_monty_arr(KB, O.poseidon2_diag(KB, 16)).tolist() == exp


def _sha3_256_via_oracle(msg: bytes) -> bytes:
    rate = 136
    padded = bytearray(msg) + b"\x06" + b"\x00" * ((-len(msg) - 2) % rate) + b"\x80" if (len(msg) + 1) % rate else bytearray(msg) + b"\x86"
    st = np.zeros(25, dtype=np.uint64)
    for off in range(0, len(padded), rate):
        blk = np.frombuffer(bytes(padded[off:off + rate]), dtype="<u8")
        st[:17] ^= blk
        st = O.keccak_f(st)
    return st[:4].astype("<u8").tobytes()


def test_keccak_f_fips202():
    # tiny-keccak is not vendored; pin Keccak-f[1600] with FIPS-202 SHA3-256 (hashlib)
    for msg in (b"", b"abc", bytes(range(200)), b"x" * 135, b"y" * 136):
        assert _sha3_256_via_oracle(msg) == hashlib.sha3_256(msg).digest()


def test_keccak_leaf_packing_and_compress():
    # field/src/integers.rs:494-509 (pair packing), sponge.rs:182-216 (overwrite, no padding), compression.rs:60-70
    hs = O.keccak_hasher()
    row = np.arange(1, 38, dtype=np.uint32)                       # odd width 37 -> 19 words -> 2 permutations
    words = [int(row[2 * i]) | (int(row[2 * i + 1]) << 32) for i in range(18)] + [int(row[36])]
    st = np.zeros(25, dtype=np.uint64)
    st[:17] = words[:17]; st = O.keccak_f(st)
    st[:2] = words[17:]; st = O.keccak_f(st)
    assert O.hash_row(hs, row).view(np.uint64).tolist() == st[:4].tolist()
    l, r = O.hash_row(hs, row), O.hash_row(hs, row[:5])
    st = np.zeros(25, dtype=np.uint64); st[:4] = l.view(np.uint64); st[4:8] = r.view(np.uint64)
    assert O.compress(hs, l, r).view(np.uint64).tolist() == O.keccak_f(st)[:4].tolist()


def _p2_hasher(f=KB, wl=16):
    return O.poseidon2_hasher(O.default_perm(f, wl), O.default_perm(f, 16))


def test_sponge_semantics():
    # symmetric/src/sponge.rs:182-216: overwrite mode, partial last block keeps previous outputs, no padding
    hs = _p2_hasher()
    pm = O.default_perm(KB, 16)
    row = O.random_matrix(KB, 1, 11, seed=5)[0]
    st = np.zeros(16, dtype=np.uint32)
    st[:8] = row[:8]; st = O.poseidon2_permute(pm, st)
    st[:3] = row[8:]; st = O.poseidon2_permute(pm, st)
    assert np.array_equal(O.hash_row(hs, row), st[:8])
    # exact multiple of the rate: no extra permutation
    st = np.zeros(16, dtype=np.uint32); st[:8] = row[:8]
    assert np.array_equal(O.hash_row(hs, row[:8]), O.poseidon2_permute(pm, st)[:8])
    # empty input: zero digest
    assert (O.hash_row(hs, row[:0]) == 0).all()


def test_merkle_structure_single_matrix():
    # merkle-tree/src/mmcs/batch.rs:333-364: root == compress(compress(h,h), compress(h,h))
    hs = _p2_hasher()
    m = O.random_matrix(KB, 4, 9, seed=2)
    layers = O.merkle_tree(hs, [m])
    h = [O.hash_row(hs, m[i]) for i in range(4)]
    root = O.compress(hs, O.compress(hs, h[0], h[1]), O.compress(hs, h[2], h[3]))
    assert [len(l) for l in layers] == [4, 2, 1]
    assert np.array_equal(layers[-1][0], root)
    assert np.array_equal(O.merkle_cap(layers, 1), layers[1])
    assert np.array_equal(O.merkle_cap(layers, 5), layers[0])       # clamp (mmcs/batch.rs:56-62)


def test_merkle_mixed_heights_and_padding():
    # merkle_tree.rs:348-460 (inject), :652-711 (zero-digest padding); two tallest matrices share a leaf hash
    hs = _p2_hasher()
    a, b = O.random_matrix(KB, 8, 3, seed=1), O.random_matrix(KB, 8, 2, seed=2)
    c = O.random_matrix(KB, 4, 5, seed=3)
    layers = O.merkle_tree(hs, [c, a, b])                            # input order: c first, but a,b are tallest
    leaf = [O.hash_row(hs, np.concatenate([a[i], b[i]])) for i in range(8)]
    l1 = [O.compress(hs, O.compress(hs, leaf[2 * i], leaf[2 * i + 1]), O.hash_row(hs, c[i])) for i in range(4)]
    l2 = [O.compress(hs, l1[0], l1[1]), O.compress(hs, l1[2], l1[3])]
    assert np.array_equal(layers[1], np.array(l1)) and np.array_equal(layers[2], np.array(l2))
    # non power of two height 5 -> padded 6 -> 3 -> padded 4 -> 2 -> 1
    m = O.random_matrix(KB, 5, 4, seed=4)
    layers = O.merkle_tree(hs, [m])
    assert [len(l) for l in layers] == [6, 4, 2, 1]
    z = np.zeros(8, dtype=np.uint32)
    assert (layers[0][5] == 0).all()
    assert np.array_equal(layers[1][2], O.compress(hs, O.hash_row(hs, m[4]), z))
    assert (layers[1][3] == 0).all()
    with pytest.raises(ValueError):
        O.merkle_tree(hs, [O.random_matrix(KB, 8, 1), O.random_matrix(KB, 3, 1)])   # 3 is off the ladder of 8


def test_fold_matrix_matches_interpolation():
    # fri/src/two_adic_pcs.rs:108-131 (fold_row by Lagrange interpolation) == fold_matrix, arity 2/4/8
    f = KB; p = O.prime(f)
    rng = np.random.default_rng(11)
    for log_arity in (1, 2, 3):
        log_len = 6
        deg = 1 << log_len
        # a polynomial with EF4 coefficients, evaluated on the subgroup in bit-reversed order
        coef = rng.integers(0, p, size=(deg, 4), dtype=np.uint32)
        ev = O.dft_batch(f, coef)                                    # each EF coordinate transformed independently
        ev = O.reverse_matrix_index_bits(ev)
        beta = rng.integers(0, p, size=4, dtype=np.uint32)
        out = O.fold_matrix(f, ev, log_arity, beta)
        # folding by arity a maps f(x) -> sum_k beta^k f_k(x^a), f_k = coefficients k mod a
        a = 1 << log_arity
        bp = [np.array([O.to_monty(f, 1), 0, 0, 0], dtype=np.uint32)]
        for _ in range(a - 1): bp.append(O.ef_mul(f, bp[-1], beta))
        folded = np.zeros((deg // a, 4), dtype=np.uint32)
        for j in range(deg // a):
            acc = np.zeros(4, dtype=np.uint32)
            for k in range(a):
                t = O.ef_mul(f, bp[k], coef[j * a + k])
                acc = np.array([O.add(f, int(x), int(y)) for x, y in zip(acc, t)], dtype=np.uint32)
            folded[j] = acc
        exp = O.reverse_matrix_index_bits(O.dft_batch(f, folded))
        assert np.array_equal(out, exp)


class OracleBackend:
    """fixture_replay backend built on the C oracle."""

    def __init__(self):
        rc_i, rc_t, rc_p = FR.fixture_constants()
        pm = O.make_perm(BB, 16, rc_i, rc_t, rc_p, monty=True)
        self.hs = O.poseidon2_hasher(pm, pm)

    def lde(self, mat, added_bits, shift): return O.coset_lde_batch(BB, mat, added_bits, shift, bitrev_out=True)
    def commit(self, mats): return O.merkle_cap(O.merkle_tree(self.hs, mats), 0)
    def fold(self, vec, log_arity, beta): return O.fold_matrix(BB, vec, log_arity, beta)

    def commit_data(self, mats):
        layers = O.merkle_tree(self.hs, mats)
        return O.merkle_cap(layers, 0), (mats, layers)

    def open_multi(self, data, indices):                          # rows per matrix + full sibling paths (mmcs/mod.rs:334-414)
        mats, layers = data
        rows = [np.array([m[i] for i in indices], dtype=np.uint32) for m in mats]
        paths = np.array([[layers[l][(i >> l) ^ 1] for l in range(len(layers) - 1)] for i in indices], dtype=np.uint32)
        return rows, paths


class OracleOpenBackend(OracleBackend):
    """Adds TwoAdicFriPcs::open's pre-FRI part, restated with the oracle primitives (two_adic_pcs.rs:413-662)."""

    def open(self, rounds, challenger, log_blowup):
        f = BB
        all_opened, inv_d = [], {}
        for mats, points in rounds:
            for m, pts in zip(mats, points):
                for z in pts:
                    inv_d[tuple(z)] = O.open_inv_denoms(f, int(np.log2(max(mm.shape[0] for ms, _ in rounds for mm in ms))), z)
        for mats, points in rounds:
            per_mat = []
            for m, pts in zip(mats, points):
                h = m.shape[0] >> log_blowup
                per_pt = []
                for z in pts:
                    ys = O.interpolate_coset(f, m[:h], z, inv_d[tuple(z)])
                    challenger.observe_algebra_slice(ys); per_pt.append(ys)
                per_mat.append(per_pt)
            all_opened.append(per_mat)
        alpha = np.array(challenger.sample_algebra_element(), dtype=np.uint32)
        ro, nred = {}, {}
        for (mats, points), opened in zip(rounds, all_opened):
            for m, pts, om in zip(mats, points, opened):
                H = m.shape[0]
                ro.setdefault(H, np.zeros((H, 4), dtype=np.uint32)); nred.setdefault(H, 0)
                r = O.rowwise_dot(f, m, alpha)
                for z, ys in zip(pts, om):
                    yred = np.zeros(4, dtype=np.uint32); pw = O.ef_from_base(f, O.to_monty(f, 1))
                    for y in ys:
                        yred = O.ef_add(f, yred, O.ef_mul(f, pw, y)); pw = O.ef_mul(f, pw, alpha)
                    ro[H] = O.open_reduce(f, ro[H], r, inv_d[tuple(z)], O.ef_pow(f, alpha, nred[H]), yred)
                    nred[H] += m.shape[1]
        return all_opened, [ro[H] for H in sorted(ro, reverse=True)]


def test_fixture_replay_with_oracle_open():
    """Same fixture, with the opened values and the FRI input produced by the oracle's `open` primitives: pins
    inverse denominators, barycentric interpolation, alpha compression and the quotient accumulation."""
    gold = json.loads((GOLD / "uni_stark_two_adic_v1.json").read_text())
    got = FR.replay(OracleOpenBackend())
    for k, v in got.items():
        assert v == gold[k], k


def test_fixture_replay_with_oracle():
    """LDE + Merkle + FRI of the oracle reproduce the reference's committed proof bit for bit."""
    gold = json.loads((GOLD / "uni_stark_two_adic_v1.json").read_text())
    got = FR.replay(OracleBackend())
    assert set(got) == set(gold) - {"source", "degree_bits"}          # every field of the proof, and its wire form
    for k, v in got.items():
        assert v == gold[k], k


def test_pruned_multiproof_host_logic():
    """prune_paths / restore_paths (merkle-tree/src/pruning.rs) and the postcard reader/writer, host-side product code:
    the reference fixture's own multiproofs restore to paths that authenticate its opened rows against its caps."""
    from plonky3_b200.merkle_tree import prune_paths, restore_paths
    from plonky3_b200.proof_io import proof_from_postcard
    gold = json.loads((GOLD / "uni_stark_two_adic_v1.json").read_text())
    be = OracleBackend()
    p = proof_from_postcard(bytes.fromhex(gold["postcard_hex"]))
    assert p["degree_bits"] == 3 and p["query_pow_witness"] == gold["query_pow_witness"]
    with pytest.raises(ValueError):
        proof_from_postcard(bytes.fromhex(gold["postcard_hex"]) + b"\x00")
    with pytest.raises(ValueError):
        proof_from_postcard(bytes.fromhex(gold["postcard_hex"])[:-7])
    # hand-checkable frontier: 8 leaves, queries {5, 4, 5}: level 0 pairs 4|5 (nothing sent), level 1 needs node 3, level 2 node 0
    rng = np.random.default_rng(0)
    layers = [rng.integers(0, 1 << 31, (8 >> l, 8), dtype=np.uint32) for l in range(4)]
    full = lambda idx: np.array([[layers[l][(i >> l) ^ 1] for l in range(3)] for i in idx], dtype=np.uint32)
    pr = prune_paths([5, 4, 5], full([5, 4, 5]))
    assert np.array_equal(pr, np.array([layers[1][3], layers[2][0]]))
    # random trees: restore(prune(paths)) agrees with the full paths wherever the verifier reads them, and the digest count
    # equals the number of distinct uncovered siblings
    for log_n, nq in [(1, 1), (3, 2), (5, 7), (10, 100), (6, 200)]:
        layers = [rng.integers(0, 1 << 31, ((1 << log_n) >> l, 8), dtype=np.uint32) for l in range(log_n + 1)]
        idx = [int(v) for v in rng.integers(0, 1 << log_n, nq)]
        paths = np.array([[layers[l][(i >> l) ^ 1] for l in range(log_n)] for i in idx], dtype=np.uint32).reshape(nq, log_n, 8)
        pr = prune_paths(idx, paths)
        expect = 0
        for l in range(log_n):
            nodes = {i >> l for i in idx}
            expect += sum(1 for v in nodes if (v ^ 1) not in nodes)
        assert pr.shape == (expect, 8)
        back = restore_paths(idx, pr, log_n)
        mask = back.any(axis=2)
        assert np.array_equal(back[mask], paths[mask]) and int(mask.sum()) == expect
        with pytest.raises(ValueError):
            restore_paths(idx, pr[:-1], log_n)
        with pytest.raises(ValueError):
            restore_paths(idx, np.concatenate([pr, pr[:1]]), log_n)


# ---------------------------------------------------------------- Poseidon2 AIR + prove replay
def test_smallrng_c_matches_fixture_pinned_python():
    """the C SmallRng + field sampler equals the Python one that reproduces the reference's proof fixture (fixture_replay.py)."""
    import fixture_replay as FR
    py = FR.SmallRng(1)
    # FR samples BabyBear (P of the fixture); draw the same stream with the C sampler for BabyBear
    exp = [py.field_monty() for _ in range(300)]
    assert list(O.SmallRng(1).field(0, 300)) == exp


def test_poseidon2_air_trace_and_quotient():
    rng = O.SmallRng(1)
    air = O.air_from_rng(1, rng)
    assert O.p2air_cols(air) == 164 and O.p2air_constraints(air) == 148           # 8 x 164 = 1312 columns
    inputs = O.SmallRng(1).field(1, 64 * 16).reshape(64, 16)
    trace = O.p2air_generate(air, inputs)
    assert trace.shape == (8, 1312) and O.p2air_check(air, trace) == 0
    # the last 16 columns of a permutation are the Poseidon2 output when the AIR constants are used as permutation constants
    pm = O.make_perm(1, 16, np.array(air.beg), np.array(air.end), np.array(air.part)[:20], monty=True)
    for p in (0, 5, 63):
        row = trace.reshape(64, 164)[p]
        assert np.array_equal(O.poseidon2_permute(pm, inputs[p]), row[148:])
    lde = O.coset_lde_batch(1, trace, 1, O.generator(1), True)
    alpha = O.random_matrix(1, 1, 4, seed=3)[0]
    q = O.p2air_quotient(air, lde, 3, alpha)
    assert not O.coset_idft_batch(1, q, O.generator(1))[14:].any()                 # deg Q <= 2N - 2
    bad = trace.copy(); bad[2, 17] ^= 1
    assert O.p2air_check(air, bad) > 0


def test_prove_replay_satisfies_the_verifier_identity():
    import p2_prove_replay as R
    rng = O.SmallRng(1)
    air = O.air_from_rng(1, rng)
    p16 = O.perm_from_rng(1, 16, rng); p24 = O.perm_from_rng(1, 24, rng)
    inputs = O.SmallRng(1).field(1, (8 << 3) * 16).reshape(-1, 16)
    pr = R.prove(air, p16, p24, inputs, num_queries=4, query_pow_bits=4)
    assert pr["log_arities"] == [3] and R.verify_constraints_at_zeta(air, pr)
    broken = dict(pr, alpha=pr["zeta"])
    assert not R.verify_constraints_at_zeta(air, broken)


# ---------------------------------------------------------------- sponge / compression structure vs the reference's mock tests
def _padding_free_sponge(permute, width, rate, out, items):
    """PaddingFreeSponge::hash_iter (symmetric/src/sponge.rs): overwrite the rate, permute after every full or trailing partial
    block, no padding, squeeze the first `out` words."""
    state = [0] * width
    items = list(items)
    for i in range(0, len(items), rate):
        block = items[i:i + rate]
        state[:len(block)] = block
        state = permute(state)
    return state[:out]


def _truncated_permutation(permute, n, chunk, width, inputs):
    """TruncatedPermutation::compress (symmetric/src/compression.rs:40-65): inputs side by side in a zeroed state, permute, truncate."""
    state = [0] * width
    for k, part in enumerate(inputs):
        state[k * chunk:(k + 1) * chunk] = part
    return permute(state)[:chunk]


def test_sponge_and_compression_structure_vs_reference_mock_tests():
    """The reference pins the sponge and the compression function with a plain-sum mock permutation (symmetric/src/sponge.rs:708-776,
    compression.rs:113-179).  The generic restatements above reproduce those literals; the C oracle's Poseidon2 leaf hash and node
    compression (what the GPU is compared with) equal the same generic code driven by the oracle's permutation."""
    mock = lambda st: [sum(st)] * len(st)
    assert _padding_free_sponge(mock, 4, 2, 2, [1, 2, 3, 4, 5]) == [44, 44]
    assert _padding_free_sponge(mock, 4, 2, 2, []) == [0, 0]
    assert _padding_free_sponge(mock, 6, 3, 2, [10, 20, 30]) == [60, 60]
    assert _truncated_permutation(mock, 2, 4, 8, [[1, 2, 3, 4], [5, 6, 7, 8]]) == [36] * 4
    assert _truncated_permutation(mock, 2, 4, 8, [[0] * 4, [0] * 4]) == [0] * 4
    assert _truncated_permutation(mock, 2, 3, 10, [[1, 2, 3], [4, 5, 6]]) == [21] * 3
    for f in (BB, KB):
        p16, p24 = O.default_perm(f, 16), O.default_perm(f, 24)
        perm = lambda pm: (lambda st: [int(v) for v in O.poseidon2_permute(pm, np.array(st, dtype=np.uint32))])
        for leaf_pm, width, rate in ((p16, 16, 8), (p24, 24, 16)):
            hs = O.poseidon2_hasher(leaf_pm, p16)
            for n in (0, 1, 7, 8, 9, 16, 17, 33, 100):
                row = O.random_matrix(f, 1, max(n, 1), seed=n + width)[0][:n]
                exp = _padding_free_sponge(perm(leaf_pm), width, rate, 8, [int(v) for v in row])
                assert [int(v) for v in O.hash_row(hs, row)] == exp, (f, width, n)
            l, r = O.random_matrix(f, 2, 8, seed=5)
            assert [int(v) for v in O.compress(hs, l, r)] == _truncated_permutation(perm(p16), 2, 8, 16, [[int(v) for v in l], [int(v) for v in r]])


class Xoroshiro128Plus:
    """rand_xoshiro Xoroshiro128Plus::seed_from_u64 (SplitMix64 fills the two state words); next_u32 = upper half of next_u64;
    field samples by rejection of the top 31 bits, the accepted value being the Montgomery word (monty-31/src/monty_31.rs:154-165)."""

    def __init__(self, seed):
        m64, x, s = (1 << 64) - 1, seed, []
        for _ in range(2):
            x = (x + 0x9E3779B97F4A7C15) & m64
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m64
            s.append(z ^ (z >> 31))
        self.s0, self.s1 = s

    def u32(self):
        m64 = (1 << 64) - 1
        rotl = lambda v, k: ((v << k) | (v >> (64 - k))) & m64
        r = (self.s0 + self.s1) & m64
        s1 = self.s1 ^ self.s0
        self.s0 = rotl(self.s0, 24) ^ s1 ^ ((s1 << 16) & m64)
        self.s1 = rotl(s1, 37)
        return r >> 32

    def field(self, p, n):
        out = []
        while len(out) < n:
            v = self.u32() >> 1
            if v < p:
                out.append(v)
        return np.array(out, dtype=np.uint32)


@pytest.mark.parametrize("f,name", [(KB, "koala_bear"), (BB, "baby_bear")])
@pytest.mark.parametrize("width", [16, 24])
def test_poseidon2_rng_constant_kats(f, name, width):
    """The reference's second family of Poseidon2 KATs: constants drawn by Poseidon2::new_from_rng_128 (4 x width initial, 4 x width
    terminal, R_P internal — the order oracle.perm_from_rng and the prove replay rely on) from Xoroshiro128Plus seed 1."""
    kat = json.loads((GOLD / "poseidon2_rng_kat.json").read_text())[f"{name}_{width}"]
    rp = {(BB, 16): 13, (BB, 24): 21, (KB, 16): 20, (KB, 24): 23}[(f, width)]
    rng = Xoroshiro128Plus(1)
    p = O.prime(f)
    init = rng.field(p, 4 * width); term = rng.field(p, 4 * width); internal = rng.field(p, rp)
    pm = O.make_perm(f, width, init, term, internal, monty=True)
    out = O.poseidon2_permute(pm, O.to_monty_arr(f, np.array(kat["input"], dtype=np.uint32)))
    assert [int(v) for v in O.from_monty_arr(f, out)] == kat["expected"]


def _fixture_verifier_setup():
    import stark_verify as V
    rc_i, rc_t, rc_p = FR.fixture_constants()
    pm = O.make_perm(BB, 16, rc_i, rc_t, rc_p, monty=True)
    cfg = dict(hasher=O.poseidon2_hasher(pm, pm), challenger_perm=pm, challenger_width=16, challenger_rate=8, log_blowup=2,
               log_final_poly_len=2, max_log_arity=1, num_queries=2, commit_pow_bits=1, query_pow_bits=1)   # fib_air.rs:134-155
    return V, V.Fld(BB), cfg


def test_verifier_accepts_the_reference_proof_fixture():
    """tests/stark_verify.py (the restated uni-stark + FRI verifier) accepts the proof the reference itself produced and verifies
    (uni-stark/tests/fib_air.rs:401-422) and rejects a wrong statement and every corruption of a proof field."""
    from plonky3_b200.proof_io import proof_from_postcard
    V, f, cfg = _fixture_verifier_setup()
    gold = json.loads((GOLD / "uni_stark_two_adic_v1.json").read_text())
    raw = bytes.fromhex(gold["postcard_hex"])
    V.verify(f, cfg, V.fibonacci_air(), proof_from_postcard(raw), [0, 1, 21])
    with pytest.raises(V.VerifyError):
        V.verify(f, cfg, V.fibonacci_air(), proof_from_postcard(raw), [0, 1, 22])
    # flip one word inside every region of the wire proof: caps, opened values, FRI commitments, witnesses, opened rows, multiproofs,
    # sibling values, final polynomial.  (A flipped length byte must fail in the parser or the shape checks.)
    rejected = 0
    for pos in list(range(2, len(raw) - 1, 29)) + [len(raw) - 6]:
        bad = bytearray(raw); bad[pos] ^= 1
        try:
            V.verify(f, cfg, V.fibonacci_air(), proof_from_postcard(bytes(bad)), [0, 1, 21])
        except (V.VerifyError, ValueError):
            rejected += 1
            continue
        raise AssertionError(f"corrupted byte {pos} was accepted")
    assert rejected >= 38


def test_product_verifier_on_the_reference_fixture():
    """plonky3_b200.verifier.verify — the product's verifier, whose hashing and transcript go through the configuration's MMCS and
    challenger (the GPU ones in tests/test_gpu_prove.py, oracle-backed stand-ins here) — accepts the reference's committed proof,
    rejects a wrong statement and every corruption."""
    from plonky3_b200.field import BabyBear
    from plonky3_b200.verifier import VerificationError, verify
    V, _, cfg = _fixture_verifier_setup()
    gold = json.loads((GOLD / "uni_stark_two_adic_v1.json").read_text())
    raw = bytes.fromhex(gold["postcard_hex"])
    config = V.product_config(BabyBear, cfg)
    verify(config, V.FibonacciAir(), raw, [0, 1, 21])
    with pytest.raises(VerificationError):
        verify(config, V.FibonacciAir(), raw, [0, 1, 22])
    for pos in list(range(2, len(raw) - 1, 29)) + [len(raw) - 6]:
        bad = bytearray(raw); bad[pos] ^= 1
        with pytest.raises(VerificationError):
            verify(config, V.FibonacciAir(), bytes(bad), [0, 1, 21])
    # one encoding per field element: the same final polynomial with its first coefficient written as w + p (>= p, fits in 32 bits
    # for BabyBear) is refused, as MontyField31::deserialize refuses it; truncations never escape as another exception type
    import struct
    off = len(raw) - 1 - 4 - 64
    w = struct.unpack_from("<I", raw, off)[0]
    assert w + BabyBear.P < 1 << 32
    alias = raw[:off] + struct.pack("<I", w + BabyBear.P) + raw[off + 4:]
    with pytest.raises(VerificationError, match="out of range"):
        verify(config, V.FibonacciAir(), alias, [0, 1, 21])
    for cut in (1, 5, 70, 400, 1000, len(raw) - 1):
        with pytest.raises(VerificationError):
            verify(config, V.FibonacciAir(), raw[:cut], [0, 1, 21])


@pytest.mark.parametrize("f", [BB, KB])
def test_fold_row_agrees_with_fold_matrix(f):
    """fri/src/two_adic_pcs.rs tests `fold_matrix_matches_fold_row`: the verifier's Lagrange-form fold_row (plonky3_b200.verifier) on
    every row equals the prover's butterfly-form fold_matrix (oracle), for every arity and several heights."""
    from plonky3_b200.field import BabyBear, KoalaBear
    from plonky3_b200.verifier import Ext, fold_row
    e = Ext(BabyBear if f == BB else KoalaBear)
    for log_arity in (1, 2, 3, 4):
        for log_h in (0, 1, 3, 5):
            n = 1 << (log_h + log_arity)
            vec = O.random_matrix(f, n, 4, seed=7 * log_arity + log_h)
            beta = O.random_matrix(f, 1, 4, seed=99)[0]
            exp = O.fold_matrix(f, vec, log_arity, beta)
            rows = vec.reshape(1 << log_h, 1 << log_arity, 4)
            for i in range(1 << log_h):
                got = fold_row(e, i, log_h, log_arity, e.ec(beta), [e.ec(v) for v in rows[i]])
                assert [e.m(v) for v in got] == [int(v) for v in exp[i]], (log_arity, log_h, i)


def test_verify_multi_batch_mixed_heights():
    """verify_multi_batch_with on a tree over matrices of three heights (injection, merkle_tree.rs:348-): openings built from the
    oracle's tree verify; a wrong row, a wrong digest, a missing or an extra digest do not."""
    import stark_verify as V
    from plonky3_b200.merkle_tree import MerkleTreeError, prune_paths
    hs = O.poseidon2_hasher(O.default_perm(KB, 24), O.default_perm(KB, 16))
    mats = [O.random_matrix(KB, 64, 5, seed=1), O.random_matrix(KB, 16, 3, seed=2), O.random_matrix(KB, 64, 2, seed=3), O.random_matrix(KB, 8, 9, seed=4)]
    layers = O.merkle_tree(hs, mats)
    mm = V.OracleMmcs(hs)
    dims = [(m.shape[1], m.shape[0]) for m in mats]
    for cap_height in (0, 2, 3):
        cap = O.merkle_cap(layers, cap_height)
        idx = [5, 40, 41, 5, 63]
        ov = [[m[i >> (6 - (m.shape[0].bit_length() - 1))] for m in mats] for i in idx]
        paths = np.array([[layers[l][(i >> l) ^ 1] for l in range(6 - cap_height)] for i in idx], dtype=np.uint32).reshape(len(idx), 6 - cap_height, 8)
        pr = prune_paths(idx, paths)
        mm.verify_multi_batch(cap, dims, idx, ov, pr)
        bad_rows = [[r.copy() for r in q] for q in ov]; bad_rows[1][3][0] ^= 1
        for args in ((cap, dims, idx, bad_rows, pr), (cap, dims, idx, ov, pr[:-1]), (cap, dims, idx, ov, np.concatenate([pr, pr[:1]])),
                     (cap, dims, idx, ov, np.concatenate([pr[:1] ^ 1, pr[1:]])), (cap, dims, [6] + idx[1:], ov, pr)):
            with pytest.raises(MerkleTreeError):
                mm.verify_multi_batch(*args)


def test_prove_replay_proof_verifies():
    """The Poseidon2-AIR proof of the CPU replay prover (the one the GPU prover is compared against bit for bit), serialised to the
    reference's wire form and read back, is accepted by the restated verifier; a tampered one is not."""
    import p2_prove_replay as R
    import stark_verify as V
    from plonky3_b200.proof_io import proof_from_postcard
    rng = O.SmallRng(1)
    air = O.air_from_rng(1, rng)
    p16 = O.perm_from_rng(1, 16, rng); p24 = O.perm_from_rng(1, 24, rng)
    inputs = O.SmallRng(1).field(1, (8 << 4) * 16).reshape(-1, 16)
    pr = R.prove(air, p16, p24, inputs, num_queries=5, query_pow_bits=3)
    raw = R.to_wire_proof(pr).to_postcard()
    cfg = R.verifier_config(p16, p24, num_queries=5, query_pow_bits=3)
    f = V.Fld(1)
    V.verify(f, cfg, V.poseidon2_air(air), proof_from_postcard(raw))
    for pos in (40, len(raw) // 3, len(raw) // 2, len(raw) - 40):
        bad = bytearray(raw); bad[pos] ^= 4
        with pytest.raises((V.VerifyError, ValueError)):
            V.verify(f, cfg, V.poseidon2_air(air), proof_from_postcard(bytes(bad)))
    # the product verifier with its own AIR class (constraint folder on the host, no device needed for verification)
    from plonky3_b200.field import KoalaBear
    from plonky3_b200.uni_stark import RoundConstants, VectorizedPoseidon2Air
    from plonky3_b200.verifier import VerificationError, verify
    pair = VectorizedPoseidon2Air(KoalaBear, RoundConstants(np.array(air.beg).reshape(4, 16), np.array(air.part)[: air.rounds_p], np.array(air.end).reshape(4, 16)), None)
    verify(V.product_config(KoalaBear, cfg), pair, raw)
    bad = bytearray(raw); bad[len(raw) // 2] ^= 4
    with pytest.raises(VerificationError):
        verify(V.product_config(KoalaBear, cfg), pair, bytes(bad))


# ---------------------------------------------------------------- Keccak: second, independent formulation
_K_RC = [1, 0x8082, 0x800000000000808a, 0x8000000080008000, 0x808b, 0x80000001, 0x8000000080008081, 0x8000000000008009, 0x8a, 0x88,
         0x80008009, 0x8000000a, 0x8000808b, 0x800000000000008b, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002,
         0x8000000000000080, 0x800a, 0x800000008000000a, 0x8000000080008081, 0x8000000000008080, 0x80000001, 0x8000000080008008]
_K_RHO = [[0, 1, 62, 28, 27], [36, 44, 6, 55, 20], [3, 10, 43, 25, 39], [41, 45, 15, 21, 8], [18, 2, 61, 56, 14]]     # avx512.rs:225-263, [y][x]
_K_PI = [[(0, 0), (1, 1), (2, 2), (3, 3), (4, 4)], [(0, 3), (1, 4), (2, 0), (3, 1), (4, 2)], [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0)],
         [(0, 4), (1, 0), (2, 1), (3, 2), (4, 3)], [(0, 2), (1, 3), (2, 4), (3, 0), (4, 1)]]                            # avx512.rs:268-306
_M64 = (1 << 64) - 1


def _rol(x, r): return ((x << r) | (x >> (64 - r))) & _M64 if r else x


def keccak_f_matrix_form(flat):
    """Keccak-f[1600] as the reference's in-repo vector implementation states it (keccak/src/avx512.rs:40-365): the state as a
    5x5 matrix state[y][x] = flat[5y + x], five SEPARATE steps per round (theta, rho, pi as an explicit index table, chi row by
    row, iota) — structurally unlike the C oracle (flat lanes, fused rho-pi walk)."""
    s = [[int(flat[5 * y + x]) for x in range(5)] for y in range(5)]
    for rnd in range(24):
        par = [s[0][x] ^ s[1][x] ^ s[2][x] ^ s[3][x] ^ s[4][x] for x in range(5)]                  # get_theta_parities
        tp = [(par[(x + 4) % 5], _rol(par[(x + 1) % 5], 1)) for x in range(5)]
        s = [[s[y][x] ^ tp[x][0] ^ tp[x][1] for x in range(5)] for y in range(5)]                   # theta
        s = [[_rol(s[y][x], _K_RHO[y][x]) for x in range(5)] for y in range(5)]                     # rho
        s = [[s[a][b] for (a, b) in row] for row in _K_PI]                                          # pi
        s = [[row[x] ^ (~row[(x + 1) % 5] & _M64 & row[(x + 2) % 5]) for x in range(5)] for row in s]   # chi (ternary 0b11010010)
        s[0][0] ^= _K_RC[rnd]                                                                       # iota
    return [s[y][x] for y in range(5) for x in range(5)]


def _keccak_leaf_second_formulation(row):
    """SerializingHasher<PaddingFreeSponge<KeccakF,25,17,4>> (serializing_hasher.rs:48-59, integers.rs:494-509, sponge.rs:182-216)
    restated on top of the matrix-form permutation."""
    words = [int(row[i]) | (int(row[i + 1]) << 32 if i + 1 < len(row) else 0) for i in range(0, len(row), 2)]
    st = [0] * 25
    for c0 in range(0, len(words), 17):
        blk = words[c0:c0 + 17]
        st[:len(blk)] = blk                                 # overwrite; a partial last block leaves the remaining rate words as they are
        st = keccak_f_matrix_form(st)
    return st[:4]


def _digest_words(d4):
    return np.array([w for v in d4 for w in (v & 0xffffffff, v >> 32)], dtype=np.uint32)


def test_keccak_f_second_formulation_agrees_with_oracle():
    rs = np.random.default_rng(5)
    for _ in range(6):
        st = rs.integers(0, 1 << 63, 25, dtype=np.uint64) * np.uint64(2) + rs.integers(0, 2, 25, dtype=np.uint64)
        assert [int(v) for v in O.keccak_f(st)] == keccak_f_matrix_form(st)
    assert [int(v) for v in O.keccak_f(np.zeros(25, dtype=np.uint64))][:2] == [0xF1258F7940E1DDE7, 0x84D5CCF933C0478A]   # Keccak team KAT, zero state


def test_keccak_mmcs_semantics_second_formulation():
    """leaf hash (all block-boundary cases: odd widths, exactly 34 = one full rate, 35, 300 = config 4) and node compression of the
    Keccak MMCS against the independent restatement."""
    hs = O.keccak_hasher()
    for w in (1, 2, 3, 33, 34, 35, 68, 69, 300):
        row = O.random_matrix(0, 1, w, seed=w)[0]
        assert np.array_equal(O.hash_row(hs, row), _digest_words(_keccak_leaf_second_formulation(row))), w
    l, r = O.random_matrix(0, 1, 8, seed=1)[0], O.random_matrix(0, 1, 8, seed=2)[0]
    # CompressionFunctionFromHasher<_, 2, 4>: sponge over the 8 u64 words left || right (compression.rs:60-70)
    words = [int(l[i]) | int(l[i + 1]) << 32 for i in range(0, 8, 2)] + [int(r[i]) | int(r[i + 1]) << 32 for i in range(0, 8, 2)]
    st = [0] * 25
    st[:8] = words
    assert np.array_equal(O.compress(hs, l, r), _digest_words(keccak_f_matrix_form(st)[:4]))


# ---------------------------------------------------------------- second reference proof fixture: multi-matrix commitment
def _batch_fixture_traces():
    """mul_trace(32, 2) and fib_trace(0, 1, 32) of batch-stark/tests/simple.rs:118-134,317-341 (Montgomery words)."""
    p = 0x78000001
    rows, reps = 32, 2
    w = reps * 3 + 1
    mul = [[0] * w for _ in range(rows)]
    for rep in range(reps):
        a, b = 0, 1
        for i in range(rows):
            mul[i][rep * 3], mul[i][rep * 3 + 1], mul[i][rep * 3 + 2] = a, b, a * b % p
            if i != rows - 1:
                mul[i][w - 1] = b
            a, b = b, (a + b) % p
    fib = [[0, 1]]
    for _ in range(rows - 1):
        fib.append([fib[-1][1], (fib[-1][0] + fib[-1][1]) % p])
    return O.to_monty_arr(0, np.array(mul, dtype=np.uint64)), O.to_monty_arr(0, np.array(fib, dtype=np.uint64))


def batch_fixture_main_cap(lde, commit):
    """main commitment of the batch proof: pcs.commit([(H, mul_trace), (H, fib_trace)]) (batch-stark/src/prover.rs:225-231)."""
    mul, fib = _batch_fixture_traces()
    g = O.generator(0)
    return commit([lde(mul, 2, g), lde(fib, 2, g)])


def test_batch_stark_fixture_main_commitment():
    """The reference's committed batch proof pins the multi-matrix leaf rule (rows of all matrices of a height concatenated in input
    order into one sponge, merkle_tree.rs:312-316) that the unlimited-matrix leaf kernel and the row-sharded commit rely on."""
    gold = json.loads((GOLD / "batch_stark_two_adic_v1.json").read_text())
    rng = O.SmallRng(777)
    pm = O.perm_from_rng(0, 16, rng)                                   # make_two_adic_compat_config(777)
    hs = O.poseidon2_hasher(pm, pm)
    cap = batch_fixture_main_cap(lambda m, bits, s: O.coset_lde_batch(0, m, bits, s, bitrev_out=True),
                                 lambda mats: O.merkle_cap(O.merkle_tree(hs, mats), 1))
    assert cap.tolist() == gold["main_cap"]


# ---------------------------------------------------------------- the product's prove driver on the CPU (device calls answered by the oracle)
def test_prove_driver_host_logic_with_mock_device(monkeypatch):
    """plonky3_b200.uni_stark.prove — the host-side sequencing of commit, quotient, commit_quotient, open, the FRI commit phase
    with its arity schedule, PoW, query openings and the wire serialiser — executed on the CPU with every device call answered by
    the oracle (tests/mock_device.py): the resulting proof equals the replay prover's byte for byte and verifies."""
    import torch
    from types import SimpleNamespace
    import mock_device as M
    import p2_prove_replay as R
    import stark_verify as V
    from plonky3_b200.dft import Radix2DitParallel
    from plonky3_b200.field import KoalaBear
    from plonky3_b200.fri import FriParameters, TwoAdicFriPcs
    from plonky3_b200.merkle_tree import MerkleTreeMmcs
    from plonky3_b200.poseidon2 import Poseidon2
    from plonky3_b200.uni_stark import RoundConstants, VectorizedPoseidon2Air, prove
    from plonky3_b200.verifier import verify
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)      # the driver's span timers synchronise the device
    rng = O.SmallRng(1)
    oair = O.air_from_rng(KB, rng)
    o16 = O.perm_from_rng(KB, 16, rng); o24 = O.perm_from_rng(KB, 24, rng)
    mk = lambda pm: Poseidon2.new(KoalaBear, pm.width, np.array(pm.rc_init)[: 4 * pm.width].reshape(4, pm.width),
                                  np.array(pm.rc_term)[: 4 * pm.width].reshape(4, pm.width), np.array(pm.rc_int)[: pm.rounds_p], monty=True)
    p16, p24 = mk(o16), mk(o24)
    for log_n, nq, pow_bits in [(3, 4, 3), (5, 9, 5)]:
        gpu = M.MockGpu()
        mmcs = MerkleTreeMmcs.poseidon2(p16, p24, cap_height=3, gpu=gpu)
        pcs = TwoAdicFriPcs(Radix2DitParallel(KoalaBear, gpu), mmcs, FriParameters(1, 0, 3, nq, 0, pow_bits, mmcs))
        config = SimpleNamespace(pcs=pcs, initialise_challenger=lambda: M.MockChallenger(o24))
        air = VectorizedPoseidon2Air(KoalaBear, RoundConstants(np.array(oair.beg).reshape(4, 16), np.array(oair.part)[: oair.rounds_p],
                                                               np.array(oair.end).reshape(4, 16)), gpu)
        inputs = O.SmallRng(1).field(KB, (8 << log_n) * 16).reshape(-1, 16)
        trace = air.generate_trace_rows(torch.from_numpy(inputs.view(np.int32)))
        proof = prove(config, air, trace)
        exp = R.prove(oair, o16, o24, inputs, num_queries=nq, query_pow_bits=pow_bits)
        raw = proof.to_postcard()
        assert raw == R.to_wire_proof(exp).to_postcard()
        assert {"coset_lde_batch", "merkle_commit", "p2air_quotient", "columnwise_dot", "rowwise_dot", "open_reduce", "fri_fold"} <= set(gpu.calls)
        verify(V.product_config(KoalaBear, R.verifier_config(o16, o24, num_queries=nq, query_pow_bits=pow_bits)), air, raw)


@pytest.mark.parametrize("log_blowup,log_final_poly_len,max_log_arity,cap_height", [(1, 0, 1, 0), (2, 1, 3, 1), (1, 3, 4, 2)])
def test_pcs_round_trip_host_logic_with_mock_device(monkeypatch, log_blowup, log_final_poly_len, max_log_arity, cap_height):
    """fri/tests/pcs.rs on the CPU: Pcs::commit -> open -> verify over two rounds with matrices of three heights in one batch and one
    or two opening points, the product's host code throughout (device calls answered by the oracle); tampering is rejected.  The
    GPU suite runs the same scenario with the real kernels (tests/test_gpu_prove.py)."""
    import copy
    import torch
    import mock_device as M
    import stark_verify as V
    from plonky3_b200.dft import Radix2DitParallel
    from plonky3_b200.field import KoalaBear
    from plonky3_b200.fri import FriParameters, TwoAdicFriPcs
    from plonky3_b200.merkle_tree import MerkleTreeMmcs
    from plonky3_b200.poseidon2 import default_poseidon2
    from plonky3_b200.verifier import VerificationError
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    f = KoalaBear
    gpu = M.MockGpu()
    p16, p24 = default_poseidon2(f, 16), default_poseidon2(f, 24)
    o16, o24 = O.default_perm(KB, 16), O.default_perm(KB, 24)
    mmcs = MerkleTreeMmcs.poseidon2(p16, p24, cap_height=cap_height, gpu=gpu)
    pcs = TwoAdicFriPcs(Radix2DitParallel(f, gpu), mmcs, FriParameters(log_blowup, log_final_poly_len, max_log_arity, 6, 1, 2, mmcs))
    vm = V.OracleMmcs(O.poseidon2_hasher(o24, o16))
    pcs_v = TwoAdicFriPcs(Radix2DitParallel(f, gpu), vm, FriParameters(log_blowup, log_final_poly_len, max_log_arity, 6, 1, 2, vm))
    shapes = [[(5, 3), (8, 2)], [(8, 1), (7, 4), (5, 2)]]
    commits, datas = [], []
    for r, rnd in enumerate(shapes):
        evals = [((f.ONE, d), torch.from_numpy(O.random_matrix(KB, 1 << d, w, seed=10 * r + d + w).view(np.int32))) for d, w in rnd]
        c, pd = pcs.commit(evals)
        commits.append(c); datas.append(pd)

    def head(ch):
        for c in commits:
            ch.observe_cap(c)
        return ch.sample_algebra_element(), ch.sample_algebra_element()
    ch = M.MockChallenger(o24)
    zeta, zeta2 = head(ch)
    points = [[[zeta], [zeta, zeta2]], [[zeta2], [zeta], [zeta, zeta2]]]
    opened, proof = pcs.open(list(zip(datas, points)), ch)

    def check(opened_values, prf):
        chv = M.MockChallenger(o24)
        head(chv)
        claims = [(c, [((f.ONE, d), list(zip(pts, ov))) for (d, _), pts, ov in zip(rnd, rpts, rov)])
                  for c, rnd, rpts, rov in zip(commits, shapes, points, opened_values)]
        pcs_v.verify(claims, prf, chv)
    check(opened, proof)
    bad = copy.deepcopy(opened); bad[1][1][0][2][1] ^= 1
    with pytest.raises(VerificationError):
        check(bad, proof)
    for mutate in (lambda p: p["input_openings"][0]["opened_values"][3][1].__setitem__(0, int(p["input_openings"][0]["opened_values"][3][1][0]) ^ 1),
                   lambda p: p["input_openings"][1]["proof"].__setitem__((0, 0), int(p["input_openings"][1]["proof"][0, 0]) ^ 1),
                   lambda p: p["commit_phase_openings"][0]["sibling_values"][2].__setitem__((0, 0), int(p["commit_phase_openings"][0]["sibling_values"][2][0, 0]) ^ 1),
                   lambda p: p["final_poly"].__setitem__((0, 0), int(p["final_poly"][0, 0]) ^ 1),
                   lambda p: p["commit_phase_openings"].pop()):
        prf = copy.deepcopy(proof)
        mutate(prf)
        with pytest.raises(VerificationError):
            check(opened, prf)


def test_get_evaluations_on_domain_host_logic_with_mock_device():
    """TwoAdicFriPcs::get_evaluations_on_domain (two_adic_pcs.rs:376-403) through the product's host code on the mock device: the fast
    path (prefix of the committed LDE) and the slow path (coset iDFT, truncate, zero-pad, coset DFT onto a foreign coset) against a
    direct oracle LDE of the same polynomials onto that coset."""
    import torch
    import mock_device as M
    from plonky3_b200.dft import Radix2DitParallel
    from plonky3_b200.field import KoalaBear as f
    from plonky3_b200.fri import FriParameters, TwoAdicFriPcs
    from plonky3_b200.merkle_tree import MerkleTreeMmcs
    from plonky3_b200.poseidon2 import default_poseidon2
    gpu = M.MockGpu()
    mmcs = MerkleTreeMmcs.poseidon2(default_poseidon2(f, 16), default_poseidon2(f, 24), cap_height=0, gpu=gpu)
    pcs = TwoAdicFriPcs(Radix2DitParallel(f, gpu), mmcs, FriParameters(1, 0, 1, 2, 0, 0, mmcs))
    evals = O.random_matrix(KB, 1 << 5, 3, seed=4)
    _, pd = pcs.commit([((f.ONE, 5), torch.from_numpy(evals.view(np.int32)))])
    fast = pcs.get_evaluations_on_domain(pd, 0, (f.generator, 6)).bit_reverse_rows()
    assert np.array_equal(fast.numpy().view(np.uint32), O.coset_lde_batch(KB, evals, 1, f.generator, bitrev_out=True))
    for shift, log_size in [(f.mul(f.generator, f.two_adic_generator(7)), 6), (f.to_monty(5), 7), (f.to_monty(7), 5)]:
        got = pcs.get_evaluations_on_domain(pd, 0, (shift, log_size)).to_row_major_matrix()
        coeffs = O.idft_batch(KB, evals)
        padded = np.zeros((1 << log_size, 3), dtype=np.uint32)
        padded[: min(32, 1 << log_size)] = coeffs[: min(32, 1 << log_size)]
        assert np.array_equal(np.asarray(got.numpy() if hasattr(got, "numpy") else got).view(np.uint32), O.coset_dft_batch(KB, padded, shift)), (shift, log_size)
